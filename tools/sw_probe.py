"""Device-resident throughput of the optional sliding-window mode next to the bit-exact mode (same plan API, same
inputs, CUDA events).  Regimes: uniform noise +-16 (6 full iterations, the headline regime) and coded blocks.
usage: python tools/sw_probe.py [--blocks 42624] [--K 6144 ...]"""
import argparse
import json
import sys

import numpy as np

sys.path.insert(0, ".")


def main():
    import torch
    from openair4g_b200 import capi
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=42624)
    ap.add_argument("--K", type=int, nargs="*", default=[6144])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--coded", action="store_true")
    ap.add_argument("--same", action="store_true")
    a = ap.parse_args()
    capi.init_td16()
    for K in a.K:
        B = a.blocks if K >= 2048 else min(65536, a.blocks * (6144 // max(K, 512)))
        row = 3 * K + 12
        g = torch.Generator(device="cuda")
        g.manual_seed(4321)
        regimes = [("noise16", torch.randint(-16, 17, (B, row), dtype=torch.int16, device="cuda", generator=g), 7)]
        if a.coded:
            import bench
            nd = 256
            for name, sig in (("waterfall", 1.08), ("clean", 0.5)):
                ys = bench.coded_inputs(K, nd, sig, 70)[0][:, :row]       # the product's own TX chain
                regimes.append((name, torch.from_numpy(ys).cuda().repeat((B + nd - 1) // nd, 1)[:B].contiguous(), None))
                if name == "waterfall" and a.same:
                    for j in range(3):
                        regimes.append(("waterfall_same%d" % j, torch.from_numpy(ys[j:j + 1]).cuda().repeat(B, 1).contiguous(), None))
        for name, y, want in regimes:
            out = torch.zeros((B, K // 8), dtype=torch.uint8, device="cuda")
            st = torch.zeros(B, dtype=torch.uint8, device="cuda")
            plan = capi.DevPlan(B, K, 6, 1)
            stream = torch.cuda.current_stream().cuda_stream
            res = {"K": K, "blocks": B, "regime": name}
            for mode, flags in (("exact", 0), ("sliding_window", capi.BATCH_SLIDING_WINDOW)):
                plan.set_mode(flags)
                for _ in range(2):
                    plan.decode(y.data_ptr(), row, out.data_ptr(), K // 8, st.data_ptr(), stream)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps):
                    plan.decode(y.data_ptr(), row, out.data_ptr(), K // 8, st.data_ptr(), stream)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.steps
                if want is not None:
                    assert (st == want).all(), (mode, st[:16])
                res[mode] = {"ms": round(ms, 3), "gbit_s": round(B * K / ms / 1e6, 2), "mean_status": round(float(st.float().mean()), 3),
                             "failed": int((st > 6).sum())}
            plan.close()
            print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
