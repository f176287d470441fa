"""Latency of one subframe's transport block (host buffers in -> bytes out, bench.subframe_latency) and of single code
blocks in the bit-exact mode and in the optional sliding-window mode."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")


def main():
    import torch
    import bench
    from openair4g_b200 import capi
    capi.init_td16()
    torch.zeros(1, device="cuda")
    for name, (tbs, G, Qm) in (("dlsim_100PRB_MCS28", (75376, 90000, 6)), ("ulsim_25PRB_MCS16", (7736, 14400, 4)),
                               ("ul_100PRB_MCS16", (30576, 57600, 4))):
        row = {"config": name}
        for mode, flags in (("exact", 0), ("sliding_window", capi.BATCH_SLIDING_WINDOW)):
            r = bench.subframe_latency(capi, tbs, G, Qm, 6, flags=flags)
            row[mode] = {"clean_ms": round(r["clean"], 3), "full_iterations_ms": round(r["full_iterations"], 3)}
            row["code_blocks"], row["K"] = r["code_blocks"], r["K"]
        print(json.dumps(row), flush=True)
    rng = np.random.default_rng(1)
    for K in (40, 512, 1024, 2048, 6144):
        for nblk in (1, 16, 148, 592):
            blocks = [{"y": rng.integers(-16, 17, size=3 * K + 12).astype(np.int16), "K": K, "max_iterations": 6, "crc_type": 1}
                      for _ in range(nblk)]
            row = {"K": K, "blocks": nblk}
            for mode, flags in (("exact", 0), ("sliding_window", capi.BATCH_SLIDING_WINDOW)):
                ts = []
                for i in range(12):
                    t0 = time.perf_counter()
                    capi.decode_batch(blocks, flags=flags)
                    ts.append(time.perf_counter() - t0)
                row[mode + "_ms"] = round(1e3 * sorted(ts[2:])[len(ts[2:]) // 2], 3)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
