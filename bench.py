#!/usr/bin/env python3
"""Headline benchmark: turbo-decoded information Mbit/s at K=6144, 6 iterations (BASELINE.json).

One "step" = one full decode of a batch of code blocks (demux, 12 MAP passes, QPP exchanges,
hard decision + CRC per iteration) on every GPU.  See DESIGN.md "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
  python bench.py --impl reference ...                             # the reference's own CPU decoder
  torchrun ... bench.py --gpus N ...                               # one rank per GPU, weak scaling
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_BITS = 6144
MAX_ITER = 6
CRC_TYPE = 1           # CRC24B
METRIC = "turbo_decoded_info_mbit_per_s_K6144_6iter"
WORKLOAD = ("isolated turbo-decoder batch (BASELINE configs[2]): K=6144, max_iterations=6, CRC24B early exit "
            "enabled, noise regime (uniform +-16 int16 LLRs, CRC never passes -> exactly 6 iterations / 12 MAP passes)")


def algorithmic_bytes_per_block(K):
    return 2 * (3 * K + 12) + K // 8 + 1      # read y once, write bytes + status


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.p = None
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _pump(self):
        for line in self.p.stdout:
            self.lines.append((time.perf_counter(), line))

    def wait_first(self, timeout=10.0):
        t0 = time.perf_counter()
        while self.p is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        time.sleep(0.12)                       # let the sample that covers the end of the region arrive
        self.p.terminate()
        t_mark = getattr(self, "t_mark", 0.0)
        sel = [l for (ts, l) in self.lines if t_mark <= ts <= t_end + 0.12]
        if not sel:
            sel = [l for (_, l) in self.lines[-1:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in sel:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/): the reference decoder on host cores
# ---------------------------------------------------------------------------------------
def cpu_decode_rate(y_blocks, total_blocks, threads):
    """Decodes `total_blocks` code blocks (cycling over the distinct inputs y_blocks) on
    `threads` host threads with the compiled reference (oracle/_ref) when present, else the
    oracle port.  Returns (Mbit/s, kind, seconds, per-block results of the distinct inputs)."""
    import numpy as np
    from oracle import loader
    R = loader.ref()
    nd = y_blocks.shape[0]
    results = [None] * nd
    if R is not None:
        kind = "reference"
        loader.ref_decode_batch(y_blocks[:1], K_BITS, MAX_ITER, CRC_TYPE)      # warm tables outside the timing
        out, ret, dt = loader.ref_decode_batch(y_blocks, K_BITS, MAX_ITER, CRC_TYPE, total=total_blocks, threads=threads)
        total_blocks = max(total_blocks, nd)
        results = [(out[i], int(ret[i])) for i in range(nd)]
    else:
        kind = "port"
        P = loader.port()
        reps = (total_blocks + nd - 1) // nd
        total_blocks = reps * nd
        stride = y_blocks.shape[1]
        out = np.zeros((nd, K_BITS // 8 + 8), dtype=np.uint8)
        ret = np.zeros(nd, dtype=np.uint8)
        yy = np.ascontiguousarray(y_blocks)
        t0 = time.perf_counter()
        for _ in range(reps):
            P.orc_turbo_decoder16_batch(yy, stride, out, out.shape[1], ret, nd, K_BITS, MAX_ITER, CRC_TYPE, threads)
        dt = time.perf_counter() - t0
        results = [(out[i, :K_BITS // 8].copy(), int(ret[i])) for i in range(nd)]
    return total_blocks * K_BITS / dt / 1e6, kind, dt, results


def coded_inputs(K, nd, sigma_over_A, seed, A=8):
    """nd distinct code blocks (random payload + CRC24B, 36.212 turbo code, BPSK LLR = A(2b-1) + sigma N(0,1), SURVEY 8d
    config 3) from the product's own TX chain (openair4g_b200/sim/txchain.py)."""
    import numpy as np
    from openair4g_b200.sim import txchain
    rng = np.random.default_rng(seed)
    payload = rng.integers(0, 2, size=(nd, K - 24)).astype(np.uint8)
    c = np.concatenate([payload, txchain.crc24b(payload)], axis=1)
    bits = txchain.turbo_encode(c).astype(np.int64)
    y = A * (2 * bits - 1) + np.rint(sigma_over_A * A * rng.standard_normal(bits.shape)).astype(np.int64)
    return np.clip(y, -32768, 32767).astype(np.int16), np.packbits(c, axis=1)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's CPU decoder, all host threads, bounded sample per step."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import numpy as np
    threads = host_threads()
    rng = np.random.default_rng(1234)
    nd = 64
    y = rng.integers(-16, 17, size=(nd, 3 * K_BITS + 12)).astype(np.int16)
    per_step = max(threads * 256, 1024)     # ~0.1 s per step on 16 threads
    for _ in range(args.warmup):
        cpu_decode_rate(y, max(threads, 64), threads)
    t_tot, kind = 0.0, "port"
    for _ in range(args.steps):
        _, kind, dt, _ = cpu_decode_rate(y, per_step, threads)
        t_tot += dt
    val = args.steps * per_step * K_BITS / t_tot / 1e6
    sample = "%d blocks/step (K=6144, noise regime, 6 iterations) on %d host threads" % (per_step, threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mbit/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "blocks_per_step": per_step},
            "cpu_baseline": {"value": val, "unit": "Mbit/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# this repo's engine
# ---------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from openair4g_b200 import capi
    capi.init_td16()

    B, K = args.blocks, (args.K if args.llr8 else K_BITS)
    if args.llr8:
        return run_llr8(args, capi, B, K, rank, world, dist)
    row = 3 * K + 12
    g = torch.Generator(device="cuda")
    g.manual_seed(1000 + rank)
    y_dev = torch.randint(-16, 17, (B, row), dtype=torch.int16, device="cuda", generator=g)
    out_dev = torch.zeros((B, K // 8), dtype=torch.uint8, device="cuda")
    st_dev = torch.zeros(B, dtype=torch.uint8, device="cuda")
    plan = capi.DevPlan(B, K, MAX_ITER, CRC_TYPE, llr8=1 if args.llr8 else 0)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        return plan.decode(y_dev.data_ptr(), row, out_dev.data_ptr(), K // 8, st_dev.data_ptr(), stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    plan.profile(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.wait_first()
    launches0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = capi.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    prof_ms, prof_cnt = plan.profile(False, fetch=True)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    st = st_dev.cpu().numpy()
    if not (st == MAX_ITER + 1).all():
        raise SystemExit("bench.py: noise-regime blocks must all run the full 6 iterations (status 7); got %s"
                         % np.unique(st))
    value = world * B * K * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the host-buffer C ABI (pinned host memory in, host memory out) ----
    Be = args.e2e_blocks
    y_pin = torch.empty((Be, row), dtype=torch.int16).pin_memory()
    y_pin.copy_(y_dev[:Be].cpu())
    # Every step is one oai_turbo_submit_batch (copies the step's inputs from page-locked host memory, decodes, copies the
    # decoded bytes and status back) and one oai_turbo_wait.  Inside a call the batch is pipelined in parts (input copy of
    # part i+1 overlaps the decode of part i).  One call at a time by default; --e2e-in-flight 2 keeps two batches in flight
    # (submit of step i+1 before the wait of step i) -- measured 8.5-8.9 Gbit/s against 8.0-8.1 in most runs but 3.4 in one
    # of six (the small metadata copies of one batch queue behind the large input copies of the other on the shared copy
    # engine), so it is not the default.
    args.e2e_serial = args.e2e_in_flight < 2
    calls = [capi.HostBatchCall(y_pin.numpy(), K, MAX_ITER, CRC_TYPE) for _ in range(1 if args.e2e_serial else 2)]
    for c in calls:
        for _ in range(2):
            c.run()
    call = calls[0]
    barrier()
    t0 = time.perf_counter()
    if args.e2e_serial:
        for _ in range(args.steps):
            out_h, st_h = call.run()
    else:
        pending = None
        for i in range(args.steps):
            c = calls[i & 1]
            h = c.submit()
            if pending is not None:
                pending[0].wait(pending[1])
            pending = (c, h)
        out_h, st_h = pending[0].wait(pending[1])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    e2e_val = world * Be * K * args.steps / dt / 1e6
    same = bool((out_h == out_dev[:Be].cpu().numpy()).all() and (st_h == st[:Be]).all())
    if not same:
        raise SystemExit("bench.py: host-buffer path and device-resident path disagree")

    # ---- side measurement: early-exit regimes of config 3 (device-resident, same plan, same timing rules) ----
    regimes = None
    if world == 1 and not args.no_regimes:
        regimes = {}
        for name, sig in (("clean", 0.5), ("waterfall", 1.08)):
            nd = 64
            ys, info = coded_inputs(K, nd, sig, 4242)
            idx = (torch.arange(B, device="cuda") * 29) % nd
            y_r = torch.from_numpy(ys).cuda()[idx].contiguous()
            out_r = torch.zeros((B, K // 8), dtype=torch.uint8, device="cuda")
            st_r = torch.zeros(B, dtype=torch.uint8, device="cuda")

            def step_r():
                return plan.decode(y_r.data_ptr(), row, out_r.data_ptr(), K // 8, st_r.data_ptr(), stream)
            for _ in range(3):
                step_r()
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(args.steps):
                step_r()
            r1.record()
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1)
            hist = torch.bincount(st_r.long(), minlength=MAX_ITER + 2).tolist()
            ok_mask = (st_r <= MAX_ITER).cpu().numpy()
            good = bool((out_r.cpu().numpy()[ok_mask] == info[idx.cpu().numpy()][ok_mask]).all())
            regimes[name] = {"value": B * K * args.steps / (rms * 1e-3) / 1e6, "unit": "Mbit/s", "sigma_over_A": sig,
                             "return_value_histogram": {str(i): h for i, h in enumerate(hist) if h},
                             "mean_iterations": float(sum(min(i, MAX_ITER) * h for i, h in enumerate(hist)) / max(sum(hist), 1)),
                             "crc_passing_blocks_equal_transmitted_bytes": good,
                             "distinct_blocks": nd}
            del y_r, out_r, st_r

    # ---- side measurement: TX mirror (encoder + sub-block interleaver + rate matching), device pointers ----
    tx_side = None
    if world == 1 and not args.no_regimes:
        G_tx = 11520                                            # 100 PRB MCS16 uplink share of one K=6144 block (configs[3])
        c_dev = torch.randint(0, 256, (B, K // 8), dtype=torch.uint8, device="cuda")
        e_dev = torch.zeros((B, G_tx), dtype=torch.uint8, device="cuda")
        descs = (capi.TxDesc * B)()
        for i in range(B):
            d = descs[i]
            d.c = c_dev.data_ptr() + i * (K // 8); d.e = e_dev.data_ptr() + i * G_tx
            d.K = K; d.G = G_tx; d.Nsoft = 1827072; d.C = 1; d.Mdlharq = 8; d.Kmimo = 1; d.Qm = 2; d.Nl = 1
        torch.cuda.synchronize()
        best = None
        for _ in range(4):                                      # the call is synchronous: wall clock around it, best of 4
            tx0 = time.perf_counter()
            rc = capi.lib.oai_turbo_tx_batch(descs, B, capi.TX_DEVICE_POINTERS, -1)
            tx_dt = time.perf_counter() - tx0
            if rc:
                raise SystemExit("bench.py: oai_turbo_tx_batch failed: " + capi.last_error())
            best = tx_dt if best is None else min(best, tx_dt)
        tx_side = {"value": B * K / best / 1e6, "unit": "Mbit/s", "blocks": B, "E": G_tx, "ms": best * 1e3,
                   "api": "oai_turbo_tx_batch, device pointers (descriptor staging inside the call)"}
        del c_dev, e_dev

    # final result gather (outside every timed region): per-rank block / bit / status counts
    from openair4g_b200 import sharding
    recs = sharding.gather_results(st_dev, B * K, dist)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    per_rank = [{"blocks": int(r[0]), "info_bits": int(r[1]), "status_7": int(r[2 + MAX_ITER + 1])} for r in recs]

    # ---- roofline of the dominant kernel (k_map16: one MAP pass over all blocks per launch) ----
    peak, peak_src = measured_peaks()
    n_map = prof_cnt[1]
    map_ms = prof_ms[1] / max(n_map, 1)
    bytes_per_launch = B * algorithmic_bytes_per_block(K) / (2.0 * MAX_ITER)
    achieved = bytes_per_launch / (map_ms * 1e-3) / 1e9 if n_map else 0.0
    kernel_ms_total = sum(prof_ms)
    # dram__bytes_read.sum + dram__bytes_write.sum per k_map16 launch of 23680 blocks, ncu --set full capture
    # profiles/r1o_ncu_full_all_summary.txt (mean of the three captured launches; unchanged since r1h), scaled to this batch size
    traffic = 1.444e9 * B / 23680.0 if K == K_BITS else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_map16", "avg_launch_ms": map_ms, "launches_timed": n_map,
                "share_of_step": prof_ms[1] / kernel_ms_total if kernel_ms_total else None,
                "kernel_ms": {"demux": prof_ms[0], "map": prof_ms[1], "x1": prof_ms[2], "x2": prof_ms[3]},
                "peak_source": peak_src,
                "actual_dram_gbs": (traffic / (map_ms * 1e-3) / 1e9) if (traffic and n_map) else None,
                "note": "algorithmic bytes per launch = blocks x (2(3K+12)+K/8+1)/12 (SURVEY 8d: y read once, bytes + "
                        "status written once, spread over the 12 MAP passes); the decoder streams its per-block state "
                        "through HBM on every pass, so the traffic actually moved per launch (`traffic`, ncu) is ~11x "
                        "the algorithmic figure; `actual_dram_gbs` = traffic / launch time.  The kernel is co-limited "
                        "by the ALU pipe (VIADDMNMX issues every 2nd clock), the L1/shared-memory data pipe and HBM "
                        "latency (ncu: 62 % / 60 % / 50 % of peak), see DESIGN.md and int_simd"}
    # integer-SIMD view: SURVEY 8(d) counts 123 int16 ops / info bit / MAP pass; the peak is the MEASURED issue rate of
    # the k_map16 instruction mix (tools/int16_peak.cu, profiles/int16_peak.json), scaled by the sampled SM clock
    ops = B * K * 123.0
    try:
        ipk = json.load(open(os.path.join(ROOT, "profiles", "int16_peak.json")))
        int_peak = ipk["int16_ops_peak_gops_kmap16_mix"] * 1e9 * ((clocks and clocks["sm_mhz"]) or 1965.0) / 1965.0
        int_note = "peak = measured issue rate of the 2 VIADDMNMX + VIADD + IMAD mix (tools/int16_peak.cu) at the sampled SM clock"
    except Exception:
        int_peak = 148 * 128 * 2 * 1965.0e6
        int_note = "paper peak = 148 SM x 128 lanes x 2 halfwords x 1965 MHz (profiles/int16_peak.json missing)"
    int_simd = {"achieved_int16_gops": ops / (map_ms * 1e-3) / 1e9 if n_map else None,
                "measured_peak_int16_gops": int_peak / 1e9, "note": int_note}
    if int_simd["achieved_int16_gops"]:
        int_simd["frac"] = int_simd["achieved_int16_gops"] / int_simd["measured_peak_int16_gops"]
    # issue-slot view from executed instructions: ncu (profiles/r1h) counts 207.8 M warp instructions per k_map16
    # launch of 23680 blocks at K=6144 = 8776 per block (the fast path needs 44 thread instructions per bit and pass,
    # not SURVEY's nominal 123 ops); the SM issues at most 4 warp instructions per clock
    if K == K_BITS and n_map:
        sm_hz = ((clocks and clocks["sm_mhz"]) or 1965.0) * 1e6
        int_simd["executed_warp_instr_per_block_pass"] = 8776
        int_simd["issue_frac"] = B * 8776.0 / (map_ms * 1e-3) / (148 * 4 * sm_hz)

    cpu = None
    if world == 1 and not args.no_cpu:
        threads = host_threads()
        nd = 64
        ys = y_dev[:nd].cpu().numpy()
        total = args.cpu_blocks
        val, kind, secs, res = cpu_decode_rate(ys, total, threads)
        ok = all(r is not None and r[1] == int(st[i]) and (r[0] == out_dev[i].cpu().numpy()).all() for i, r in enumerate(res))
        cpu = {"value": val, "unit": "Mbit/s", "cores": threads, "kind": kind,
               "sample": "%d blocks of the same workload (first %d distinct inputs of the GPU batch, cycled) in %.2f s; "
                         "GPU output bit-exact vs this CPU run: %s" % (total, nd, secs, ok)}
        if not ok:
            raise SystemExit("bench.py: GPU result differs from the CPU reference on the sampled blocks")

    line = {"metric": METRIC, "value": value, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "blocks_per_gpu_per_step": B, "e2e_blocks_per_gpu_per_step": Be,
                       "l2": "inputs larger than L2 (%.0f MB of LLRs per GPU per step, workspace %.0f MB)"
                             % (B * row * 2 / 1e6, B * 6 * K * 2 / 1e6),
                       "timing": "CUDA events on the launching stream, barrier + synchronize both sides, max over ranks",
                       "sharding": "independent code blocks, one shard per rank, no data-path collective", "per_rank": per_rank},
            "roofline": roofline, "int_simd": int_simd, "cpu_baseline": cpu, "early_exit_regimes": regimes, "tx_mirror": tx_side,
            "e2e": {"value": e2e_val, "unit": "Mbit/s", "h2d_bytes_per_step": call.h2d_bytes,
                    "d2h_bytes_per_step": call.d2h_bytes, "ms_per_step": 1e3 * dt / args.steps,
                    "in_flight": 1 if args.e2e_serial else 2,
                    "api": "oai_turbo_submit_batch + oai_turbo_wait per step, page-locked host input and output buffers; "
                           "bound by the PCIe copy of 36.9 KB of int16 LLRs per 6144 decoded bits"},
            "gpu_launches": launches, "clocks": clocks}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_llr8(args, capi, B, K, rank, world, dist):
    """Side measurement (BASELINE configs[4]): 8-bit decoder throughput, device-resident, same timing rules."""
    import torch
    row = 3 * K + 12 + 4
    g = torch.Generator(device="cuda")
    g.manual_seed(1000 + rank)
    y_dev = torch.randint(-16, 17, (B, row), dtype=torch.int16, device="cuda", generator=g)
    out_dev = torch.zeros((B, K // 8), dtype=torch.uint8, device="cuda")
    st_dev = torch.zeros(B, dtype=torch.uint8, device="cuda")
    plan = capi.DevPlan(B, K, MAX_ITER, CRC_TYPE, llr8=1)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(args.warmup):
        plan.decode(y_dev.data_ptr(), row, out_dev.data_ptr(), K // 8, st_dev.data_ptr(), stream)
    torch.cuda.synchronize()
    plan.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.decode(y_dev.data_ptr(), row, out_dev.data_ptr(), K // 8, st_dev.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof_ms, _ = plan.profile(False, fetch=True)
    print(json.dumps({"metric": "turbo_decoded_info_mbit_per_s_8bit_decoder", "value": B * K * args.steps / (ms * 1e-3) / 1e6,
                      "kernel_ms": {"demux": prof_ms[0], "map": prof_ms[1], "x1": prof_ms[2], "x2": prof_ms[3]},
                      "unit": "Mbit/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                      "dtype": "int8", "data": "synthetic",
                      "config": {"workload": "8-bit decoder, K=%d, max_iterations=6, noise regime" % K, "blocks": B},
                      "status_hist": torch.bincount(st_dev.long()).nonzero().flatten().tolist()}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=42624,
                    help="code blocks per GPU per step (device-resident); 42624 = 3 full waves of the MAP kernel "
                         "(148 SMs x 6 resident CTAs x 16 blocks)")
    ap.add_argument("--e2e-blocks", type=int, default=42624, help="code blocks per GPU per step (host-buffer API)")
    ap.add_argument("--cpu-blocks", type=int, default=262144,
                    help="bounded CPU-baseline sample (blocks): ~4 s of wall time on 16 host threads (~60 s of CPU work)")
    ap.add_argument("--e2e-in-flight", type=int, default=1, choices=[1, 2], help="host-buffer batches in flight in the e2e loop")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-regimes", action="store_true", help="skip the clean / waterfall side measurements")
    ap.add_argument("--llr8", action="store_true", help="measure the 8-bit decoder (BASELINE configs[4]) instead")
    ap.add_argument("--K", type=int, default=K_BITS, help="block size for --llr8 / side measurements")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
