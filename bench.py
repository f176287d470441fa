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


def bench_config(B, Be, K):
    """the workload description both arms print (the driver compares the two `config` objects)"""
    row = 3 * K + 12
    return {"workload": WORKLOAD, "blocks_per_gpu_per_step": B, "e2e_blocks_per_gpu_per_step": Be,
            "l2": "inputs larger than L2 (%.0f MB of LLRs per GPU per step, workspace %.0f MB)" % (B * row * 2 / 1e6, B * 6 * K * 2 / 1e6)}


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
    per_step = args.blocks                  # the repo arm's batch (42624 blocks: ~0.6 s per step on 16 host threads)
    for _ in range(args.warmup):
        cpu_decode_rate(y, max(threads * 16, 64), threads)      # warm-up steps are short (tables, thread pool, caches)
    t_tot, kind = 0.0, "port"
    for _ in range(args.steps):
        _, kind, dt, _ = cpu_decode_rate(y, per_step, threads)
        t_tot += dt
    val = args.steps * per_step * K_BITS / t_tot / 1e6
    sample = "%d blocks/step (K=6144, noise regime, 6 iterations) on %d host threads" % (per_step, threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mbit/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": bench_config(per_step, per_step, K_BITS),
            "cpu_baseline": {"value": val, "unit": "Mbit/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


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
    # part i+1 overlaps the decode of part i on three compute streams).  Measured twice: one call at a time (`serial`), and
    # with two batches in flight (submit of step i+1 before the wait of step i, two handles -- what a streaming receiver
    # does with an asynchronous submit/wait API): the link then stays busy across step boundaries.  `--e2e-in-flight 1`
    # makes the serial figure the headline one.
    calls = [capi.HostBatchCall(y_pin.numpy(), K, MAX_ITER, CRC_TYPE) for _ in range(2)]
    for c in calls:
        for _ in range(2):
            c.run()

    def e2e_loop(in_flight):
        barrier()
        t0 = time.perf_counter()
        if in_flight < 2:
            for _ in range(args.steps):
                res = calls[0].run()
        else:
            pending = None
            for i in range(args.steps):
                c = calls[i & 1]
                h = c.submit()
                if pending is not None:
                    pending[0].wait(pending[1])
                pending = (c, h)
            res = pending[0].wait(pending[1])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), res
    steps_timed = args.steps
    if args.e2e_in_flight >= 2:
        # untimed: the second in-flight batch object (device workspace, streams) is created on first use -- in round 1 that
        # allocation fell into the timed loop and showed up as a bimodal "two in flight" figure
        args.steps = 4
        e2e_loop(2)
        args.steps = steps_timed
    dt_serial, (out_h, st_h) = e2e_loop(1)
    dt, (out_h2, st_h2) = e2e_loop(2) if args.e2e_in_flight >= 2 else (dt_serial, (out_h, st_h))
    args.e2e_serial = args.e2e_in_flight < 2
    call = calls[0]
    e2e_val = world * Be * K * args.steps / dt / 1e6
    e2e_serial_val = world * Be * K * args.steps / dt_serial / 1e6
    if not ((out_h2 == out_h).all() and (st_h2 == st_h).all()):
        raise SystemExit("bench.py: the two host-buffer loops disagree")
    same = bool((out_h == out_dev[:Be].cpu().numpy()).all() and (st_h == st[:Be]).all())
    if not same:
        raise SystemExit("bench.py: host-buffer path and device-resident path disagree")
    e2e_h2d, e2e_d2h = call.h2d_bytes, call.d2h_bytes
    call_keep = None
    del out_h, st_h, out_h2, st_h2, call

    # ---- BASELINE configs[3]: multi-cell uplink through the front end, all ranks (strong scaling over 64 cells) ----
    del y_pin, calls, call_keep
    multicell = None if args.no_multicell else run_multicell_ul(args, capi, rank, world, dist, local_rank)

    # ---- side measurement: early-exit regimes of config 3 (device-resident, same plan, same timing rules) ----
    regimes = None
    if world == 1 and not args.no_regimes:
        regimes = {}
        # (A = 8 keeps |LLR| <= 127; A = 128 / 256 are amplitudes of real demapper outputs: beyond the a-priori fast-path guard,
        # decoded through tracked fast passes with an a-posteriori range certificate)
        for name, sig, amp in (("clean", 0.5, 8), ("waterfall", 1.08, 8), ("waterfall_A128", 1.08, 128), ("waterfall_A256", 1.08, 256)):
            nd = 64
            ys, info = coded_inputs(K, nd, sig, 4242, A=amp)
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
            regimes[name] = {"value": B * K * args.steps / (rms * 1e-3) / 1e6, "unit": "Mbit/s", "sigma_over_A": sig, "A": amp,
                             "return_value_histogram": {str(i): h for i, h in enumerate(hist) if h},
                             "mean_iterations": float(sum(min(i, MAX_ITER) * h for i, h in enumerate(hist)) / max(sum(hist), 1)),
                             "crc_passing_blocks_equal_transmitted_bytes": good,
                             "distinct_blocks": nd}
            del y_r, out_r, st_r

    # ---- BASELINE configs[0] / [1]: one subframe through the drop-in call (latency), configs[4]: the 8-bit decoder ----
    subframes = llr8_side = None
    if world == 1 and not args.no_regimes:
        def guarded(fn, *a, **kw):                               # a side measurement never takes the headline line down
            try:
                return fn(*a, **kw)
            except Exception as ex:
                return {"error": str(ex)}
        subframes = {"ulsim_subframe_latency_ms": guarded(subframe_latency, capi, 7736, 14400, 4, 6),     # 25 PRB MCS16: 2 x K=3904
                     "dlsim_subframe_latency_ms": guarded(subframe_latency, capi, 75376, 90000, 6, 4),    # 100 PRB MCS28: 13 x K=5824
                     "dlsim_subframe_latency_ms_8bit": guarded(subframe_latency, capi, 75376, 90000, 6, 4, llr8=1)}
        llr8_side = {}
        B8 = 21312
        for K8 in (5824, 6144):
            r8, r16 = guarded(device_rate, capi, B8, K8, 1, args.steps), guarded(device_rate, capi, B8, K8, 0, args.steps)
            if "value" in r8 and "value" in r16:
                r8["ratio_to_16bit"] = r8["value"] / r16["value"]
                r8["value_16bit"] = r16["value"]
            llr8_side["K%d" % K8] = r8
        llr8_side["note"] = ("device-resident, %d blocks, noise regime, 6 iterations; ratio = 8-bit / 16-bit decoder at the same K and "
                             "batch (on the reference CPU the 8-bit decoder is ~1.4x the 16-bit one, SURVEY 6)" % B8)

    # ---- side measurement: the OPTIONAL sliding-window mode (north_star: reported separately, with its BLER delta) ----
    sw_side = None
    if world == 1 and not args.no_regimes:
        SW = capi.BATCH_SLIDING_WINDOW
        sw_side = {"bit_exact_with_reference": False,
                   "device_resident": guarded(device_rate, capi, B, K, 0, args.steps, flags=SW),
                   "dlsim_subframe_latency_ms": guarded(subframe_latency, capi, 75376, 90000, 6, 6, flags=SW),
                   "dlsim_subframe_latency_ms_bit_exact_mode": guarded(subframe_latency, capi, 75376, 90000, 6, 6),
                   "ulsim_subframe_latency_ms": guarded(subframe_latency, capi, 7736, 14400, 4, 6, flags=SW),
                   "bler_delta": "profiles/r2s_sw_bler_delta.txt (tools/sw_bler_delta.py; same received subframes in both modes): ulsim 25 PRB "
                                 "MCS16 -0.03 dB at BLER 10 % / -0.08 dB at 1 % with 6 iterations, -0.06 dB with 4; dlsim MCS28: BLER 0.02 at "
                                 "19 dB and 0.005 at 20 dB where the reference's algorithm still loses 58 / 40 % of the transport blocks",
                   "note": "OAI_BATCH_SLIDING_WINDOW: one warp decodes a block out of shared memory in one launch (td16_sw.cuh); never "
                           "part of `value` / `e2e`, which are the bit-exact mode"}

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
    # measured figures of the kernel from the committed ncu --set full capture of this bench command
    # (profiles/ncu_kmap16.json, written by tools/ncu_summary.py from the .ncu-rep): DRAM bytes, executed warp instructions
    # and pipe utilisations per launch, keyed by the block count of the captured launch and scaled linearly to this batch
    nk = {}
    try:
        nk = json.load(open(os.path.join(ROOT, "profiles", "ncu_kmap16.json")))
    except Exception:
        pass
    cap_blocks = float(nk.get("blocks_per_launch") or 0)
    traffic = (nk["dram_bytes_per_launch"] * B / cap_blocks) if (cap_blocks and K == nk.get("K")) else None
    ratio = (traffic / bytes_per_launch) if traffic else None
    actual_gbs = (traffic / (map_ms * 1e-3) / 1e9) if (traffic and n_map) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_map16", "avg_launch_ms": map_ms, "launches_timed": n_map,
                "share_of_step": prof_ms[1] / kernel_ms_total if kernel_ms_total else None,
                "kernel_ms": {"demux": prof_ms[0], "map": prof_ms[1], "x1": prof_ms[2], "x2": prof_ms[3]},
                "peak_source": peak_src,
                "traffic_source": nk.get("source"),
                "traffic_over_algorithmic": ratio,
                "actual_dram_gbs": actual_gbs, "actual_dram_frac": (actual_gbs / peak) if actual_gbs else None,
                "ncu_utilisation_pct": nk.get("utilisation_pct"),
                "timing_note": "per-launch CUDA events are recorded around every kernel INSIDE the timed region (they are what "
                               "`avg_launch_ms` and `kernel_ms` come from), so `value` includes their small overhead",
                "note": "algorithmic bytes per launch = blocks x (2(3K+12)+K/8+1)/12 (SURVEY 8d: y read once, bytes + status "
                        "written once, spread over the 12 MAP passes).  The decoder streams its per-block state through HBM "
                        "on every pass, so the bytes really moved per launch (`traffic`, ncu dram__bytes of the committed "
                        "capture) are `traffic_over_algorithmic` x that figure; `actual_dram_gbs` = traffic / launch time.  "
                        "Of the kernel's resources HBM is the one closest to its measured peak (`actual_dram_frac`), ahead of "
                        "the ALU pipe and the issue slots (`ncu_utilisation_pct`, `int_simd.issue_frac`): `bound` = hbm on "
                        "the traffic the design moves, not on the algorithmic bytes"}
    # integer-SIMD view from EXECUTED instructions (ncu smsp__inst_executed of the same capture): the SM issues at most
    # 4 warp instructions per clock; the measured peak of the kernel's own instruction mix is in profiles/int16_peak.json
    int_simd = {}
    if K == nk.get("K") and n_map and cap_blocks:
        sm_hz = ((clocks and clocks["sm_mhz"]) or 1965.0) * 1e6
        wi = nk["warp_instructions_per_launch"] / cap_blocks
        int_simd["executed_warp_instr_per_block_pass"] = wi
        int_simd["issue_frac"] = B * wi / (map_ms * 1e-3) / (148 * 4 * sm_hz)
        try:
            ipk = json.load(open(os.path.join(ROOT, "profiles", "int16_peak.json")))
            mix = ipk["warp_instr_per_clk_per_sm"]["2 VIADDMNMX+VIADD+IMAD (k_map16 mix)"]
            int_simd["issue_frac_of_measured_mix_peak"] = B * wi / (map_ms * 1e-3) / (148 * mix * sm_hz)
            int_simd["note"] = ("executed warp instructions per second / (148 SMs x 4 issue slots x SM clock); the measured peak of "
                                "the kernel's instruction mix is %.2f of 4 slots (tools/int16_peak.cu)" % mix)
        except Exception:
            pass

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
            # `config` is the workload only and is identical in both arms (bench.py --impl reference); how this arm measures
            # is in `method`
            "config": bench_config(B, Be, K),
            "method": {"timing": "CUDA events on the launching stream, barrier + synchronize both sides, max over ranks",
                       "sharding": "independent code blocks, one shard per rank, no data-path collective", "per_rank": per_rank},
            "roofline": roofline, "int_simd": int_simd, "cpu_baseline": cpu, "early_exit_regimes": regimes, "tx_mirror": tx_side,
            "multicell_ul": multicell, "subframe_latency": subframes, "llr8": llr8_side, "sliding_window_mode": sw_side,
            "e2e": {"value": e2e_val, "unit": "Mbit/s", "h2d_bytes_per_step": e2e_h2d,
                    "d2h_bytes_per_step": e2e_d2h, "ms_per_step": 1e3 * dt / args.steps,
                    "in_flight": 1 if args.e2e_serial else 2, "serial_value": e2e_serial_val,
                    "serial_ms_per_step": 1e3 * dt_serial / args.steps,
                    "api": "oai_turbo_submit_batch + oai_turbo_wait per step, page-locked host input and output buffers; `value`: two "
                           "batches in flight (submit of step i+1 before the wait of step i), `serial_value`: one call at a time; "
                           "both bound by the host->device copy of 36.9 KB of int16 LLRs per 6144 decoded bits "
                           "(profiles/r2d_link_ceiling.txt: 55 GB/s for one GPU = 9.2 Gbit/s)"},
            "gpu_launches": launches, "clocks": clocks}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def device_rate(capi, B, K, llr8, steps, flags=0):
    """device-resident decoder throughput (noise regime, 6 iterations) of B blocks of size K; CUDA events"""
    import torch
    row = 3 * K + 12 + (4 if llr8 else 0)
    g = torch.Generator(device="cuda")
    g.manual_seed(4321)
    y = torch.randint(-16, 17, (B, row), dtype=torch.int16, device="cuda", generator=g)
    out = torch.zeros((B, K // 8), dtype=torch.uint8, device="cuda")
    st = torch.zeros(B, dtype=torch.uint8, device="cuda")
    plan = capi.DevPlan(B, K, MAX_ITER, CRC_TYPE, llr8=llr8)
    if flags:
        plan.set_mode(flags)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.decode(y.data_ptr(), row, out.data_ptr(), K // 8, st.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.decode(y.data_ptr(), row, out.data_ptr(), K // 8, st.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if not (st == MAX_ITER + 1).all():
        raise RuntimeError("bench.py: device_rate: noise-regime blocks must report status 7")
    plan.close()
    return {"value": B * K * steps / (ms * 1e-3) / 1e6, "unit": "Mbit/s", "ms_per_step": ms / steps}


def subframe_latency(capi, tbs, G, Qm, max_it, llr8=0, reps=30, flags=0):
    """BASELINE configs[0]/[1]: wall-clock latency of ONE subframe's transport block through the host-buffer call
    (page-locked soft bits e in -> fused front end + decoder -> bytes out), the drop-in replacement of the per-code-block
    loops of ulsch_decoding.c:1222-1369 / dlsch_decoding.c:303-453.  Inputs come from the product's own TX chain
    (openair4g_b200/sim/txchain.py): `clean` (sigma/A = 0.25: every block leaves at iteration 2) and `full_iterations`
    (pure noise: max_it iterations).  Returns medians in ms."""
    import ctypes as C
    import numpy as np
    from openair4g_b200.sim import txchain as tx
    rc, seg = capi.lte_segmentation_params(tbs + 24)
    assert rc == 0
    Cn, F = seg["C"], seg["F"]
    Ks = [seg["Kminus"] if r < seg["Cminus"] else seg["Kplus"] for r in range(Cn)]
    rng = np.random.default_rng(tbs)
    a = rng.integers(0, 2, size=(1, tbs)).astype(np.uint8)
    b = np.concatenate([a, tx.crc24a(a)], axis=1)
    L = 24 if Cn > 1 else 0
    es, pos = [], 0
    for r, K in enumerate(Ks):
        cb = np.zeros((1, K), dtype=np.uint8)
        f = F if r == 0 else 0
        cb[:, f:K - L] = b[:, pos:pos + K - L - f]
        pos += K - L - f
        if Cn > 1:
            cb[:, K - 24:] = tx.crc24b(cb[:, :K - 24])
        bits, E = tx.rate_match(tx.turbo_encode(cb), K, f, G, Cn, Qm, 1, r, 0)
        es.append(bits[0].astype(np.int64))
    tx_bits = np.concatenate(es)
    out = {"code_blocks": Cn, "K": Ks[-1], "max_iterations": max_it, "decoder": "8-bit" if llr8 else "16-bit"}
    pin = capi.PinnedArray((G,), np.int16)
    obuf = capi.PinnedArray((Cn, 768), np.uint8)
    status = np.zeros(Cn, dtype=np.uint8)
    pool = capi.HarqPool(Cn, max(Ks))
    descs = (capi.CbDesc * Cn)()
    off = 0
    for r, K in enumerate(Ks):
        d = descs[r]
        d.in_ = pin.array.ctypes.data + 2 * off
        off += es[r].size
        d.decoded_bytes = obuf.array.ctypes.data + r * 768
        d.status = status.ctypes.data + r
        d.K, d.max_iterations, d.crc_type, d.F, d.decode_enable, d.dematch_enable = K, max_it, (0 if Cn == 1 else 1), (F if r == 0 else 0), 1, 1
        d.G, d.C, d.r, d.rvidx, d.clear, d.Qm, d.Nl, d.Mdlharq, d.Kmimo, d.Nsoft = G, Cn, r, 0, 1, Qm, 1, 8, 1, 1827072
        d.llr8 = llr8
        d.harq_pool = pool.handle
        d.harq_slot = r
    for name, sig in (("clean", 0.25), ("full_iterations", None)):
        if sig is None:
            pin.array[...] = rng.integers(-16, 17, size=G).astype(np.int16)
        else:
            pin.array[...] = np.clip(8 * (2 * tx_bits - 1) + np.rint(sig * 8 * rng.standard_normal(G)), -32768, 32767).astype(np.int16)
        ts = []
        for i in range(reps + 5):
            h = C.c_void_p()
            t0 = time.perf_counter()
            if capi.lib.oai_turbo_submit_batch(descs, Cn, flags | (capi.BATCH_DL_STOP_AFTER_FAILURE if sig is not None else 0), -1, C.byref(h)) \
                    or capi.lib.oai_turbo_wait(h):
                raise RuntimeError("bench.py: subframe batch failed: " + capi.last_error())
            if i >= 5:
                ts.append(time.perf_counter() - t0)
        want = 2 if sig is not None else max_it + 1
        if not (status == want).all():
            raise RuntimeError("bench.py: subframe latency (%s): unexpected status %s" % (name, status))
        out[name] = 1e3 * statistics.median(ts)
    pool.close()
    return out


def run_multicell_ul(args, capi, rank, world, dist, gpu):
    """BASELINE configs[3] through the product path: 64 cells, every cell-subframe one 100 PRB MCS16 allocation
    (ulsch_decoding.c:1222-1369 shape: C = 5 code blocks of K = 6144, G = 57600, Qm = 4, E = 11520 soft bits per block),
    cells assigned to GPUs with sharding.assign_by_cell, one device-resident HARQ pool per GPU, rate-matched soft bits e
    in page-locked host memory -> fused front end + decoder -> decoded bytes in page-locked host memory.  The 64 cells are
    spread over the GPUs and the number of subframes batched per step grows with N (weak scaling in time depth); timed on
    the host around submit + wait like `e2e`, max over ranks.
    Two feeds: the reference's int16 soft bits and the narrow int8 feed (oai_cb_desc_t.in_fmt = 1).
    The GPUs are identical but their host links are not (profiles/r2d_link_ceiling.txt), so for N > 1 the cells are first
    spread evenly, every rank's rate is measured on two untimed steps, and the cells are re-assigned in proportion."""
    import ctypes as C
    import numpy as np
    import torch
    from openair4g_b200 import sharding
    K, G, Cb, Qm, cells = K_BITS, 57600, 5, 4, 64
    E = G // Cb
    S = args.mc_subframes * world          # subframes batched per step grow with N: the 64 cells are fixed
    cell_of = [c for c in range(cells) for _ in range(S)]                                   # one entry per cell-subframe

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def prepare(owner, dt_np, fmt, ul=False):
        """two lanes (descriptor sets with their own output buffers and HARQ pool) over one page-locked input buffer, so
        that two batches can be in flight: consecutive subframe batches belong to different HARQ processes.
        ul: the input is the demodulator's soft bits of the whole allocation (lte_eNB_pusch_vars->llr, 1 HARQ-ACK bit
        multiplexed in), the uplink front end runs on the GPU (oai_ul_front_t) and the transport blocks come back assembled
        (oai_turbo_submit_tbs; no per-block bytes cross the link)"""
        mine = [i for i, r in enumerate(owner) if r == rank]
        n_ue = len(mine)
        n = n_ue * Cb
        st = {"n": n, "n_ue": n_ue, "isz": np.dtype(dt_np).itemsize, "lanes": []}
        st["pin"] = capi.PinnedArray((max(n_ue, 1), G), dt_np)
        g = torch.Generator()
        g.manual_seed(77 + rank)
        st["pin"].array[...] = torch.randint(-16, 17, (max(n_ue, 1), G), dtype=torch.int16, generator=g).numpy().astype(dt_np)
        base = st["pin"].array.ctypes.data
        for _ in range(2):
            ln = {"out": capi.PinnedArray((max(n, 1), K // 8), np.uint8), "status": np.zeros(max(n, 1), dtype=np.uint8),
                  "pool": capi.HarqPool(max(n, 1), K, gpu=gpu)}
            descs = (capi.CbDesc * max(n, 1))()
            ob, sb = ln["out"].array.ctypes.data, ln["status"].ctypes.data
            for u in range(n_ue):
                for r in range(Cb):
                    i = u * Cb + r
                    d = descs[i]
                    d.in_ = base + (u * G + r * E) * st["isz"]
                    d.decoded_bytes = ob + i * (K // 8)
                    d.status = sb + i
                    d.K, d.max_iterations, d.crc_type, d.F, d.decode_enable, d.dematch_enable = K, MAX_ITER, CRC_TYPE, 0, 1, 1
                    d.G, d.C, d.r, d.rvidx, d.clear, d.Qm, d.Nl, d.Mdlharq, d.Kmimo, d.Nsoft = G, Cb, r, 0, 1, Qm, 1, 8, 1, 1827072
                    d.tb_id = mine[u]
                    d.harq_pool = ln["pool"].handle
                    d.harq_slot = i
                    d.in_fmt = fmt
            ln["descs"] = descs
            if ul:
                tb_bytes = Cb * (K // 8) - 3 * Cb
                ln["b"] = capi.PinnedArray((max(n_ue, 1), tb_bytes), np.uint8)
                ln["ret"] = np.zeros(max(n_ue, 1), dtype=np.uint8)
                ln["oack"] = np.zeros((max(n_ue, 1), 2), dtype=np.uint8)
                ln["ufs"] = (capi.UlFront * max(n_ue, 1))()
                ln["tbs"] = (capi.TbDesc * max(n_ue, 1))()
                for u in range(n_ue):
                    f = ln["ufs"][u]
                    f.llr, f.llr_fmt, f.c_init = base + u * G * st["isz"], fmt, (0x1234 << 14) + (3 << 9) + (mine[u] % 504)
                    f.Qm, f.Ncp, f.O_ACK, f.O_RI, f.bundling, f.Nbundled, f.Cmux = Qm, 0, 1, 0, 0, 1, 12
                    f.Qprime_RI, f.Qprime_ACK, f.Qprime_CQI, f.Hprime = 0, ul_sizes["Qprime_ACK"], 0, ul_sizes["Hprime"]
                    f.o_ACK = ln["oack"].ctypes.data + 2 * u
                    t = ln["tbs"][u]
                    t.first_cb, t.C, t.uplink = u * Cb, Cb, 1
                    t.b, t.b_capacity, t.ret = ln["b"].array.ctypes.data + u * tb_bytes, tb_bytes, ln["ret"].ctypes.data + u
                    t.ul_front = C.pointer(f)
                    for r in range(Cb):
                        descs[u * Cb + r].in_ = None
                        descs[u * Cb + r].decoded_bytes = None
                        descs[u * Cb + r].in_fmt = 0
            st["lanes"].append(ln)
        st["ul"] = ul
        return st

    def submit(st, lane):
        if st["n"] == 0:
            return None
        h = C.c_void_p()
        ln = st["lanes"][lane]
        if st["ul"]:
            rc = capi.lib.oai_turbo_submit_tbs(ln["descs"], st["n"], ln["tbs"], st["n_ue"], 0, gpu, C.byref(h))
        else:
            rc = capi.lib.oai_turbo_submit_batch(ln["descs"], st["n"], 0, gpu, C.byref(h))
        if rc:
            raise SystemExit("bench.py: multicell_ul submit failed: " + capi.last_error())
        return h

    def wait(h):
        if h is not None and capi.lib.oai_turbo_wait(h):
            raise SystemExit("bench.py: multicell_ul wait failed: " + capi.last_error())

    def timed(st, steps):
        """two batches in flight: submit of step i+1 before the wait of step i"""
        barrier()
        t0 = time.perf_counter()
        pending = None
        for i in range(steps):
            h = submit(st, i & 1)
            wait(pending)
            pending = h
        wait(pending)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def close(st):
        for ln in st["lanes"]:
            ln["pool"].close()

    rc, ul_sizes = capi.ulsch_control_sizes(0, 1, 0, 1200, 12, 40, 40, 16, Cb * K, 100, Qm, 12)     # 1 HARQ-ACK bit, no RI / CQI
    if rc or ul_sizes["G"] != G:
        raise SystemExit("bench.py: unexpected uplink control sizes %s" % ul_sizes)
    res = {}
    for name, dt_np, fmt, ul in (("e_int16", np.int16, 0, False), ("e_int8", np.int8, 1, False),
                                 ("llr_int16_ul_front", np.int16, 0, True), ("llr_int8_ul_front", np.int8, 1, True)):
        owner = sharding.assign_by_cell(cell_of, world)
        st = prepare(owner, dt_np, fmt, ul)
        timed(st, 4)                                                # untimed: both batch objects are created and warm
        weights = None
        if world > 1:
            dt_cal = timed(st, 4)                                   # untimed calibration: this rank's blocks per second
            weights = sharding.measured_weights(st["n"] / dt_cal, dist)
            if max(weights) / min(weights) > 1.05:
                owner = sharding.assign_by_cell(cell_of, world, weights=weights)
                close(st)
                del st
                st = prepare(owner, dt_np, fmt, ul)
                timed(st, 4)
        dt = timed(st, args.steps)
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        n, n_ue, isz = st["n"], st["n_ue"], st["isz"]
        if n and not all((ln["status"][:n] == MAX_ITER + 1).all() for ln in st["lanes"]):
            raise SystemExit("bench.py: multicell_ul noise-regime blocks must report status 7")
        if ul and n and not all((ln["ret"][:n_ue] == MAX_ITER + 1).all() for ln in st["lanes"]):
            raise SystemExit("bench.py: multicell_ul transport blocks of pure noise must report ret 7")
        counts = [sum(1 for r in owner if r == k) * Cb for k in range(world)]
        res[name] = {"value": cells * S * Cb * K * args.steps / dt / 1e6, "unit": "Mbit/s", "ms_per_step": 1e3 * dt / args.steps,
                     "h2d_bytes_per_step_this_gpu": n_ue * G * isz,
                     "d2h_bytes_per_step_this_gpu": (n_ue * (Cb * (K // 8) - 3 * Cb + 8) + n) if ul else n * (K // 8 + 1),
                     "h2d_gbs_this_gpu": n_ue * G * isz * args.steps / dt / 1e9, "blocks_per_gpu": counts,
                     "rank_weights": None if weights is None else [round(w, 3) for w in weights], "in_flight": 2}
        close(st)
        del st
    res["config"] = {"workload": "BASELINE configs[3]: %d cells x %d subframes x (100 PRB MCS16 = 5 x K=6144, E=11520), noise regime, "
                                 "%d iterations; cells -> GPUs by sharding.assign_by_cell (weighted by each rank's measured rate "
                                 "when N > 1), one HARQ pool per GPU, page-locked soft bits in, bytes out.  e_*: rate-matched soft bits e per code "
                                 "block (what ulsch_decoding.c:1259 hands to lte_rate_matching_turbo_rx) -> decoded code blocks; "
                                 "llr_*_ul_front: the demodulator's soft bits of the allocation (1 HARQ-ACK bit multiplexed in) -> uplink "
                                 "front end on the GPU -> assembled transport blocks + return values (the whole data path of "
                                 "ulsch_decoding.c:600-1409 in one call)" % (cells, S, MAX_ITER),
                     "blocks_total": cells * S * Cb,
                     "scaling": "weak (64 cells fixed; subframes batched per step = %d x n_gpus)" % args.mc_subframes,
                     "api": "oai_turbo_submit_batch(dematch_enable, harq_pool, gpu) + oai_turbo_wait per step, two batches in flight"}
    return res


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
    emit(({"metric": "turbo_decoded_info_mbit_per_s_8bit_decoder", "value": B * K * args.steps / (ms * 1e-3) / 1e6,
                      "kernel_ms": {"demux": prof_ms[0], "map": prof_ms[1], "x1": prof_ms[2], "x2": prof_ms[3]},
                      "unit": "Mbit/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                      "dtype": "int8", "data": "synthetic",
                      "config": {"workload": "8-bit decoder, K=%d, max_iterations=6, noise regime" % K, "blocks": B},
                      "status_hist": torch.bincount(st_dev.long()).nonzero().flatten().tolist()}))


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the run, on the process's real stdout (see main)"""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    # stdout carries exactly one JSON line: whatever libraries print on the way (NCCL's version banner under torchrun, ...)
    # goes to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
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
    ap.add_argument("--e2e-in-flight", type=int, default=2, choices=[1, 2], help="host-buffer batches in flight in the e2e loop")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-multicell", action="store_true", help="skip the BASELINE configs[3] multi-cell uplink measurement")
    ap.add_argument("--mc-subframes", type=int, default=128, help="subframes per cell and step of the multi-cell measurement "
                    "(64 cells x 128 subframes x 5 blocks = 40960 code blocks per step in total)")
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
