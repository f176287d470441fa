"""Host->device link ceiling of the box: every rank copies from its own page-locked buffer to its own GPU at the same time
(the pattern of bench.py's host-fed legs, one process per GPU).  Run under torchrun with 1, 2, 4, 8 ranks:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/link_ceiling_probe.py

Rank 0 prints one line per transfer size: per-GPU GB/s (min / median / max over ranks) and the aggregate, for H2D alone,
D2H alone and both directions at once.  Timed with CUDA events per rank between two barriers; aggregate = total bytes /
max over ranks of the elapsed time."""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world == 1:
        return [x]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


res = []
for mb in (64, 512):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    h_in.fill_(3)
    s2 = torch.cuda.Stream()
    for mode in ("h2d", "d2h", "both"):
        reps = 8
        for timed in (False, True):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                if mode in ("h2d", "both"):
                    d_in.copy_(h_in, non_blocking=True)
                if mode == "d2h":
                    h_out.copy_(d_out, non_blocking=True)
                if mode == "both":
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        per = gather(reps * n / (ms * 1e-3) / 1e9)           # GB/s per direction on this rank
        tmax = max(gather(ms))
        if rank == 0:
            per_sorted = sorted(per)
            r = {"ranks": world, "mode": mode, "MB_per_copy": mb, "per_gpu_gbs_min": per_sorted[0],
                 "per_gpu_gbs_median": per_sorted[len(per) // 2], "per_gpu_gbs_max": per_sorted[-1],
                 "aggregate_gbs_per_direction": world * reps * n / (tmax * 1e-3) / 1e9, "per_gpu_gbs": [round(x, 1) for x in per]}
            res.append(r)
            print(json.dumps(r), flush=True)
    del h_in, h_out, d_in, d_out
if world > 1:
    dist.destroy_process_group()
