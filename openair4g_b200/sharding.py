"""Multi-GPU sharding of independent code blocks.

Every code block is independent (read-only tables apart, a HARQ buffer belongs to one
(cell, UE, harq, r)), so the data path needs no collective: blocks are partitioned across
ranks, each rank decodes its shard on its own GPU, and only per-rank counters / results are
gathered after the timed region.  Functions here are backend-agnostic (`nccl` on GPUs,
`gloo` in the CPU tests).
"""
from typing import List, Sequence, Tuple


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n_items for `rank` (sizes differ by at most one)."""
    assert 0 <= rank < world
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def assign_by_cell(cells: Sequence[int], world: int, weights: Sequence[float] = None) -> List[int]:
    """Rank of every block when blocks are tagged with a cell id: all blocks of a cell go to
    one GPU so that a UE's HARQ buffers stay resident on one device; cells are placed on the
    rank that finishes them earliest (largest cells first).

    weights: relative throughput of the ranks (default: all equal).  The GPUs of one box are
    identical but their host links are not (profiles/r2d_link_ceiling.txt: with all eight GPUs
    fed from host memory, four of them get 23 GB/s each and four 35 GB/s), so a host-fed job
    balances by measured rate: a rank's load counts as blocks / weight."""
    w = [1.0] * world if weights is None else [float(x) for x in weights]
    assert len(w) == world and all(x > 0 for x in w)
    load = [0.0] * world
    count = {}
    for c in cells:
        count[c] = count.get(c, 0) + 1
    owner = {}
    for c, n in sorted(count.items(), key=lambda kv: (-kv[1], kv[0])):
        r = min(range(world), key=lambda i: ((load[i] + n) / w[i], i))
        owner[c] = r
        load[r] += n
    return [owner[c] for c in cells]


def measured_weights(local_rate: float, dist=None) -> List[float]:
    """All-gathers one positive number per rank (e.g. blocks per second of a calibration step) and
    returns the list, normalised to mean 1, on every rank."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [1.0]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(local_rate)], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    v = [float(o.item()) for o in out]
    m = sum(v) / len(v)
    return [x / m for x in v]


def gather_results(local_status, local_bits, dist=None):
    """Final result gather, outside the timed region: returns on every rank the list of
    per-rank (blocks, info bits, status histogram) tuples."""
    import torch
    hist = torch.bincount(torch.as_tensor(local_status, dtype=torch.int64).flatten().cpu(), minlength=256)
    rec = torch.cat([torch.tensor([int(hist.sum()), int(local_bits)], dtype=torch.int64), hist])
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [rec]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    rec = rec.to(dev)
    out = [torch.zeros_like(rec) for _ in range(dist.get_world_size())]
    dist.all_gather(out, rec)
    return [o.cpu() for o in out]
