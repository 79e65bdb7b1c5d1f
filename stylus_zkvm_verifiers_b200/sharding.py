"""Proof-range sharding for the one-process-per-GPU form.

Proofs are independent, so a batch of n proofs is cut into contiguous ranges [g*n/G, (g+1)*n/G), one per rank; there is
no collective on the data path.  The only communication is the gather of one status byte per proof on the rank that
asked (and, in bench.py, the max-over-ranks of the timings).  The same cut is used inside one process by
`for_each_device` in csrc/zkv.cu when a handle is created with several device indices.
"""
import numpy as np


def shard_range(n, rank, world):
    """[begin, end) of the proofs rank `rank` of `world` verifies (same arithmetic as for_each_device in zkv.cu)."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard request n=%r rank=%r world=%r" % (n, rank, world))
    return n * rank // world, n * (rank + 1) // world


def shard_ranges(n, world):
    return [shard_range(n, g, world) for g in range(world)]


def gather_status(local_status, n, dist=None, dst=0):
    """Concatenate per-rank status arrays in rank order on rank `dst` (None elsewhere).  `dist` is torch.distributed
    (already initialised) or None for a single process."""
    local = np.ascontiguousarray(local_status, dtype=np.uint8)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.size == n
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    b, e = shard_range(n, rank, world)
    assert local.size == e - b, "rank %d holds %d statuses for a shard of %d" % (rank, local.size, e - b)
    cap = max(e_ - b_ for b_, e_ in shard_ranges(n, world))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    pad = torch.full((cap,), 255, dtype=torch.uint8, device=dev)
    pad[: local.size] = torch.from_numpy(local).to(dev)
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, parts, dst=dst)
    if rank != dst:
        return None
    out = np.empty(n, dtype=np.uint8)
    for g, (b_, e_) in enumerate(shard_ranges(n, world)):
        out[b_:e_] = parts[g][: e_ - b_].cpu().numpy()
    return out


def verify_sharded(verify_fn, n, dist=None, dst=0):
    """Run `verify_fn(begin, end) -> status[end-begin]` on this rank's range and gather the bytes on `dst`."""
    if dist is None or not dist.is_initialized():
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(), dist.get_world_size()
    b, e = shard_range(n, rank, world)
    return gather_status(verify_fn(b, e), n, dist, dst)
