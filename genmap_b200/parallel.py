"""Multi-GPU plumbing for the map path: one process per GPU, `torch.distributed` only for the one-off
index broadcast and for timing reductions.  The path itself has no data-path collective: every position's
count depends only on the read-only index (the reference's independent OpenMP iterations,
src/algo.hpp:434-439), so the index blob is replicated and text positions are range-partitioned."""
import numpy as np


def shard_range(n_positions, rank, world):
    """Contiguous range of file-local positions owned by `rank` (ranges tile [0, n) exactly)."""
    return n_positions * rank // world, n_positions * (rank + 1) // world


def step_batches(shard_begin, shard_end, batch, count):
    """`count` consecutive batches of `batch` positions inside the shard, wrapping around at its end.  When the
    shard holds fewer than two whole batches, the batch that would run over the end is moved back to end at it, so
    consecutive steps still search different ranges."""
    span = shard_end - shard_begin
    batch = min(batch, span)
    res, pos, moved_back = [], 0, False
    for _ in range(count):
        if pos + batch > span:
            moved_back = not moved_back and 0 < span - batch < batch
            pos = span - batch if moved_back else 0
        res.append((shard_begin + pos, shard_begin + pos + batch))
        pos += batch
    return res


def broadcast_blob(blob, dist, device, src=0):
    """Replicate the index blob from rank `src` to every rank.  `blob`: a uint8 torch tensor on `device`
    on the source rank (may be a zero-copy view of library-owned memory), None elsewhere.  Returns the
    tensor holding the blob on this rank (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    import torch
    nbytes = torch.zeros(1, dtype=torch.int64, device=device)
    if dist.get_rank() == src:
        nbytes[0] = blob.numel()
    dist.broadcast(nbytes, src)
    n = int(nbytes.item())
    if dist.get_rank() != src:
        blob = torch.empty(n, dtype=torch.uint8, device=device)
    dist.broadcast(blob, src)
    return blob


def max_over_ranks(value, dist, device):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, dist, device):
    import torch
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
