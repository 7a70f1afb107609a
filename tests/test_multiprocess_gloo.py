"""world_size-2 gloo test of the N>1 plumbing on CPU: rank 0 builds the index blob and broadcasts it,
every rank searches its own position shard, the shards tile the whole output.  The search itself is done
by the host-compiled state machine (tests/hostsim) because the product has no CPU path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gmtest as T


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, T.ROOT)
    sys.path.insert(0, os.path.join(T.ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import genmap_b200
    from genmap_b200 import parallel
    seqs = T.repeat_rich(77, 3, 4000)  # every rank knows the genome layout (as in bench.py), only rank 0 indexes it
    blob = None
    if rank == 0:
        blob = torch.from_numpy(genmap_b200.Index.build_blob(seqs))
    blob = parallel.broadcast_blob(blob, dist, torch.device("cpu"))
    ref = genmap_b200.Index.build_blob(seqs)
    assert blob.numpy().tobytes() == ref.tobytes()
    n = sum(len(s) for s in seqs)
    b, e = parallel.shard_range(n, rank, world)
    hs = T.HostSim(seqs)
    part = hs.map(20, 1, pos_begin=b, pos_end=e)
    whole = torch.from_numpy(part.astype(np.int32))
    dist.all_reduce(whole)  # shards are disjoint and zero elsewhere: the sum is the full vector
    if rank == 0:
        np.save(os.path.join(tmp, "whole.npy"), whole.numpy())
        np.save(os.path.join(tmp, "want.npy"), hs.map(20, 1))
    assert parallel.max_over_ranks(rank + 1, dist, torch.device("cpu")) == world
    assert parallel.sum_over_ranks(e - b, dist, torch.device("cpu")) == n
    dist.destroy_process_group()


def test_two_rank_broadcast_and_sharding(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert np.array_equal(np.load(tmp_path / "whole.npy"), np.load(tmp_path / "want.npy").astype(np.int32))


def test_shard_ranges_tile_exactly():
    from genmap_b200 import parallel
    for n in (0, 1, 7, 1000, 3_000_000_000):
        for w in (1, 2, 3, 4, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    bl = parallel.step_batches(100, 1100, 300, 5)
    assert bl == [(100, 400), (400, 700), (700, 1000), (100, 400), (400, 700)]
    assert parallel.step_batches(0, 375, 268, 4) == [(0, 268), (107, 375), (0, 268), (107, 375)]
    assert parallel.step_batches(0, 300, 300, 2) == [(0, 300), (0, 300)]
