#!/usr/bin/env python
"""Tuning sweep: relative lengths of the search-scheme parts (GMB_PART_WEIGHTS) x block size, one GPU.
Any split keeps the scheme exhaustive and non-redundant, so the counts must not change; only the size of the
search tree (rank-block fetches per position) does.  Prints one line per setting."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genmap_b200 as gm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genome-mbp", type=float, default=3000)
ap.add_argument("--nchr", type=int, default=24)
ap.add_argument("--seed", type=int, default=45)
ap.add_argument("-K", type=int, default=30)
ap.add_argument("-E", type=int, default=2)
ap.add_argument("--batch-mpos", type=float, default=8)
ap.add_argument("--blocks", default="6")
ap.add_argument("--weights", default="1,1,1,1;4,4,6,6;4,5,6,6;5,5,6,6;5,5,7,7;4,4,5,5;3,3,4,4;6,6,7,7;5,6,7,7;6,5,7,7;5,5,6,8;5,5,8,6")
args = ap.parse_args()

seqs = gm.synth_genome(int(args.genome_mbp * 1e6), args.nchr, args.seed)
ix = gm.Index.build(seqs, on_gpu=True)
n = ix.n_text
out = torch.zeros(n, dtype=torch.int16, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
batch = min(int(args.batch_mpos * (1 << 20)), n // 2)
ref = None
for blk in [int(x) for x in args.blocks.split(",")]:
    for w in args.weights.split(";"):
        if w == "default":
            os.environ.pop("GMB_PART_WEIGHTS", None)
        else:
            os.environ["GMB_PART_WEIGHTS"] = w
        p = gm.SearchParams(args.K, args.E, block_kmers=blk)
        ix.compute_mappability_device(p, out.data_ptr(), pos_begin=0, pos_end=1 << 16, stream=stream)
        ms = []
        for r in range(2):
            b = n // 3 + r * batch
            st = ix.compute_mappability_device(p, out.data_ptr(), pos_begin=b, pos_end=b + batch, stream=stream)
            ms.append(st.kernel_ms)
        got = out[n // 3:n // 3 + batch].clone()
        if ref is None:
            ref = got
        same = bool(torch.equal(ref, got))
        st = ix.compute_mappability_device(p, out.data_ptr(), pos_begin=n // 3, pos_end=n // 3 + batch, stream=stream, count_fetches=True)
        print("K=%d E=%d B=%d weights=%-12s %.2f ms  %.1f Mpos/s  fetch/pos=%.1f  depth=%d  counts %s"
              % (args.K, args.E, blk, w, min(ms), st.positions / min(ms) / 1e3, st.rank_block_fetches / st.positions, st.jump_depth,
                 "same" if same else "DIFFER"), flush=True)
