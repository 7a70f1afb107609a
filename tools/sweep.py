#!/usr/bin/env python
"""Tuning sweep on one GPU: build genome + index once, then time the map kernel for several
(E, jump depth, batch) settings.  Prints one line per setting (not a bench line)."""
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
ap.add_argument("--configs", default="0:-1:256,0:0:256,0:13:256,0:14:256,0:16:256,1:-1:64,1:0:64,2:-1:8,2:0:8")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--kmer", type=int, default=30)
ap.add_argument("--rep-frac", type=float, default=0.05, help="fraction of every chromosome covered by planted repeat copies")
ap.add_argument("--rep-mut", type=float, default=0.02, help="substitution rate of the planted copies")
ap.add_argument("--with-sa", action="store_true", help="keep the suffix array in the index (Dna5: enables the searches that skip the text's N + the N pass)")
ap.add_argument("--n-frac", type=float, default=0.0, help="fraction of every chromosome turned into runs of N (-> Dna5 index)")
args = ap.parse_args()

total = int(args.genome_mbp * 1e6)
t0 = time.time()
seqs = gm.synth_genome(total, args.nchr, args.seed, rep_frac=args.rep_frac, mut=args.rep_mut)
if args.n_frac > 0:  # assembly-gap model: one long run per chromosome (centromere) + a few short ones
    rng = np.random.default_rng(args.seed + 1)
    for s in seqs:
        big = int(len(s) * args.n_frac * 0.9)
        a = int(rng.integers(0, len(s) - big))
        s[a:a + big] = 4
        small = max(1, int(len(s) * args.n_frac * 0.1) // 20)
        for a in rng.integers(0, len(s) - small, 20):
            s[int(a):int(a) + small] = 4
ix = gm.Index.build(seqs, on_gpu=True, with_sa=args.with_sa)
print("genome+index %.1f s, build %s" % (time.time() - t0, ix.build_timings_ms), flush=True)
n = ix.n_text
out = torch.zeros(n, dtype=torch.int16, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for cfg in args.configs.split(","):
    parts = cfg.split(":")
    E, depth, batch = int(parts[0]), int(parts[1]), int(float(parts[2]) * (1 << 20))
    blk = int(parts[3]) if len(parts) > 3 else 0
    batch = min(batch, n // 2)
    ix.set_jump_depth(depth)
    p = gm.SearchParams(args.kmer, E, block_kmers=blk)
    t1 = time.time()
    st = ix.compute_mappability_device(p, out.data_ptr(), pos_begin=0, pos_end=1 << 16, stream=stream)  # builds tables
    setup = time.time() - t1
    ms, fetch, lut, npos = [], 0, 0, 0
    for r in range(args.reps):
        b = (r * batch) % (n - batch)
        st = ix.compute_mappability_device(p, out.data_ptr(), pos_begin=b, pos_end=b + batch, stream=stream)
        ms.append(st.kernel_ms); npos = st.positions
    b = 0
    st = ix.compute_mappability_device(p, out.data_ptr(), pos_begin=b, pos_end=b + batch, stream=stream, count_fetches=True)
    print("E=%d B=%d depth=%2d(%2d) batch=%d  %.2f ms  %.1f Mpos/s  fetch/pos=%.1f lut/pos=%.2f  algGB/s=%.0f  setup=%.2fs free=%.1fGB"
          % (E, blk, depth, st.jump_depth, batch, np.median(ms), npos / np.median(ms) / 1e3, st.rank_block_fetches / st.positions,
             st.jump_table_reads / st.positions, st.rank_block_fetches * ix.info.rank_block_bytes / st.positions * npos / np.median(ms) / 1e6,
             setup, torch.cuda.mem_get_info()[0] / 1e9), flush=True)
    tot = max(1, st.rank_block_fetches)
    print("    fetches by interval size 1|2|3-4|5-8|9-16|17-32|33-64|65+ : %s   thin paths/pos=%.2f (%.1f fetches each)  iterations/pos=%.1f  located/pos=%.2f text reads/pos=%.3f"
          % (" ".join("%.1f%%" % (100.0 * x / tot) for x in st.fetches_by_size), st.thin_paths / st.positions,
             st.fetches_by_size[0] / max(1, st.thin_paths), st.iterations / st.positions, st.located_entries / st.positions,
             st.text_reads / st.positions), flush=True)
