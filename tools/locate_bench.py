#!/usr/bin/env python
"""Timing of the locate path (gmb_map_locations: what `genmap map -d` runs) on one GPU: occurrences/s and
positions/s for a few (K, E) on a repeat-rich synthetic genome.  Not a bench line."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genmap_b200 as gm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genome-mbp", type=float, default=250)
ap.add_argument("--nchr", type=int, default=5)
ap.add_argument("--seed", type=int, default=44)
ap.add_argument("--positions", type=int, default=16 << 20)
args = ap.parse_args()

seqs = gm.synth_genome(int(args.genome_mbp * 1e6), args.nchr, args.seed)
ix = gm.Index.build(seqs, with_sa=True, on_gpu=True)
print("index built:", ix.build_timings_ms, flush=True)
for K, E in ((30, 0), (30, 1), (30, 2), (50, 2)):
    p = gm.SearchParams(K, E)
    ix.compute_locations(p, pos_begin=0, pos_end=1 << 16)  # builds the jump tables
    t0 = time.time()
    off, loc = ix.compute_locations(p, pos_begin=1 << 20, pos_end=(1 << 20) + args.positions)
    dt = time.time() - t0
    print("K=%d E=%d: %d positions, %d occurrences, %.3f s end to end (host arrays included): %.1f Mpos/s, %.1f Mocc/s"
          % (K, E, args.positions, len(loc), dt, args.positions / dt / 1e6, len(loc) / dt / 1e6), flush=True)
