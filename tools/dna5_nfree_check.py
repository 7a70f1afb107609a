#!/usr/bin/env python
"""At-scale self-check of the Dna5 N pass on one GPU: the 3 Gbp bench genome with 5 % of every
chromosome turned into assembly gaps, index with the suffix array; the same positions mapped twice — searches that
skip the text's N + the N pass, and (GMB_DNA5_NFREE=0) the N children walked — and compared on the device.
Prints one line per (E, range); not a bench line."""
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
ap.add_argument("--n-frac", type=float, default=0.05)
ap.add_argument("--kmer", type=int, default=30)
ap.add_argument("--configs", default="1:0:3000,2:0:400,2:1400:1800,3:100:110", help="E:first Mbp:last Mbp of the text, comma separated")
args = ap.parse_args()

total = int(args.genome_mbp * 1e6)
t0 = time.time()
seqs = gm.synth_genome(total, args.nchr, args.seed)
rng = np.random.default_rng(args.seed + 1)  # the gap model of tools/sweep.py --n-frac
edges = []
off = 0
for s in seqs:
    big = int(len(s) * args.n_frac * 0.9)
    a = int(rng.integers(0, len(s) - big))
    s[a:a + big] = 4
    edges += [off + a, off + a + big]
    small = max(1, int(len(s) * args.n_frac * 0.1) // 20)
    for a in rng.integers(0, len(s) - small, 20):
        s[int(a):int(a) + small] = 4
        edges += [off + int(a), off + int(a) + small]
    off += len(s)
edges = np.array(sorted(edges), dtype=np.int64)
ix = gm.Index.build(seqs, on_gpu=True, with_sa=True)
print("genome+index %.1f s, alphabet %d, %d gap edges" % (time.time() - t0, ix.info.alphabet_size, len(edges)), flush=True)
n = ix.n_text
stream = torch.cuda.current_stream().cuda_stream
a_out = torch.zeros(n, dtype=torch.int16, device="cuda")
b_out = torch.zeros(n, dtype=torch.int16, device="cuda")
bad = 0
for cfg in args.configs.split(","):
    E, lo, hi = cfg.split(":")
    E, lo, hi = int(E), min(n, int(float(lo) * 1e6)), min(n, int(float(hi) * 1e6))
    p = gm.SearchParams(args.kmer, E)
    res = []
    for out, env in ((a_out, None), (b_out, "0")):
        if env is None:
            os.environ.pop("GMB_DNA5_NFREE", None)
        else:
            os.environ["GMB_DNA5_NFREE"] = env
        out.zero_()
        ix.compute_mappability_device(p, out.data_ptr(), pos_begin=lo, pos_end=min(hi, lo + (1 << 16)), stream=stream)  # tables, N pass
        st = ix.compute_mappability_device(p, out.data_ptr(), pos_begin=lo, pos_end=hi, stream=stream)
        res.append(st)
    os.environ.pop("GMB_DNA5_NFREE", None)
    diff = int((a_out[lo:hi] != b_out[lo:hi]).sum().item())
    inside = edges[(edges >= lo) & (edges < hi)]
    near = 0
    for e in inside:  # values next to the gap edges: where the N pass writes
        near += int((a_out[max(lo, e - args.kmer):min(hi, e + args.kmer)] != 0).sum().item())
    bad += diff
    print("E=%d positions [%d, %d): %d differences; %d gap edges inside, %d non-zero counts within K of them; "
          "N pass %.2f ms (%d launches, %.0f M positions/s) vs walked %.2f ms (%.0f M positions/s)"
          % (E, lo, hi, diff, len(inside), near, res[0].kernel_ms, res[0].kernel_launches, res[0].positions / res[0].kernel_ms / 1e3,
             res[1].kernel_ms, res[1].positions / res[1].kernel_ms / 1e3), flush=True)
print("TOTAL differences: %d" % bad)
sys.exit(1 if bad else 0)
