#!/usr/bin/env python
"""TEST INFRASTRUCTURE (lives under tests/ because its --cpu leg runs the reference binary of oracle/_ref).
BASELINE config 5 on one GPU: 10 FASTA files x 300 Mbp synthetic pan-genome, K=50 E=2 --exclude-pseudo.

File g = the base genome (3 x 100 Mbp, seed 46) with g % iid substitutions, all ten indexed
together (3 Gbp + the full suffix array).  Times the search kernel on batches of positions of several files and,
with --cpu, the unmodified reference binary on a scale model (10 x --cpu-mbp) of the same construction, checking
the counts of the scale model against it.  Prints plain lines (not a bench.py line).
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # gmtest (reference-binary runner) for the --cpu leg
import genmap_b200 as gm  # noqa: E402


def pangenome(file_mbp, n_files, nchr=3, seed=46):
    return gm.synth_pangenome(file_mbp * 1e6, n_files, nchr, seed)


ap = argparse.ArgumentParser()
ap.add_argument("--file-mbp", type=float, default=300)
ap.add_argument("--files", type=int, default=10)
ap.add_argument("-K", type=int, default=50)
ap.add_argument("-E", type=int, default=2)
ap.add_argument("--batch-mpos", type=float, default=4)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--blocks", default="0", help="comma-separated k-mers per block to try (0 = the planner's choice)")
ap.add_argument("--cpu", action="store_true", help="also run the reference binary on a scale model and compare")
ap.add_argument("--cpu-mbp", type=float, default=3.0, help="per-file size of the CPU scale model")
args = ap.parse_args()

import torch  # noqa: E402

t0 = time.time()
seqs, stf = pangenome(args.file_mbp, args.files)
limits = np.zeros(len(seqs) + 1, dtype=np.uint64)
limits[1:] = np.cumsum([len(s) for s in seqs])
print("pan-genome: %d files x %.0f Mbp = %.2f Gbp in %d sequences (%.1f s)" % (args.files, args.file_mbp, limits[-1] / 1e9, len(seqs), time.time() - t0), flush=True)
t0 = time.time()
ix = gm.Index.build(seqs, with_sa=True, on_gpu=True, seq_to_file=stf)
print("index with suffix array built on the GPU in %.1f s %s, blob %.2f GB" % (time.time() - t0, ix.build_timings_ms, ix.info.blob_bytes / 1e9), flush=True)
batch = int(args.batch_mpos * (1 << 20))
per_file = len(seqs) // args.files
out = torch.zeros(int(limits[per_file]), dtype=torch.int16, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for ep, blk in [(e, int(b)) for b in args.blocks.split(",") for e in (True, False)]:
    p = gm.SearchParams(args.K, args.E, True, ep, 16, block_kmers=blk)
    print("-- k-mers per block: %s" % (blk or "planner"), flush=True)
    rates = []
    for fi in (0, args.files // 2, args.files - 1):
        s0 = fi * per_file
        tb, tl = int(limits[s0]), int(limits[s0 + per_file] - limits[s0])
        cum = np.ascontiguousarray(limits[s0:s0 + per_file + 1] - limits[s0])
        ix.compute_mappability_device(p, out.data_ptr(), text_begin=tb, text_len=tl, chrom_cum_lengths=cum, pos_begin=0, pos_end=1 << 14, stream=stream)
        ms = []
        for r in range(args.reps):
            b = (r * batch) % max(1, tl - batch)
            st = ix.compute_mappability_device(p, out.data_ptr(), text_begin=tb, text_len=tl, chrom_cum_lengths=cum,
                                               pos_begin=b, pos_end=min(tl, b + batch), stream=stream)
            ms.append(st.kernel_ms)
        rate = st.positions / np.median(ms) / 1e3
        rates.append(rate)
        vals = out[:min(tl, batch)].cpu().numpy().view(np.uint16)
        print("K=%d E=%d %s file %d: %.2f ms per %d positions = %.2f Mpos/s; mean value %.2f, max %d"
              % (args.K, args.E, "--exclude-pseudo" if ep else "(plain counts)", fi, np.median(ms), st.positions, rate, vals[:st.positions].mean(), vals.max()), flush=True)
    print("  -> %.2f Mpos/s per GPU (mean over files)" % np.mean(rates), flush=True)
ix.close()

if args.cpu:
    import gmtest as T
    seqs2, stf2 = pangenome(args.cpu_mbp, args.files)
    tmp = tempfile.mkdtemp(prefix="gmb_pan_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fdir = os.path.join(tmp, "fa"); os.mkdir(fdir)
    per = len(seqs2) // args.files
    for g in range(args.files):
        T.write_fasta(os.path.join(fdir, "g%02d.fa" % g), seqs2[g * per:(g + 1) * per], names=["g%02d_chr%d" % (g, i + 1) for i in range(per)])
    t0 = time.time()
    subprocess.run([T.REF_BIN, "index", "-FD", fdir, "-I", os.path.join(tmp, "index")], check=True, stdout=subprocess.DEVNULL)
    t_index = time.time() - t0
    odir = os.path.join(tmp, "out"); os.mkdir(odir)
    cores = os.cpu_count()
    res = subprocess.run([T.REF_BIN, "map", "-I", os.path.join(tmp, "index"), "-O", odir, "-K", str(args.K), "-E", str(args.E), "-ep", "-r", "-fl",
                          "-T", str(cores), "-v"], check=True, stdout=subprocess.PIPE, text=True)
    secs = [float(l.split()[3]) for l in res.stdout.replace("\r", "\n").split("\n") if l.startswith("Mappability computed in")][0]
    n_all = sum(len(s) for s in seqs2)
    print("reference binary, scale model %d x %.1f Mbp, -ep, -T %d: index %.1f s, map %.2f s = %.3f Mpos/s"
          % (args.files, args.cpu_mbp, cores, t_index, secs, n_all / secs / 1e6), flush=True)
    lim2 = np.zeros(len(seqs2) + 1, dtype=np.uint64); lim2[1:] = np.cumsum([len(s) for s in seqs2])
    ix2 = gm.Index.build(seqs2, with_sa=True, on_gpu=True, seq_to_file=stf2)
    ok, t_gpu = True, 0.0
    for g in range(args.files):
        s0 = g * per
        tb, tl = int(lim2[s0]), int(lim2[s0 + per] - lim2[s0])
        got, st = ix2.compute_mappability(gm.SearchParams(args.K, args.E, True, True, 16), text_begin=tb, text_len=tl,
                                          chrom_cum_lengths=np.ascontiguousarray(lim2[s0:s0 + per + 1] - lim2[s0]), return_stats=True)
        t_gpu += st.kernel_ms
        ref = np.fromfile(os.path.join(odir, "g%02d.genmap.freq16" % g), dtype=np.uint16)
        ok = ok and np.array_equal(got, ref)
    print("same scale model on the GPU: %.1f ms of kernel time = %.1f Mpos/s; counts %s the reference's"
          % (t_gpu, n_all / t_gpu / 1e3, "EQUAL" if ok else "DIFFER FROM"), flush=True)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
