#!/usr/bin/env python
"""Evidence for bench.py's reference arm: the reference-format index written from OUR BWT + SA is byte-identical
to the one the unmodified reference's own `genmap index` writes, at a size well beyond the unit tests
(tests/test_seqan_index_writer.py stops at 160 kbp).  CPU only.

    python scripts/r02_writer_identity.py [mbp=40] [nchr=4] > profiles/r02/writer_identity.txt
"""
import filecmp, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gmtest as T
import genmap_b200 as gm

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
nchr = int(sys.argv[2]) if len(sys.argv) > 2 else 4
seqs = gm.synth_genome(int(mbp * 1e6), nchr, 45)
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
    fa = os.path.join(tmp, "genome.fa")
    T.write_fasta(fa, seqs)
    t0 = time.time()
    subprocess.run([T.REF_BIN, "index", "-F", fa, "-I", os.path.join(tmp, "ref_index")], check=True, stdout=subprocess.DEVNULL)
    t_ref = time.time() - t0
    t0 = time.time()
    hs = T.HostSim(seqs, with_sa=True)
    files = [("genome.fa", [("chr%d" % (i + 1), s) for i, s in enumerate(seqs)])]
    ours = T.write_seqan_index(os.path.join(tmp, "our_index"), files, hs.bwt(False), hs.bwt(True), hs.sa())
    t_our = time.time() - t0
    names = sorted(os.listdir(os.path.join(tmp, "ref_index")))
    same, diff, err = filecmp.cmpfiles(os.path.join(tmp, "ref_index"), ours, names, shallow=False)
    print("genome: %g Mbp, %d sequences, frozen generator seed 45" % (mbp, nchr))
    print("genmap_ref index: %.1f s; host SA-IS builder + reference-format writer: %.1f s" % (t_ref, t_our))
    for n in names:
        print("  %-22s %12d bytes  %s" % (n, os.path.getsize(os.path.join(ours, n)), "identical" if n in same else "DIFFERENT"))
    print("files written by us only:", sorted(set(os.listdir(ours)) - set(names)))
    print("RESULT:", "byte-identical (%d files)" % len(same) if not diff and not err and sorted(os.listdir(ours)) == names else "MISMATCH %s %s" % (diff, err))
