#!/usr/bin/env python
"""Generates tests/golden/ref_synth.npz: outputs of the UNMODIFIED reference binary
(oracle/_ref/genmap_ref, built by oracle/build_ref.sh from /root/reference) on small seeded
genomes.  Run in the build container (the reference binary is needed); the .npz is committed so the
tests that consume it run anywhere.  Inputs are stored next to the outputs, so the fixtures do not
depend on the NumPy RNG stream.

    python tests/golden/make_fixtures.py
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gmtest as T  # noqa: E402

OUT = os.path.join(T.GOLDEN, "ref_synth.npz")


def main():
    assert T.have_reference(), "build oracle/_ref/genmap_ref first (oracle/build_ref.sh)"
    store = {}
    cases = []

    def add(name, files, K, E, flags=(), bits=16):
        """files: list of list of code arrays (one inner list per FASTA file)"""
        with tempfile.TemporaryDirectory() as tmp:
            if len(files) == 1:
                src = os.path.join(tmp, "genome.fa")
                T.write_fasta(src, files[0], names=["s%d" % i for i in range(len(files[0]))], width=60)
            else:
                src = os.path.join(tmp, "fastas")
                os.mkdir(src)
                for fi, seqs in enumerate(files):
                    T.write_fasta(os.path.join(src, "g%02d.fa" % fi), seqs,
                                  names=["f%ds%d" % (fi, i) for i in range(len(seqs))], width=60)
            outs = T.run_reference(src, K, E, flags=flags, value_bits=bits)
        for fi, seqs in enumerate(files):
            for si, s in enumerate(seqs):
                store["%s/in/%d/%d" % (name, fi, si)] = s
        for fi, key in enumerate(sorted(outs)):
            store["%s/out/%d" % (name, fi)] = outs[key]
        cases.append("%s|%d|%d|%s|%d|%d" % (name, K, E, " ".join(flags), bits, len(files)))
        print(name, K, E, flags, bits, [int(v.sum()) for v in outs.values()])

    g4 = T.repeat_rich(7, 3, 6000)
    for K, E in [(30, 0), (30, 1), (30, 2), (21, 3), (16, 4), (50, 2), (12, 2), (9, 1), (8, 0), (33, 1), (64, 2)]:
        add("dna4_K%d_E%d" % (K, E), [g4], K, E)
    add("dna4_K30_E2_nc", [g4], 30, 2, flags=("-nc",))
    add("dna4_K10_E2_fs", [g4], 10, 2, bits=8)  # saturates at 255
    g5 = T.repeat_rich(11, 3, 5000, with_n=True)
    for K, E in [(20, 0), (20, 1), (20, 2), (14, 3)]:
        add("dna5_K%d_E%d" % (K, E), [g5], K, E)
    base = T.repeat_rich(13, 2, 4000)
    rng = np.random.default_rng(5)
    multi = []
    for g in range(3):
        seqs = []
        for s in base:
            s = s.copy()
            m = rng.random(len(s)) < 0.02 * g
            s[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
            seqs.append(s)
        multi.append(seqs)
    add("multi_K25_E2", multi, 25, 2)
    add("multi_K25_E2_ep", multi, 25, 2, flags=("-ep",))
    add("multi_K25_E1_ep_nc", multi, 25, 1, flags=("-ep", "-nc"))
    pal = [np.tile(np.array([0, 1, 2, 3], dtype=np.uint8), 2000)]  # (ACGT)x2000: palindromes + saturation
    add("acgt_K8_E0", [pal], 8, 0)
    add("acgt_K8_E0_fs", [pal], 8, 0, bits=8)
    short = [np.array([0, 1, 2], np.uint8), T.repeat_rich(3, 1, 300)[0], np.array([3, 3], np.uint8),
             T.repeat_rich(4, 1, 200)[0], np.array([2], np.uint8)]
    add("short_seqs_K12_E1", [short], 12, 1)  # sequences shorter than K between longer ones
    store["cases"] = np.array(cases)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
