#!/usr/bin/env python
"""profiles/r02/sass_summary.txt: instruction counts and memory / warp-level mnemonics of the hot kernel instantiations
in the built library (cuobjdump -sass; works without a GPU).    python scripts/sass_summary.py > profiles/r02/sass_summary.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "genmap_b200", "lib", "libgenmap_b200.so")], capture_output=True, text=True).stdout
want = {"exact_kernelILi1ELb0EtLi4E": "exact_kernel<KW=1, COUNT=0, uint16, SIGMA=4> (E = 0, K <= 32)",
        "exact_kernelILi1ELb0EtLi5E": "exact_kernel<KW=1, COUNT=0, uint16, SIGMA=5> (E = 0, K <= 32, Dna5 index)",
        "block_kernelILi2ELb0EtLb0ELi3ELi4E": "block_kernel<KW=2, COUNT=0, uint16, EP=0, MINB=3, SIGMA=4> (E = 1, 2 at K = 30)",
        "block_kernelILi1ELb0EtLb0ELi3ELi5E": "block_kernel<KW=1, COUNT=0, uint16, EP=0, MINB=3, SIGMA=5> (E = 1, 2 at K = 30, Dna5 index with the suffix array)",
        "map_kernelILi2ELb0EtLb0ELb1ELi4ELb0ELi4E": "map_kernel<KW=2, COUNT=0, uint16, EP=0, BLK=1, SIGMA=4, LOC=0, MINB=4> (general kernel, E >= 3)",
        "k_locate_singletons": "k_locate_singletons (text pass of the table builder)"}
print("SASS summary of the hot instantiations in genmap_b200/lib/libgenmap_b200.so (cuobjdump -sass, sm_100a; scripts/sass_summary.py).")
print("Per kernel: instruction count and the memory / warp-level mnemonics.  LDG.E.ENL2.LTC64B.256 = one 32-byte rank block per")
print("request, LDG.E.LTC64B.128 = one 16-byte table entry.  No UTMALDG / UTC*MMA is expected: per-thread 16- and 32-byte random")
print("reads, no dense contraction.\n")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0]
    for k, label in want.items():
        if k not in name:
            continue
        ops = collections.Counter()
        n = 0
        for l in f.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                n += 1
                ops[m.group(1)] += 1
        print("%s\n    %s\n    instructions: %d" % (label, name[:120], n))
        for o, c in sorted(ops.items()):
            if re.match(r"(LDG|STG|LDS|STS|ATOM|RED|SHFL|POPC|VOTE|BAR|LDL|STL|MATCH|REDUX|UTMA|UTC)", o):
                print("    %-34s %d" % (o, c))
        print()
