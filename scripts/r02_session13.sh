#!/bin/bash
# round 2, session 13: halving pieces of the host pipeline (e2e), Dna5 numbers at HEAD, CLI index timing with the new FASTA reader
mkdir -p gpurun_out
nvidia-smi -L
echo "== kernel + range tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "range or sharded or golden or edge or plan_cache" 2>&1 | tail -3
echo "== bench (e2e with halving pieces)"; timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/r02_s13_bench_n1.json 2> gpurun_out/r02_s13_bench_n1.log; echo "rc=$?"; python - <<'PY'
import json
j = json.load(open("gpurun_out/r02_s13_bench_n1.json"))
print("E=0 value %.3g e2e %.3g" % (j["value"], j["e2e"]["value"]))
for k, v in j["extra"].items(): print(k, "value %.4g e2e %.4g" % (v["value"], v["e2e"]["value"]))
PY
echo "== Dna5 (5 % N)"; timeout 900 python tools/sweep.py --n-frac 0.05 --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8 2>&1 | tee gpurun_out/r02_s13_sweep_dna5.log | grep -v "fetches by"
echo "== genmap index at 3 Gbp"; W=/dev/shm/gmb_idx3g; rm -rf $W; mkdir -p $W
python - <<PY
import sys, time
sys.path.insert(0, ".")
import genmap_b200 as gm
from genmap_b200 import synth
seqs = gm.synth_genome(3_000_000_000, 24, 45)
t = time.time(); synth.write_fasta("$W/genome.fa", seqs); print("FASTA written in %.1f s" % (time.time() - t), flush=True)
PY
a=$(date +%s%N); genmap_b200/bin/genmap index -F $W/genome.fa -I $W/index -v -xn; b=$(date +%s%N); echo "genmap index -xn: $(( (b - a) / 1000000 )) ms wall"
rm -rf $W
