#!/bin/bash
# after restricting variant entries to the blocked instantiation: full parity suite, A/B of E=0 against the previous
# library on the same box, defaults sweep (Dna4, Dna5), config 5
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s25_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/s25_pytest.log
echo "== previous library"; GMB_LIB_PATH=$PWD/genmap_b200/lib/variants/libgenmap_b200_prev.so timeout 600 python tools/sweep.py --reps 3 --configs 0:-1:256 > gpurun_out/s25_sweep_prev.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s25_sweep_prev.log
echo "== current library"; timeout 900 python tools/sweep.py --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5 > gpurun_out/s25_sweep.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s25_sweep.log
echo "== dna5"; timeout 900 python tools/sweep.py --n-frac 0.05 --reps 2 --configs 0:-1:256,1:-1:64,2:-1:8 > gpurun_out/s25_sweep_dna5.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s25_sweep_dna5.log
echo "== pangenome"; timeout 1500 python tests/pangenome_bench.py > gpurun_out/s25_pangenome.log 2>&1; echo "rc=$?"; grep -v "^pan-genome\|^index" gpurun_out/s25_pangenome.log
