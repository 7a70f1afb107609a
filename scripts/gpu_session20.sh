#!/bin/bash
# model-chosen block size: parity (full suite), defaults sweep (Dna4 + Dna5), config 5 (pan-genome) again
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s20_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/s20_pytest.log
echo "== sweep defaults"; timeout 900 python tools/sweep.py --reps 2 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5,4:-1:0.03125 > gpurun_out/s20_sweep.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s20_sweep.log
echo "== sweep defaults dna5"; timeout 900 python tools/sweep.py --n-frac 0.05 --reps 2 --configs 0:-1:256,1:-1:64,2:-1:8 > gpurun_out/s20_sweep_dna5.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/s20_sweep_dna5.log
echo "== pangenome"; timeout 1500 python tests/pangenome_bench.py > gpurun_out/s20_pangenome.log 2>&1; echo "rc=$?"; cat gpurun_out/s20_pangenome.log
