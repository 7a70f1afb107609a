#!/bin/bash
# Dna5 32-byte blocks: full GPU suite + Dna5 sweep; BASELINE config 5 (pan-genome, -ep) on one GPU
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/s14_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/s14_pytest.log
echo "== sweep dna5"; timeout 600 python tools/sweep.py --n-frac 0.05 --reps 2 --configs 0:-1:256,1:-1:64,2:-1:8 > gpurun_out/s14_sweep_dna5.log 2>&1; echo "rc=$?"; cat gpurun_out/s14_sweep_dna5.log
echo "== pangenome"; timeout 1500 python tests/pangenome_bench.py --cpu > gpurun_out/s14_pangenome.log 2>&1; echo "rc=$?"; cat gpurun_out/s14_pangenome.log
