#!/bin/bash
# Dna5 (genomes with N) on the device: parity suite, Dna4 regression check, first Dna5 timings
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/s10_pytest.log
echo "== sweep dna4"; timeout 600 python tools/sweep.py --configs 0:-1:256,1:-1:64,2:-1:8 > gpurun_out/s10_sweep_dna4.log 2>&1; echo "rc=$?"; cat gpurun_out/s10_sweep_dna4.log
echo "== sweep dna5"; timeout 600 python tools/sweep.py --n-frac 0.05 --configs 0:-1:256,0:-1:256:1,1:-1:64,1:-1:64:1,2:-1:8,2:-1:8:1 > gpurun_out/s10_sweep_dna5.log 2>&1; echo "rc=$?"; cat gpurun_out/s10_sweep_dna5.log
