#!/bin/bash
# round 2, session 26: Dna5 indices with the suffix array — searches that skip the text's N (substituted keys, two-phase
# kernel) + the N pass: the Dna5 tests of the GPU suite, then the 3 Gbp genome with 5 % N against session 24's numbers
# (E = 1: 1.93 G, E = 2: 140 M positions/s with the N children walked)
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -k dna5"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dna5 or located" > gpurun_out/r02_s26_pytest_dna5.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r02_s26_pytest_dna5.log
echo "== Dna5 (5 % N) with the suffix array"; GMB_VERBOSE=1 timeout 600 python tools/sweep.py --with-sa --n-frac 0.05 --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8,3:-1:0.5 2>&1 | tee gpurun_out/r02_s26_sweep_dna5_npass.log
echo "== the same, N children walked"; GMB_DNA5_NFREE=0 timeout 600 python tools/sweep.py --with-sa --n-frac 0.05 --reps 3 --configs 1:-1:64,2:-1:8,3:-1:0.5 2>&1 | tee gpurun_out/r02_s26_sweep_dna5_walked.log | grep -v "fetches by"
