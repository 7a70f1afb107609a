#!/bin/bash
# round 2, session 24: located table entries and the straight-line E = 0 kernel on Dna5 indices — full GPU suite, smoke(),
# the Dna5 sweep of session 13 again (general kernel, tables without located entries: 12.0 G / 1.90 G / 141 M positions/s)
#
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_s24_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02_s24_pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== Dna5 (5 % N), HEAD"; timeout 600 python tools/sweep.py --n-frac 0.05 --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8 2>&1 | tee gpurun_out/r02_s24_sweep_dna5.log | grep -v "fetches by"
echo "== Dna5 (5 % N), GMB_LOCATE=0 (tables of session 13)"; GMB_LOCATE=0 timeout 600 python tools/sweep.py --n-frac 0.05 --reps 3 --configs 0:-1:256,1:-1:64,2:-1:8 2>&1 | tee gpurun_out/r02_s24_sweep_dna5_locate0.log | grep -v "fetches by"
