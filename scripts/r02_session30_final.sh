#!/bin/bash
# round 2, session 30: HEAD verified — full GPU suite, smoke(), and the at-scale self-check of the Dna5 N pass
# (3 Gbp with 5 % N: searches that skip the text's N + N pass against the N children walked, compared on the device)
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_s30_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02_s30_pytest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== Dna5 N pass at 3 Gbp"; GMB_VERBOSE=1 timeout 600 python tools/dna5_nfree_check.py 2>&1 | tee gpurun_out/r02_s30_dna5_nfree_check.log | tail -12
