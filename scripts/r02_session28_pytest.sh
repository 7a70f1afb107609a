#!/bin/bash
# round 2, session 28: the full GPU suite at HEAD (session 27 stopped at a wrong assertion of a new test: its parity part had passed)
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_s28_pytest.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r02_s28_pytest.log
