#!/bin/bash
# round 2, session 29: after has_n_in (the "N in the common infix" test of the two-phase kernel for infixes of 32+ characters:
# device and host mirror had skipped different blocks — same counts, different counters): the Dna5 / located / counter tests
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dna5 or located or fetch_counter" > gpurun_out/r02_s29_pytest_dna5.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02_s29_pytest_dna5.log
