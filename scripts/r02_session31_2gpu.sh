#!/bin/bash
# round 2, session 31 (2 GPUs): the two tests of the suite that need a second GPU, at HEAD — index replica over NVLink and
# `genmap map --gpus 2` on a Dna4 and on a Dna5 genome with gaps (every GPU builds the N pass and applies it to its slice)
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -q -k "second_gpu or multi_gpu" > gpurun_out/r02_s31_pytest_2gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02_s31_pytest_2gpu.log
