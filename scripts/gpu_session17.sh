#!/bin/bash
# 2-GPU session: index replica over NVLink, multi-GPU CLI (device runs merged across slices), torchrun bench N=2
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest 2-GPU tests"; timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or replica or sharded" > gpurun_out/s17_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/s17_pytest.log
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -4
echo "== bench N=2"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/s17_bench_n2.json 2> gpurun_out/s17_bench_n2.log; echo "rc=$?"; tail -3 gpurun_out/s17_bench_n2.log; cat gpurun_out/s17_bench_n2.json
