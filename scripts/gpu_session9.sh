#!/bin/bash
# 2-GPU session: torchrun bench (NCCL index broadcast + position sharding), multi-GPU CLI test
mkdir -p gpurun_out
nvidia-smi -L
echo "== bench N=2"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.log; echo "rc=$?"; tail -12 gpurun_out/bench_n2.log; cat gpurun_out/bench_n2.json
echo "== reference arm under torchrun"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.log; echo "rc=$?"; tail -3 gpurun_out/bench_ref_n2.log; cat gpurun_out/bench_ref_n2.json
echo "== multi-GPU CLI test"; timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -q -k "multi_gpu or prefix" > gpurun_out/pytest_n2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_n2.log
echo "== bench N=1 (pipelined e2e)"; timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "rc=$?"; cat gpurun_out/bench_n1.json
