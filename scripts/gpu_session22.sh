#!/bin/bash
# scaling check: torchrun bench on all GPUs of the box (NCCL index broadcast, positions range-partitioned)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi -L
echo "== bench N=$N"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/s22_bench_n$N.json 2> gpurun_out/s22_bench_n$N.log; echo "rc=$?"; grep -i "broadcast\|error" gpurun_out/s22_bench_n$N.log | head -12; cat gpurun_out/s22_bench_n$N.json | cut -c1-900
echo "== CLI --gpus $N (index replicas over NVLink, device runs merged)"; timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -q -k multi_gpu 2>&1 | tail -2
