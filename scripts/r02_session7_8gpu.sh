#!/bin/bash
# round 2, session 7 (8 GPUs): multi-GPU tests, bench at N=8 with e2e for E=0/1/2 and the D2H-only ceiling,
# BASELINE config 5 at N=8, one full command-line run at 3 Gbp on 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | head -8
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== multi-GPU tests"; timeout 600 python -m pytest tests -m gpu -q -k "replica or gpus or multi" > gpurun_out/r02_s7_pytest_multigpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02_s7_pytest_multigpu.log
echo "== bench N=8"; timeout 900 $RUN --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_s7_bench_n8.json 2> gpurun_out/r02_s7_bench_n8.log; echo "rc=$?"; grep "\[bench\]" gpurun_out/r02_s7_bench_n8.log | tail -12; cut -c1-400 gpurun_out/r02_s7_bench_n8.json
echo "== bench config 5 N=8"; timeout 900 $RUN --master-port 29512 bench.py --config pangenome --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_s7_bench_pangenome_n8.json 2> gpurun_out/r02_s7_bench_pangenome_n8.log; echo "rc=$?"; grep "\[bench\]" gpurun_out/r02_s7_bench_pangenome_n8.log | tail -12; cut -c1-400 gpurun_out/r02_s7_bench_pangenome_n8.json
echo "== command line, 3 Gbp, 8 GPUs"; timeout 900 bash scripts/r02_cli_3gbp.sh 8 > gpurun_out/r02_s7_cli_3gbp_n8.log 2>&1; echo "rc=$?"; grep -v "^-rw\|^total\|^drwx" gpurun_out/r02_s7_cli_3gbp_n8.log | tail -60
