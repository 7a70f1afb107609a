#!/bin/bash
# round 2, session 3: the straight-line E = 0 kernel (exact_kernel.cu) — kernel parity suite, A/B against the
# general kernel on the same tables, ncu counters
mkdir -p gpurun_out
nvidia-smi -L
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
echo "== pytest kernels"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_s3_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02_s3_pytest.log
echo "== E=0 general kernel"; GMB_EXACT_KERNEL=0 timeout 600 python tools/sweep.py --reps 5 --configs 0:-1:256 > gpurun_out/r02_s3_sweep_general.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_s3_sweep_general.log
echo "== E=0 exact kernel"; timeout 600 python tools/sweep.py --reps 5 --configs 0:-1:256,0:14:256,0:15:256 > gpurun_out/r02_s3_sweep_exact.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_s3_sweep_exact.log
echo "== E=0 exact kernel, K=50 (text reads)"; timeout 600 python tools/sweep.py --kmer 50 --reps 3 --configs 0:-1:256 > gpurun_out/r02_s3_sweep_exact_k50.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_s3_sweep_exact_k50.log
echo "== ncu counters"; timeout 600 ncu --metrics $M --clock-control none -k regex:exact_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02_s3_ncu_e0.csv python tools/sweep.py --configs 0:-1:256 --reps 2 > gpurun_out/r02_s3_ncu_e0.log 2>&1; echo "rc=$?"; grep -E "^\"0\"" gpurun_out/r02_s3_ncu_e0.csv | cut -d, -f13,15
