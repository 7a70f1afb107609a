#!/bin/bash
# round 2, session 4: the two-phase kernel for E >= 1 (block_kernel.cu) — kernel parity suite, A/B against the general
# kernel on the same tables, block sizes, ncu counters
mkdir -p gpurun_out
nvidia-smi -L
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
echo "== pytest kernels"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_s4_pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r02_s4_pytest.log
echo "== general kernel"; GMB_BLOCK_KERNEL=0 timeout 600 python tools/sweep.py --reps 3 --configs 1:-1:64,2:-1:8,3:-1:0.5 > gpurun_out/r02_s4_sweep_general.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/r02_s4_sweep_general.log
echo "== block kernel"; timeout 600 python tools/sweep.py --reps 3 --configs 1:-1:64,2:-1:8,3:-1:0.5 > gpurun_out/r02_s4_sweep_block.log 2>&1; echo "rc=$?"; cat gpurun_out/r02_s4_sweep_block.log
echo "== block sizes"; timeout 900 python tools/sweep.py --reps 2 --configs 1:-1:64:4,1:-1:64:6,1:-1:64:8,1:-1:64:10,1:-1:64:12,2:-1:8:2,2:-1:8:3,2:-1:8:4,2:-1:8:5,2:-1:8:6 > gpurun_out/r02_s4_sweep_blocks.log 2>&1; echo "rc=$?"; grep -v "fetches by" gpurun_out/r02_s4_sweep_blocks.log
echo "== ncu counters"
for cfg in 1:-1:64 2:-1:8; do
  E=${cfg%%:*}
  timeout 600 ncu --metrics $M --clock-control none -k regex:block_kernel -s 2 -c 1 --csv --log-file gpurun_out/r02_s4_ncu_e$E.csv python tools/sweep.py --configs $cfg --reps 2 > gpurun_out/r02_s4_ncu_e$E.log 2>&1; echo "E=$E rc=$?"
  grep -E "^\"0\"" gpurun_out/r02_s4_ncu_e$E.csv | cut -d, -f5,13,15 | cut -c1-160
done
