#!/bin/bash
# round 2, session 32: DRAM traffic per launch of the three bench kernels at HEAD (-> profiles/ncu_traffic.json, what the
# bench line reports as roofline.traffic) — the kernels' template signatures changed with the Dna5 work — and the same
# counters for the two-phase kernel on the Dna5 genome (searches that skip the text's N)
mkdir -p gpurun_out
nvidia-smi -L
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
timeout 300 ncu --metrics $M --clock-control none -k regex:exact_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_s32_ncu_e0.csv python tools/sweep.py --configs 0:-1:256 --reps 2 > gpurun_out/r02_s32_ncu_e0.log 2>&1; echo "E=0 rc=$?"
timeout 300 ncu --metrics $M --clock-control none -k regex:block_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_s32_ncu_e1.csv python tools/sweep.py --configs 1:-1:64 --reps 2 > gpurun_out/r02_s32_ncu_e1.log 2>&1; echo "E=1 rc=$?"
timeout 300 ncu --metrics $M --clock-control none -k regex:block_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_s32_ncu_e2.csv python tools/sweep.py --configs 2:-1:8 --reps 2 > gpurun_out/r02_s32_ncu_e2.log 2>&1; echo "E=2 rc=$?"
timeout 300 ncu --metrics $M --clock-control none -k regex:block_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_s32_ncu_dna5_e1.csv python tools/sweep.py --with-sa --n-frac 0.05 --configs 1:-1:64 --reps 2 > gpurun_out/r02_s32_ncu_dna5_e1.log 2>&1; echo "Dna5 E=1 rc=$?"
grep -h "^E=" gpurun_out/r02_s32_ncu_*.log
