#!/bin/bash
# Per-kernel ncu table of every live kernel.  Run on the GPU box through gpurun from the repo root:
#   gpurun --timeout 900 -- bash tools/ncu_per_kernel.sh
# then, here:  python tools/ncu_per_kernel.py gpurun_out/r2_all_kernels.csv > profiles/r2_per_kernel.md
set -e
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
M=$M,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
M=$M,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct
python tools/all_kernels.py > gpurun_out/r2_all_kernels_plain.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_all_kernels.csv python tools/all_kernels.py > gpurun_out/r2_all_kernels_ncu.log 2>&1
