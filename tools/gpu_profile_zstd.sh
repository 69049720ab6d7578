#!/bin/bash
# run on the GPU box: cheap ncu passes over the residual-coder kernels (wide kernel on the 1.35 MB raw-group pack, narrow
# kernel on a batch of small frames); full --set captures of the wide kernel cost minutes of replay, see profiles/r01f_*
mkdir -p gpurun_out
M=gpu__time_duration.sum,launch__registers_per_thread,launch__block_size,launch__grid_size,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
ZS_ONLY=x0d0 ncu --metrics $M --clock-control none -k regex:k_zstd -s 1 -c 1 --csv --log-file gpurun_out/zstd_wide_x0d_metrics.csv python tools/zs_parts_prof.py > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_zstd_narrow -s 1 -c 1 -o gpurun_out/zstd_narrow_full python tools/zs_throughput.py 600 > gpurun_out/ncu_narrow.log 2>&1
ls -la gpurun_out | tail -5
