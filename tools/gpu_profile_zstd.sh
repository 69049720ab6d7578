#!/bin/bash
mkdir -p gpurun_out
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:k_zstd -c 1 -f -o gpurun_out/zstd_prof_raw python tools/zs_prof.py 300000 1 raw > gpurun_out/ncu_zstd.log 2>&1; tail -2 gpurun_out/ncu_zstd.log
