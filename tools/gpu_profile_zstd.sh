#!/bin/bash
mkdir -p gpurun_out
python tools/zs_prof.py 60000 2 > gpurun_out/zstd_timing.txt 2>&1; cat gpurun_out/zstd_timing.txt
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section MemoryWorkloadAnalysis --import-source on --clock-control none -k regex:k_zstd -c 1 -o gpurun_out/zstd_prof python tools/zs_prof.py 30000 1 > gpurun_out/ncu_zstd.log 2>&1; tail -3 gpurun_out/ncu_zstd.log
