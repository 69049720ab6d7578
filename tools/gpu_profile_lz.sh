#!/bin/bash
# run on the GPU box: timing first (no profiler), then the ncu launch list and one full capture of the LZ kernel
mkdir -p gpurun_out
python tools/profile_lz.py 4096 60031 0.001 64 5 > gpurun_out/lz_timing_hpp.txt 2>&1
python tools/profile_lz.py 1000 30000 0.01 1 5 > gpurun_out/lz_timing_viral.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lz_launches.csv python tools/profile_lz.py 4096 60031 0.001 64 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lz_packed -c 1 -o gpurun_out/lz_packed_full python tools/profile_lz.py 4096 60031 0.001 64 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
