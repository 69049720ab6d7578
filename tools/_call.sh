mkdir -p gpurun_out
ZS_DUMP=1 AGCGPU_TRACE=1 timeout 600 python tools/zs_c3_prof.py > gpurun_out/c3_zsprof.log 2>&1
grep -vE "^\[agcgpu\]   frame|phase" gpurun_out/c3_zsprof.log | cut -c1-400 | tail -40
