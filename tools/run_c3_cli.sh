#!/bin/bash
# BASELINE configs[2] at full size on the GPU box: 64 synthetic 5 Mb bacterial genomes, adaptive mode (-a), k=29;
# ours vs the reference binary, archives compared byte for byte.  N_SAMPLES / REF_LEN shrink it for a quick run.
python - <<PY
import sys; sys.path.insert(0,'tools')
import gen_data
files = gen_data.bacterial_adaptive('/dev/shm/c3', seed=2, n_samples=${N_SAMPLES:-63}, ref_len=${REF_LEN:-5000000})
open('/dev/shm/c3/list.txt','w').write("\n".join(files[1:])+"\n")
print(gen_data.total_bases(files), "bases")
PY
mkdir -p gpurun_out
( time AGCGPU_TRACE=1 AGCGPU_TRACE_LZ=${TRACE_LZ:-} agc_b200/bin/agc-b200 create -a -k 29 -o /dev/shm/c3/our.agc -i /dev/shm/c3/list.txt /dev/shm/c3/ref.fa ) > gpurun_out/c3_trace.log 2>&1
grep -E "real|total|LZ encode|phase (close|residual coder collect)" gpurun_out/c3_trace.log | tail -14
( time oracle/_ref/agc create -a -k 29 -t ${THREADS:-16} -o /dev/shm/c3/ref.agc -i /dev/shm/c3/list.txt /dev/shm/c3/ref.fa ) 2>&1 | grep -E "real"
cmp /dev/shm/c3/our.agc /dev/shm/c3/ref.agc && echo IDENTICAL; ls -la /dev/shm/c3/*.agc
