#!/bin/bash
# C2 through the CLI with tracing (diagnostics)
python - <<'PY'
import sys; sys.path.insert(0,'tools')
import gen_data
files,_ = gen_data.viral('/dev/shm/c2', n_samples=1000, ref_len=30000, p=0.01, seed=1)
open('/dev/shm/c2/list.txt','w').write("\n".join(files[1:])+"\n")
PY
time AGCGPU_TRACE=1 agc_b200/bin/agc-b200 create -k 25 -o /dev/shm/c2/our.agc -i /dev/shm/c2/list.txt /dev/shm/c2/ref.fa
time oracle/_ref/agc create -k 25 -t 32 -o /dev/shm/c2/ref.agc -i /dev/shm/c2/list.txt /dev/shm/c2/ref.fa
cmp /dev/shm/c2/our.agc /dev/shm/c2/ref.agc && echo IDENTICAL; ls -la /dev/shm/c2/*.agc
