#!/bin/bash
# mid-size whole-archive parity + timing on the GPU box: complex collection scaled up (3 contigs x 4 Mb, 10 samples with SNPs,
# indels, big deletions, reverse complements, trimmed ends, novel contigs; default k=31, 60 kb segments => ~200 groups, thousands
# of segments), ours vs the reference binary, archives compared byte for byte
python - <<'PY'
import sys; sys.path.insert(0,'tools')
import gen_data
files = gen_data.complex_collection('/dev/shm/mid', seed=11, n_samples=10, ctg_len=4000000, n_ctg=3)
open('/dev/shm/mid/list.txt','w').write("\n".join(files[1:])+"\n")
print(gen_data.total_bases(files), "bases")
PY
( time AGCGPU_TRACE=1 agc_b200/bin/agc-b200 create -o /dev/shm/mid/our.agc -i /dev/shm/mid/list.txt /dev/shm/mid/ref.fa ) 2>&1 | grep -E "real|wave|phase (scan|add_seg)" | tail -8
( time oracle/_ref/agc create -t 16 -o /dev/shm/mid/ref.agc -i /dev/shm/mid/list.txt /dev/shm/mid/ref.fa ) 2>&1 | grep -E "real"
cmp /dev/shm/mid/our.agc /dev/shm/mid/ref.agc && echo IDENTICAL; ls -la /dev/shm/mid/*.agc
