mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0,'tools')
import gen_data
files = gen_data.human_chromosome('/dev/shm/c4s', seed=3, n_samples=4, ctg_len=50_000_000, n_repeats=40)
open('/dev/shm/c4s/list.txt','w').write("\n".join(files[1:])+"\n")
PY
( time AGCGPU_TRACE=1 agc_b200/bin/agc-b200 create -k 31 -o /dev/shm/c4s/our.agc -i /dev/shm/c4s/list.txt /dev/shm/c4s/ref.fa ) > gpurun_out/c4s_trace.log 2>&1
grep -vE "zstd wave|frame " gpurun_out/c4s_trace.log | tail -40
( time AGCGPU_TRACE=1 agc_b200/bin/agc-b200 create -k 31 -o /dev/shm/c4s/our.agc -i /dev/shm/c4s/list.txt /dev/shm/c4s/ref.fa ) 2>&1 | grep real
