mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0,'tools')
import gen_data
files = gen_data.human_chromosome('/dev/shm/c4', seed=3, n_samples=8, ctg_len=250_000_000, n_repeats=200)
open('/dev/shm/c4/list.txt','w').write("\n".join(files[1:])+"\n")
PY
( time AGCGPU_TRACE=1 agc_b200/bin/agc-b200 create -k 31 -o /dev/shm/c4/our.agc -i /dev/shm/c4/list.txt /dev/shm/c4/ref.fa ) > gpurun_out/c4_trace.log 2>&1
grep -vE "zstd wave|frame |cost split" gpurun_out/c4_trace.log | tail -60
