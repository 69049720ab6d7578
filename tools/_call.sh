mkdir -p gpurun_out
nvidia-smi -L | wc -l
for w in c4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 2 --warmup 1 --workload $w > gpurun_out/r02u_bench_${w}_n4.json 2> gpurun_out/r02u_bench_${w}_n4.err
tail -2 gpurun_out/r02u_bench_${w}_n4.err | cut -c1-200; python -c "
import json; d=json.loads(open('gpurun_out/r02u_bench_${w}_n4.json').read().strip().splitlines()[-1]); print('$w n4', round(d['value'],4), round(d['ms_per_step']), round(d['e2e']['value'],4), round(d['e2e']['ms_per_step']), round(d['cpu_baseline']['value'],4), 'zstd', round(d['residual_coder']['ms_per_step']), 'lz', round(d['roofline']['kernel_ms_per_step'],1), d['comm']['collectives'], d['comm']['bytes_gathered'], d['bit_exact'])"
done
