mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_pipeline.py tests/test_gpu_zz_sharded.py -m gpu -x -q 2>&1 | tail -3
for w in c4 c3; do timeout 1500 python bench.py --workload $w --steps 2 --warmup 1 --no-extra > gpurun_out/r02v_bench_${w}_n1.json 2> gpurun_out/r02v_bench_${w}_n1.err; tail -2 gpurun_out/r02v_bench_${w}_n1.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r02v_bench_${w}_n1.json')); print('$w', round(d['value'],4), round(d['ms_per_step']), round(d['e2e']['value'],4), round(d['e2e']['ms_per_step']), round(d['cpu_baseline']['value'],4), 'zstd', round(d['residual_coder']['ms_per_step']), round(d['residual_coder']['host_wait_ms_per_step']), 'lz', round(d['roofline']['kernel_ms_per_step'],1), d['bit_exact'])"; done
