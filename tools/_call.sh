mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/r02x_bench_n1.json 2> gpurun_out/r02x_bench_n1.err ) 2>&1 | grep real; tail -2 gpurun_out/r02x_bench_n1.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r02x_bench_n1.json')); print('default', round(d['value'],4), round(d['ms_per_step']), round(d['e2e']['value'],4), round(d['e2e']['ms_per_step']), round(d['cpu_baseline']['value'],4)); print(json.dumps(d['residual_coder'])[:900]); print({k:(v.get('value'),v.get('e2e'),v.get('error')) for k,v in d['other_workloads'].items()})"
