mkdir -p gpurun_out
( time timeout 1500 python bench.py > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err ) 2>&1 | grep real; tail -2 gpurun_out/r02w_bench_n1.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r02w_bench_n1.json')); print('default', round(d['value'],4), round(d['ms_per_step']), round(d['e2e']['value'],4), round(d['e2e']['ms_per_step']), round(d['cpu_baseline']['value'],4), d['roofline']['batch_frac']); print(json.dumps(d['other_workloads'])[:1500])"
