mkdir -p gpurun_out
timeout 1500 python bench.py --workload c4s --steps 2 --warmup 1 --no-extra > gpurun_out/r02g_bench_c4s_n1.json 2> gpurun_out/r02g_bench_c4s_n1.err; tail -3 gpurun_out/r02g_bench_c4s_n1.err; cut -c1-3500 gpurun_out/r02g_bench_c4s_n1.json
