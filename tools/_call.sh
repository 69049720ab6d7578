mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_zstd.py tests/test_gpu_zx_decoders.py tests/test_gpu_zz_sharded.py -m gpu -x -q 2>&1 | tail -3
sed -i 's/| tail -12//; s/grep -E "real|wave|phase"/grep -vE "phase (scan|assign|find_new|add_seg)|agcgpu.   frame"/' tools/run_c3_cli.sh
THREADS=$(nproc) timeout 600 bash tools/run_c3_cli.sh 2>&1 | tail -40
timeout 900 python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err; tail -3 gpurun_out/c20_bench.err; cut -c1-1500 gpurun_out/c20_bench.json
