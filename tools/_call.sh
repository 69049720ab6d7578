mkdir -p gpurun_out
timeout 300 python tools/lzc_debug.py 1 9000 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/c8_tests.log; cat gpurun_out/c8_tests.log
(timeout 300 python tools/lz_hpp_bench.py 4096 0.001 64; timeout 300 python tools/lz_hpp_bench.py 16384 0.001 256; timeout 300 python tools/lz_hpp_bench.py 4096 0.01 64; AGCGPU_LZC_NOSTAGE=1 timeout 300 python tools/lz_hpp_bench.py 16384 0.001 256) > gpurun_out/c8_lz.log 2>&1; cat gpurun_out/c8_lz.log
sed -i 's/| tail -12//; s/grep -E "real|wave|phase"/grep -vE "phase (scan|assign|find_new|add_seg)|agcgpu.   frame"/' tools/run_c3_cli.sh
THREADS=$(nproc) timeout 600 bash tools/run_c3_cli.sh > gpurun_out/c8_c3.log 2>&1
tail -22 gpurun_out/c8_c3.log
