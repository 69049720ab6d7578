mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_zstd.py -m gpu -x -q 2>&1 | tail -2
(timeout 300 python tools/lz_hpp_bench.py 96 0.01 96; timeout 300 python tools/lz_hpp_bench.py 4096 0.01 64; timeout 300 python tools/lz_hpp_bench.py 16384 0.001 256) 2>&1 | cut -c1-330
sed -i 's/| tail -12//; s/grep -E "real|wave|phase"/grep -vE "phase (scan|assign|find_new|add_seg)|agcgpu.   frame"/' tools/run_c3_cli.sh
THREADS=$(nproc) timeout 600 bash tools/run_c3_cli.sh 2>&1 | tail -19
bash tools/run_c2_cli.sh 2>&1 | grep -E "real|IDENT|zstd wave|wide:" 
