mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c16_tests.log; cat gpurun_out/c16_tests.log
(timeout 300 python tools/lz_hpp_bench.py 16384 0.001 256; timeout 300 python tools/lz_hpp_bench.py 4096 0.01 64) 2>&1 | cut -c1-330
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; tail -3 gpurun_out/c16_bench.err; cat gpurun_out/c16_bench.json | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 | cut -c1-600
