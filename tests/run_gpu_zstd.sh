#!/bin/bash
# quick GPU loop for the residual coder: zstd parity + whole-archive parity tests, then C2 through the CLI with wave tracing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -3
python tools/zs_prof.py 60000 1; python tools/zs_prof.py 300000 1 raw
bash tools/run_c2_cli.sh 2>&1 | grep -E "zstd wave|IDENT|real"
