#!/bin/bash
# quick GPU loop for the residual coder: zstd parity + whole-archive parity tests, then C2 through the CLI with wave tracing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -5
bash tools/run_c2_cli.sh 2>&1 | tail -14
