#!/bin/bash
# helper for gpurun: run (a subset of) the GPU test-suite and keep the log
mkdir -p gpurun_out
python -m pytest ${@:-tests} -m gpu -x -q 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
