#!/bin/bash
# First GPU call of a round: everything that has so far only run against the mocked device ABI (tests/mock) or the host build,
# in the order "most likely to be right" first, each under its own timeout, one log under gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_first_call.sh'
mkdir -p gpurun_out
L=gpurun_out/first_call.log
: > $L
run() { echo "=== $*" | tee -a $L; timeout 300 "$@" 2>&1 | tail -15 | tee -a $L; }
run python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_zstd.py -m gpu -q -x
run python -m pytest tests/test_gpu_zx_decoders.py -m gpu -q          # lz / zstd decoders, --verify, append (never run on hardware)
run python -m pytest tests/test_gpu_zy_fuzz.py -m gpu -q              # random collections x -a -c -f
run python -m pytest tests/test_gpu_zz_sharded.py -m gpu -q           # two ranks, one archive (both on cuda:0, gloo exchange)
run python bench.py --steps 3 --warmup 3
