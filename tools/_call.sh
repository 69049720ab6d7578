mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
THREADS=$(nproc) timeout 600 bash tools/run_c3_cli.sh | tail -18
grep -E "phase" gpurun_out/c3_trace.log | awk '{a[$3" "$4" "$5]+=$(NF-1)} END {for (k in a) print a[k], k}' | sort -rn | head -20
