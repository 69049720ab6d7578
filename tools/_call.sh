mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lzc_parse -s 2 -c 1 -o gpurun_out/c12_parse_1pct python tools/lz_hpp_bench.py 4096 0.01 64 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lzc_parse -s 2 -c 1 -o gpurun_out/c12_parse_01pct python tools/lz_hpp_bench.py 16384 0.001 256 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lzc_stitch -s 2 -c 1 -o gpurun_out/c12_stitch_01pct python tools/lz_hpp_bench.py 16384 0.001 256 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
