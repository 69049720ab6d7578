python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 2> gpurun_out/bench_n2.err | cut -c1-600
tail -3 gpurun_out/bench_n2.err
