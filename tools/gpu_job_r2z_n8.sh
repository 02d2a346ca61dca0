#!/bin/bash
# bench at N = 8 with 32 x 1 warp tiles for 8-bit peer bands
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2z_bench_n8.json 2> gpurun_out/r2z_bench_n8.err; echo "bench n8 rc=$?"
tail -2 gpurun_out/r2z_bench_n8.err
cut -c1-400 gpurun_out/r2z_bench_n8.json
