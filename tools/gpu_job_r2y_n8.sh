#!/bin/bash
# bench at N = 8 and N = 4 on one 8-GPU box (trimmed kernels)
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2y_bench_n$n.json 2> gpurun_out/r2y_bench_n$n.err; echo "bench n$n rc=$?"
tail -2 gpurun_out/r2y_bench_n$n.err
cut -c1-400 gpurun_out/r2y_bench_n$n.json
done
