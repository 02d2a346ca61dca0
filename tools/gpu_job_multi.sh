#!/bin/bash
# Final multi-GPU lines of round 2 on one 8-GPU box: bench at N = 2, 4, 8 and the 2-GPU tests
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2ab_pytest_multi_2gpu.log 2>&1; echo "multi rc=$?"; tail -2 gpurun_out/r2ab_pytest_multi_2gpu.log
for n in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2ab_bench_n$n.json 2> gpurun_out/r2ab_bench_n$n.err; echo "bench n$n rc=$?"
cut -c1-330 gpurun_out/r2ab_bench_n$n.json; echo
done
