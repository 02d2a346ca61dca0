#!/bin/bash
# 2-GPU pass on the trimmed kernels: multi-GPU tests and the bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2y_pytest_multi_2gpu.log 2>&1; echo "multi rc=$?"; tail -8 gpurun_out/r2y_pytest_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2y_bench_n2.json 2> gpurun_out/r2y_bench_n2.err; echo "bench n2 rc=$?"
tail -5 gpurun_out/r2y_bench_n2.err
cut -c1-1200 gpurun_out/r2y_bench_n2.json
