#!/bin/bash
# 2-GPU pass: multi-GPU tests (bit identity, streaming peer frames, sharded host frames) and the bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2_n2_pytest_multi.log 2>&1; echo "multi rc=$?"; tail -30 gpurun_out/r2_n2_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err; echo "bench n2 rc=$?"
tail -5 gpurun_out/r2_n2_bench.err
cut -c1-3000 gpurun_out/r2_n2_bench.json
