#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rk45.py -q -m gpu -k dense > gpurun_out/r2d_pytest_dense.log 2>&1; echo "dense rc=$?"; tail -3 gpurun_out/r2d_pytest_dense.log
LP_REMAP_TMA=1 timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_main.py -q -m gpu -x > gpurun_out/r2d_pytest_frame_tma.log 2>&1; echo "frame tests with TMA remap rc=$?"
tail -15 gpurun_out/r2d_pytest_frame_tma.log
LP_REMAP_TMA=0 timeout 300 python tools/remap_perf.py > gpurun_out/r2d_remap_perf.log 2>&1
LP_REMAP_TMA=1 timeout 300 python tools/remap_perf.py >> gpurun_out/r2d_remap_perf.log 2>&1
cat gpurun_out/r2d_remap_perf.log
