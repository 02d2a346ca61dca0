#!/bin/bash
# Step-size controllers on the squared error norm (inv_tenth_root instead of sqrt / pow): RK45 and Kerr tests + timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_rk45.py tests/test_gpu_kerr.py tests/test_gpu_main.py -q -m gpu -x > gpurun_out/r2w_pytest_rk45_kerr.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2w_pytest_rk45_kerr.log
timeout 600 python tools/rk45_perf.py > gpurun_out/r2w_rk45_perf.log 2>&1; cat gpurun_out/r2w_rk45_perf.log
timeout 600 python tools/kerr_perf.py > gpurun_out/r2w_kerr_perf.log 2>&1; tail -12 gpurun_out/r2w_kerr_perf.log
