#!/bin/bash
# Kerr: error norm over a straight-line reciprocal, select-based min / max: tests, fuzz, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kerr.py tests/test_gpu_main.py tests/test_gpu_rk45.py -q -m gpu -x > gpurun_out/r2aj_pytest_kerr.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2aj_pytest_kerr.log
timeout 900 python tools/parity_fuzz_kerr.py 33 > gpurun_out/r2_parity_fuzz_kerr_seed33.log 2>&1; echo "kerr fuzz rc=$?"; tail -1 gpurun_out/r2_parity_fuzz_kerr_seed33.log
timeout 900 python tools/parity_fuzz_kerr.py 43 > gpurun_out/r2_parity_fuzz_kerr_seed43.log 2>&1; echo "kerr fuzz 43 rc=$?"; tail -1 gpurun_out/r2_parity_fuzz_kerr_seed43.log
timeout 600 python tools/kerr_perf.py 3 3 > gpurun_out/r2aj_kerr_perf.log 2>&1; tail -2 gpurun_out/r2aj_kerr_perf.log
