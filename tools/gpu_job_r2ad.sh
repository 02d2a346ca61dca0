#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_main.py tests/test_gpu_guards.py -q -m gpu -x > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ad_pytest.log
timeout 600 python tools/remap_perf.py > gpurun_out/r2ad_remap_perf.log 2>&1; grep "float32 nearest" gpurun_out/r2ad_remap_perf.log
