#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rk45.py tests/test_gpu_frame.py -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; echo "rk45+frame rc=$?"; tail -15 gpurun_out/r2i_pytest.log
grep -h "worst rel err\|distinct alphas\|equatorial vs" gpurun_out/r2i_pytest.log
