#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rk45.py tests/test_gpu_schedule.py -q -m gpu > gpurun_out/r2j_pytest.log 2>&1; echo "rk45+schedule rc=$?"; tail -8 gpurun_out/r2j_pytest.log
grep -h "worst rel err\|equatorial vs" gpurun_out/r2j_pytest.log
