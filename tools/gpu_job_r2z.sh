#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_schedule.py tests/test_gpu_knobs.py tests/test_gpu_frame.py -q -m gpu -x > gpurun_out/r2z_pytest_sched.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2z_pytest_sched.log
