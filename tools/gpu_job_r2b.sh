#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_schedule.py -q -m gpu > gpurun_out/r2b_pytest_schedule.log 2>&1; echo "schedule rc=$?"
tail -5 gpurun_out/r2b_pytest_schedule.log
for r in 1 2 4 8; do LP_REPACK_REFILL=$r timeout 300 python tools/repack_perf.py >> gpurun_out/r2b_repack_perf.log 2>&1; done
cat gpurun_out/r2b_repack_perf.log
