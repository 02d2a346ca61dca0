#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py tests/test_gpu_rk45.py tests/test_gpu_binet.py -q -m gpu -x > gpurun_out/r2c_pytest_new.log 2>&1; echo "new tests rc=$?"
tail -25 gpurun_out/r2c_pytest_new.log
for t in 1 2 4; do LP_RENDER_TILE_H=$t timeout 300 python tools/repack_perf.py >> gpurun_out/r2c_tile_perf.log 2>&1; done
cat gpurun_out/r2c_tile_perf.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_schedule.py --deselect tests/test_gpu_rk45.py --deselect tests/test_gpu_binet.py > gpurun_out/r2c_pytest_rest.log 2>&1; echo "rest rc=$?"
tail -5 gpurun_out/r2c_pytest_rest.log
