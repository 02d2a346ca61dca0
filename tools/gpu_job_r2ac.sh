#!/bin/bash
# Stand-alone remap with table constants: GPU suite, remap timings, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2ac_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r2ac_pytest_gpu.log
timeout 600 python tools/remap_perf.py > gpurun_out/r2ac_remap_perf.log 2>&1; cat gpurun_out/r2ac_remap_perf.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ac_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2ac_smoke.log
