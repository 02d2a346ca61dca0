#!/bin/bash
# Scaled-variable (v = 3M u) FMA loop: full GPU suite, bench line, mode timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -15 gpurun_out/r2s_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err; echo "bench rc=$?"; head -c 1500 gpurun_out/r2s_bench_n1.json; tail -3 gpurun_out/r2s_bench_n1.err
timeout 600 python tools/quick_perf3.py > gpurun_out/r2s_modes_perf3.log 2>&1; cat gpurun_out/r2s_modes_perf3.log
