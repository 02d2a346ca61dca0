#!/bin/bash
# Instruction trims outside the loop (write-out geometry from the host, integer range guards, crossing in the scaled
# variable, half-orbit count from the sincos quadrant): GPU suite, mode timings, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -15 gpurun_out/r2u_pytest_gpu.log
timeout 600 python tools/quick_perf3.py > gpurun_out/r2u_modes_perf3.log 2>&1; cat gpurun_out/r2u_modes_perf3.log
timeout 600 python bench.py > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2u_bench_n1.json; tail -3 gpurun_out/r2u_bench_n1.err
