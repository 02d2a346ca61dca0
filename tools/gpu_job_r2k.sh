#!/bin/bash
# re-entry validation of the committed tree: full GPU suite, bench lines, then the round-2 profiles
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 gpurun_out/r2k_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2k_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err; echo "bench ref rc=$?"
LP_REMAP_TMA=1 timeout 300 python tools/remap_perf.py > gpurun_out/r2k_remap_perf_tma.log 2>&1; cat gpurun_out/r2k_remap_perf_tma.log
timeout 300 python tools/remap_perf.py > gpurun_out/r2k_remap_perf.log 2>&1; cat gpurun_out/r2k_remap_perf.log
bash tools/gpu_job_r2_prof.sh
