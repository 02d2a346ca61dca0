#!/bin/bash
# Nystrom-form FMA loop: full GPU suite, bench line, mode timings, ncu capture (with source) of the frame kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -15 gpurun_out/r2m_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; echo "bench rc=$?"; head -c 1500 gpurun_out/r2m_bench_n1.json; tail -3 gpurun_out/r2m_bench_n1.err
timeout 600 python tools/quick_perf3.py > gpurun_out/r2m_modes_perf3.log 2>&1; cat gpurun_out/r2m_modes_perf3.log
python tools/ncu_case.py render_u8 > gpurun_out/r2m_plain_render_u8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lp_render_kernel" -s 1 -c 1 -f -o gpurun_out/prof_r2m_render_u8 python tools/ncu_case.py render_u8 > gpurun_out/r2m_ncu_render_u8.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
