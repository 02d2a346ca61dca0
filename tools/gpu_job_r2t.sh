#!/bin/bash
# Head / tail trims of the frame kernel (own acos, remap from the tracer's direction, straight-line reciprocals in
# the FMA path): GPU suite, mode timings, bench line, ncu capture (with source) of the frame kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2t_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -15 gpurun_out/r2t_pytest_gpu.log
timeout 600 python tools/quick_perf3.py > gpurun_out/r2t_modes_perf3.log 2>&1; cat gpurun_out/r2t_modes_perf3.log
timeout 600 python bench.py > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2t_bench_n1.json; tail -3 gpurun_out/r2t_bench_n1.err
python tools/ncu_case.py render_u8 > gpurun_out/r2t_plain_render_u8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lp_render_kernel" -s 1 -c 1 -f -o gpurun_out/prof_r2t_render_u8 python tools/ncu_case.py render_u8 > gpurun_out/r2t_ncu_render_u8.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
