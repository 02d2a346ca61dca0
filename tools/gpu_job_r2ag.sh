#!/bin/bash
# New launch defaults (256-thread CTAs, steps per trip by frame size): GPU suite, smoke, mode timings, knob tool at defaults, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2ag_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r2ag_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ag_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2ag_smoke.log
timeout 600 python tools/quick_perf3.py > gpurun_out/r2ag_modes_perf3.log 2>&1; cat gpurun_out/r2ag_modes_perf3.log
python tools/render_knob_perf.py LP_NONE default > gpurun_out/r2ag_render_perf.log 2>&1; cat gpurun_out/r2ag_render_perf.log
timeout 600 python bench.py > gpurun_out/r2ag_bench_n1.json 2> gpurun_out/r2ag_bench_n1.err; echo "bench rc=$?"; head -c 420 gpurun_out/r2ag_bench_n1.json; echo; tail -1 gpurun_out/r2ag_bench_n1.err
