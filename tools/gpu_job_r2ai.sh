#!/bin/bash
# last check of the driver's own sequence on the final tree: smoke, bench (default flags), reference arm
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ai_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2ai_smoke.log
timeout 600 python bench.py > gpurun_out/r2ai_bench_n1.json 2> gpurun_out/r2ai_bench_n1.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2ai_bench_n1.json; echo
timeout 600 python bench.py --impl reference > gpurun_out/r2ai_bench_reference_arm.json 2> gpurun_out/r2ai_bench_reference_arm.err; echo "reference arm rc=$?"; head -c 300 gpurun_out/r2ai_bench_reference_arm.json; echo
