#!/bin/bash
# round 2, first GPU pass: smoke, new schedule tests, full GPU suite, bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/r2a_smoke.log
timeout 600 python -m pytest tests/test_gpu_schedule.py -q -m gpu > gpurun_out/r2a_pytest_schedule.log 2>&1; echo "schedule rc=$?"
tail -15 gpurun_out/r2a_pytest_schedule.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_schedule.py > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "suite rc=$?"
tail -15 gpurun_out/r2a_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2a_bench_n1.err
cut -c1-1500 gpurun_out/r2a_bench_n1.json
