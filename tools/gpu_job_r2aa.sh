#!/bin/bash
# Two-threshold hybrid re-trace, straight-line sincos in the Kerr right-hand side: GPU suite, Binet fuzz (seed 31 again
# and a new one), Kerr fuzz and timing, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2aa_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -6 gpurun_out/r2aa_pytest_gpu.log
timeout 900 python tools/parity_fuzz.py 31 32 > gpurun_out/r2_parity_fuzz_binet_seed31.log 2>&1; echo "binet fuzz 31 rc=$?"; tail -2 gpurun_out/r2_parity_fuzz_binet_seed31.log
timeout 900 python tools/parity_fuzz.py 41 32 > gpurun_out/r2_parity_fuzz_binet_seed41.log 2>&1; echo "binet fuzz 41 rc=$?"; tail -2 gpurun_out/r2_parity_fuzz_binet_seed41.log
timeout 900 python tools/parity_fuzz_kerr.py 33 > gpurun_out/r2_parity_fuzz_kerr_seed33.log 2>&1; echo "kerr fuzz rc=$?"; tail -2 gpurun_out/r2_parity_fuzz_kerr_seed33.log
timeout 600 python tools/kerr_perf.py 3 3 > gpurun_out/r2aa_kerr_perf.log 2>&1; tail -4 gpurun_out/r2aa_kerr_perf.log
timeout 600 python bench.py > gpurun_out/r2aa_bench_n1.json 2> gpurun_out/r2aa_bench_n1.err; echo "bench rc=$?"; head -c 500 gpurun_out/r2aa_bench_n1.json; tail -3 gpurun_out/r2aa_bench_n1.err
