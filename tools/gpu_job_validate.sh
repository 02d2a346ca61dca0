#!/bin/bash
# Final 1-GPU validation of round 2: GPU suite, parity fuzz (Binet, frames, Kerr), bench line, reference arm, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2ab_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r2ab_pytest_gpu.log
timeout 900 python tools/parity_fuzz.py 31 32 > gpurun_out/r2_parity_fuzz_binet_seed31.log 2>&1; echo "binet fuzz 31 rc=$?"; tail -1 gpurun_out/r2_parity_fuzz_binet_seed31.log
timeout 900 python tools/parity_fuzz.py 41 32 > gpurun_out/r2_parity_fuzz_binet_seed41.log 2>&1; echo "binet fuzz 41 rc=$?"; tail -1 gpurun_out/r2_parity_fuzz_binet_seed41.log
timeout 900 python tools/parity_fuzz_frames.py 32 > gpurun_out/r2_parity_fuzz_frames_seed32.log 2>&1; echo "frames fuzz rc=$?"; tail -1 gpurun_out/r2_parity_fuzz_frames_seed32.log
timeout 900 python tools/parity_fuzz_kerr.py 33 > gpurun_out/r2_parity_fuzz_kerr_seed33.log 2>&1; echo "kerr fuzz rc=$?"; tail -1 gpurun_out/r2_parity_fuzz_kerr_seed33.log
timeout 600 python bench.py > gpurun_out/r2ab_bench_n1.json 2> gpurun_out/r2ab_bench_n1.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2ab_bench_n1.json; echo; tail -2 gpurun_out/r2ab_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ab_bench_reference_arm.json 2> gpurun_out/r2ab_bench_reference_arm.err; echo "reference arm rc=$?"; head -c 400 gpurun_out/r2ab_bench_reference_arm.json; echo
timeout 600 python tools/quick_perf3.py > gpurun_out/r2ab_modes_perf3.log 2>&1; head -3 gpurun_out/r2ab_modes_perf3.log
for c in render_u8:lp_render_kernel rk45:lp_rk45_eq_kernel; do
  case=${c%%:*}; kern=${c##*:}
  python tools/ncu_case.py $case > gpurun_out/r2ab_plain_$case.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$kern" -s 1 -c 1 -f -o gpurun_out/prof_r2ab_$case python tools/ncu_case.py $case > gpurun_out/r2ab_ncu_$case.log 2>&1
  echo "$case ncu rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ab_launches_bench_4k.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2ab_ncu_launches.log 2>&1; echo "launch list rc=$?"
