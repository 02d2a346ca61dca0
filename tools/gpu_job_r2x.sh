#!/bin/bash
# RK45 equatorial kernel: occupancy / refill sweep on the trimmed kernel, ncu with source counts
mkdir -p gpurun_out
for mb in 3 4 5; do for rf in 4 8 16; do
  echo "LP_RK45_EQ_MINB=$mb LP_RK45_REFILL=$rf"; LP_RK45_EQ_MINB=$mb LP_RK45_REFILL=$rf timeout 300 python tools/rk45_perf.py 3 2>&1 | tail -1
done; done > gpurun_out/r2x_rk45_sweep.log 2>&1
cat gpurun_out/r2x_rk45_sweep.log
python tools/ncu_case.py rk45 > gpurun_out/r2x_plain_rk45.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lp_rk45_eq_kernel" -s 1 -c 1 -f -o gpurun_out/prof_r2x_rk45 python tools/ncu_case.py rk45 > gpurun_out/r2x_ncu_rk45.log 2>&1
echo "ncu rc=$?"
