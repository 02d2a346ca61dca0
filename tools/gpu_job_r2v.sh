#!/bin/bash
# ncu captures with source-level counters: frame kernel (current build), equatorial RK45 kernel, queued Kerr kernel
mkdir -p gpurun_out
for c in render_u8:lp_render_kernel rk45:lp_rk45_eq_kernel kerr:lp_kerr_queued_kernel; do
  case=${c%%:*}; kern=${c##*:}
  python tools/ncu_case.py $case > gpurun_out/r2v_plain_$case.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"$kern" -s 1 -c 1 -f -o gpurun_out/prof_r2v_$case python tools/ncu_case.py $case > gpurun_out/r2v_ncu_$case.log 2>&1
  echo "$case ncu rc=$?"
done
ls -la gpurun_out/*.ncu-rep
