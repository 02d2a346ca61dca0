#!/bin/bash
# round 2 profiles: launch list of the bench command, ncu --set full of the kernels that changed
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_4k.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
for c in render_u8 render_hybrid remap remap_tma rk45 rk45_full; do
  python tools/ncu_case.py $c > gpurun_out/r2_plain_$c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lp_render_kernel|lp_remap|lp_rk45" -s 1 -c 1 -f -o gpurun_out/prof_r2_$c python tools/ncu_case.py $c > gpurun_out/r2_ncu_$c.log 2>&1
  echo "$c rc=$?"
done
ls -la gpurun_out/*.ncu-rep
