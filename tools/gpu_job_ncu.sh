#!/bin/bash
# ncu capture (source counters) of the frame kernel with the final launch defaults; launch list of the bench
mkdir -p gpurun_out
python tools/ncu_case.py render_u8 > gpurun_out/r2ah_plain_render_u8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lp_render_kernel" -s 1 -c 1 -f -o gpurun_out/prof_r2ah_render_u8 python tools/ncu_case.py render_u8 > gpurun_out/r2ah_ncu_render_u8.log 2>&1
echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ah_launches_bench_4k.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2ah_ncu_launches.log 2>&1; echo "launch list rc=$?"
