set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_final.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
for c in render_hybrid remap rk45 trace_hybrid kerr; do
  python tools/ncu_case.py $c > gpurun_out/plain_$c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lp_render_kernel|lp_remap|lp_rk45_kernel|lp_trace_kernel|lp_kerr_kernel" -s 1 -c 1 -f -o gpurun_out/prof_r1d_$c python tools/ncu_case.py $c > gpurun_out/ncu_$c.log 2>&1
done
