#!/bin/bash
# A/B timing of several builds of liblightpath.so on the same box: put the variants as tools/ab/liblightpath_*.so
# (git-ignored, they travel with the snapshot); each is copied over _C/liblightpath.so in turn, twice
mkdir -p gpurun_out
cp light_path_tracer_b200/_C/liblightpath.so /tmp/lib_orig.so
for rep in 1 2; do for f in tools/ab/liblightpath_*.so; do
  cp $f light_path_tracer_b200/_C/liblightpath.so
  echo "$(basename $f): $(python tools/render_knob_perf.py LP_NONE default 2>&1 | tail -1 | sed 's/ | 1920.*//')"
done; done > gpurun_out/ab_render_perf.log 2>&1
cp /tmp/lib_orig.so light_path_tracer_b200/_C/liblightpath.so
cat gpurun_out/ab_render_perf.log
