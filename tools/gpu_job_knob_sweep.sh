#!/bin/bash
# knob sweep of the frame kernel on the final build: CTA size x steps per trip
mkdir -p gpurun_out
for B in 64 128 256; do
  LP_TRACE_BLOCK=$B python tools/render_knob_perf.py LP_RENDER_TRIP 2 4 2>&1 | sed "s/^/LP_TRACE_BLOCK=$B /"
done > gpurun_out/r2af_knob_sweep2.log
cat gpurun_out/r2af_knob_sweep2.log
