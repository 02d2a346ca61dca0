#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/tma_small.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from light_path_tracer_b200 import image_lens as il
from light_path_tracer_b200.metrics import Schwarzschild
H, W = 270, 480
vfov = np.radians(40.0); fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
m = Schwarzschild(1.0)
src = torch.rand(H, W, 3, device='cuda')
a = il.build_alpha_lookup((H, W), fov, device=True)
fa, w = m.trace_alpha_table(a, 100.0)
out = il.render_lensed_image(src, a, fa, w, 0.0, fov)
torch.cuda.synchronize()
print('ok', float(out.sum()))
PY
LP_REMAP_TMA=1 timeout 120 python /tmp/tma_small.py > gpurun_out/r2e_tma_small.log 2>&1; echo "tma small rc=$?"; tail -3 gpurun_out/r2e_tma_small.log
if ! grep -q "^ok" gpurun_out/r2e_tma_small.log; then
  LP_REMAP_TMA=1 timeout 300 compute-sanitizer --tool memcheck python /tmp/tma_small.py > gpurun_out/r2e_tma_sanitizer.log 2>&1; tail -40 gpurun_out/r2e_tma_sanitizer.log
  exit 0
fi
LP_REMAP_TMA=1 timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_main.py -q -m gpu -x > gpurun_out/r2e_pytest_frame_tma.log 2>&1; echo "frame tests with TMA remap rc=$?"
tail -15 gpurun_out/r2e_pytest_frame_tma.log
LP_REMAP_TMA=1 timeout 300 python tools/remap_perf.py > gpurun_out/r2e_remap_perf_tma.log 2>&1
cat gpurun_out/r2e_remap_perf_tma.log
