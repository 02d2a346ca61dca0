#!/bin/bash
mkdir -p gpurun_out
LP_REMAP_TMA=1 timeout 120 python /dev/stdin > gpurun_out/r2h_tma_small.log 2>&1 <<'PY'
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
tail -2 gpurun_out/r2h_tma_small.log
if grep -q "^ok" gpurun_out/r2h_tma_small.log; then
  LP_REMAP_TMA=1 timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_main.py -q -m gpu > gpurun_out/r2h_pytest_frame_tma.log 2>&1; echo "frame tests with TMA remap rc=$?"
  tail -12 gpurun_out/r2h_pytest_frame_tma.log
  LP_REMAP_TMA=1 timeout 300 python tools/remap_perf.py > gpurun_out/r2h_remap_perf_tma.log 2>&1
  cat gpurun_out/r2h_remap_perf_tma.log
fi
timeout 900 python -m pytest tests/test_gpu_frame.py -q -m gpu -k "oracle" > gpurun_out/r2h_pytest_frame_attrib.log 2>&1; echo "attribution tests rc=$?"; tail -12 gpurun_out/r2h_pytest_frame_attrib.log
echo "rk45 eq variants"
for m in 3 4 5 6; do for p in 0 1; do echo "eq minb=$m pow=$p"; LP_RK45_EQ=1 LP_RK45_EQ_MINB=$m LP_RK45_POW=$p timeout 300 python tools/rk45_perf.py 2>&1 | tail -1; done; done
LP_RK45_EQ=1 LP_RK45_POW=1 timeout 600 python -m pytest tests/test_gpu_rk45.py -q -m gpu > gpurun_out/r2h_pytest_rk45_eq.log 2>&1; echo "rk45 tests with EQ+fast pow rc=$?"; tail -8 gpurun_out/r2h_pytest_rk45_eq.log
