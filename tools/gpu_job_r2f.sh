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
for d in 1 4 2 0; do echo "debug=$d"; LP_REMAP_TMA=1 LP_REMAP_TMA_DEBUG=$d timeout 120 python /tmp/tma_small.py 2>&1 | tail -4 | cut -c1-200; done
echo "rk45 variants"
for v in 0 1 2 3 4 5 6; do for p in 0 1; do echo "variant=$v pow=$p"; LP_RK45_VARIANT=$v LP_RK45_POW=$p timeout 300 python tools/rk45_perf.py 2>&1 | tail -3; done; done
