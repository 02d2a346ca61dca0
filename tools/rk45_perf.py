"""Timing of kernel (1b) on a frame-shaped alpha table (BASELINE config 3)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_path_tracer_b200 import image_lens as il, geodesic_tracer as gt
from light_path_tracer_b200.metrics import Schwarzschild
m = Schwarzschild(1.0)
for (H, W) in [(270, 480), (1080, 1920), (2160, 3840)][: int(sys.argv[1]) if len(sys.argv) > 1 else 3]:
    vfov = np.radians(40.0); fov = (2*np.arctan(np.tan(vfov/2)*W/H), vfov)
    a = il.build_alpha_lookup((H, W), fov, device=True).double()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        state, lam, oc, ns = gt.trace_rays(m, 100.0, a)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nsf = ns.reshape(-1, 2).double()
    attempts = ((nsf[:, 1] - 2) / 6).sum().item()
    print("%dx%d: %.2f ms, %.3e rays/s, mean points %.1f, mean attempts %.1f, attempts/s %.3e, escaped %d captured %d invalid %d"
          % (W, H, ms, H*W/ms*1e3, nsf[:, 0].mean().item(), attempts/(H*W), attempts/ms*1e3,
             (oc == 1).sum().item(), (oc == -1).sum().item(), (oc == 0).sum().item()), flush=True)
