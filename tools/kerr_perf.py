"""Timing of the Kerr tracer on a frame-shaped lookup (a = 0.9, r_obs = 100, equatorial)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_path_tracer_b200 import image_lens as il, _device as dev
from light_path_tracer_b200.metrics import Kerr
m = Kerr(1.0, 0.9)
for (H, W) in [(270, 480), (1080, 1920), (2160, 3840)][: int(sys.argv[1]) if len(sys.argv) > 1 else 3]:
    vfov = np.radians(40.0); fov = (2*np.arctan(np.tan(vfov/2)*W/H), vfov)
    a = il.build_alpha_lookup((H, W), fov, device=True)
    cam = dev.camera_vector((H, W), fov, (0.0, 0.0), il._psi_frame)
    steps = torch.empty((H, W, 2), dtype=torch.int32, device="cuda")
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    all_ms = []
    for rep in range(1 + reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fa, w = m.trace_alpha_table_2d(a, cam, 100.0, np.pi/2, steps=steps)
        e1.record(); torch.cuda.synchronize()
        if rep:
            all_ms.append(e0.elapsed_time(e1))
    ms = float(np.median(all_ms))
    if reps > 1:
        print("   runs (ms):", " ".join("%.2f" % t for t in all_ms))
    att = steps[..., 1].double()
    print("%dx%d full frame (no mirror): %.2f ms, %.3e rays/s, mean attempts %.1f (max %d), attempts/s %.3e, escaped %d"
          % (W, H, ms, H*W/ms*1e3, att.mean().item(), int(att.max().item()), att.sum().item()/ms*1e3,
             int(torch.isfinite(fa).sum().item())), flush=True)
