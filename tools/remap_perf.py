import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_path_tracer_b200 import image_lens as il
from light_path_tracer_b200.metrics import Schwarzschild
H, W = 2160, 3840
vfov = np.radians(40.0); fov = (2*np.arctan(np.tan(vfov/2)*W/H), vfov)
m = Schwarzschild(1.0)
a = il.build_alpha_lookup((H, W), fov, device=True)
fa, w = m.trace_alpha_table(a, 100.0)
src = torch.rand(H, W, 3, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(25):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); il.render_lensed_image(src, a, fa, w, 0.0, fov); e1.record(); torch.cuda.synchronize()
    if i >= 5: ts.append(e0.elapsed_time(e1))
print(os.environ.get("LP_REMAP_MINB"), "min %.4f mean %.4f ms -> %.0f GB/s" % (min(ts), np.mean(ts), H*W*30/np.mean(ts)/1e6))
