"""Small driver for ncu captures: N launches of the 4K strict tracer on a resident alpha table."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_path_tracer_b200 import image_lens as il
from light_path_tracer_b200.metrics import Schwarzschild
H, W = 2160, 3840
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
vfov = np.radians(40.0); fov = (2*np.arctan(np.tan(vfov/2)*W/H), vfov)
m = Schwarzschild(1.0)
a = il.build_alpha_lookup((H, W), fov, device=True)
for _ in range(n):
    fa, w = m.trace_alpha_table(a, 100.0, flags=flags)
torch.cuda.synchronize()
# sin agreement between the device libm and the host libm on frame-like alphas
x = a.double().flatten()[::7].contiguous()
s_dev = torch.sin(x).cpu().numpy(); s_host = np.sin(x.cpu().numpy())
print("device sin != host sin: %.4f %% of %d" % (100.0*np.mean(s_dev != s_host), x.numel()))
print("ok")
