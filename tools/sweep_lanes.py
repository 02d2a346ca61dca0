"""Config 5 on one GPU: the 512-frame sweep as a CUDA graph with 1 / 2 / 3 / 4 concurrent chains."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_path_tracer_b200 import image_lens as il, dist as lpdist
from light_path_tracer_b200.metrics import Schwarzschild
Hs = Ws = 1024
yy, xx = torch.meshgrid(torch.arange(Hs, device="cuda"), torch.arange(Ws, device="cuda"), indexing="ij")
r = (((yy // 32) + (xx // 32)) & 1).float()
src = torch.stack([r, 1 - r, xx.float() / Ws], dim=-1).contiguous()
pipe = il.LensPipeline(src, 40.0, Schwarzschild(1.0))
grid = lpdist.sweep_grid()
frames = torch.empty((len(grid), Hs, Ws, 3), device="cuda")
ref = None
for lanes in (1, 2, 3, 4):
    g = pipe.capture_sweep(grid, frames, lanes=lanes)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    chk = frames[::37].clone()
    same = True if ref is None else bool(torch.equal(chk, ref))
    ref = chk if ref is None else ref
    print(json.dumps({"lanes": lanes, "ms_total": float(np.median(ts)), "rays_per_s": len(grid) * Hs * Ws / np.median(ts) * 1e3,
                      "same_frames": same}), flush=True)
