"""Scratch timing of the tracer kernels on device-resident data (not the bench)."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_path_tracer_b200 import _device as dev, _lib, image_lens as il
from light_path_tracer_b200.metrics import Schwarzschild

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

e = _lib.ext()
# FP64 peak
sink = torch.empty(148*32*256, dtype=torch.float64, device='cuda')
iters = 20000
t, _ = timeit(lambda: e.bench_dfma(148*32, 256, iters, sink))
print("DFMA peak: %.2f TFLOP/s (%.3f ms)" % (148*32*256*iters*16/ (t*1e-3)/1e12, t))
for (H, W, r_obs) in [(1080,1920,100.0),(2160,3840,100.0),(2160,3840,15.0),(2160,3840,1000.0),(4320,7680,100.0)]:
    vfov = np.radians(40.0); fov = (2*np.arctan(np.tan(vfov/2)*W/H), vfov)
    m = Schwarzschild(1.0)
    a = il.build_alpha_lookup((H,W), fov, device=True)
    stats = dev.new_stats()
    fa, w = m.trace_alpha_table(a, r_obs, stats=stats)
    s = dev.read_stats(stats)
    flops = 43*s['sum_steps'] + 40*s['n_rays']
    for flags in (0, 1):
        t, tm = timeit(lambda: m.trace_alpha_table(a, r_obs, flags=flags))
        print("%dx%d r_obs=%g flags=%d: %.3f ms (med %.3f)  %.2f Grays/s  %.2f TFLOP/s(alg)  lane_eff=%.3f steps/ray=%.1f" % (W,H,r_obs,flags,t,tm,H*W/t/1e6, flops/t/1e9, s['lane_efficiency'], s['sum_steps']/s['n_rays']))
    src = torch.rand(H,W,3,device='cuda')
    t, tm = timeit(lambda: il.render_frame(src, fov, r_obs, m))
    print("   fused render_frame: %.3f ms" % t)
    t, tm = timeit(lambda: il.render_lensed_image(src, a, fa, w, 0.0, fov))
    print("   remap alone: %.3f ms  (%.1f GB/s at 30 B/px)" % (t, H*W*30/t/1e6))
    t, tm = timeit(lambda: il.build_alpha_lookup((H,W), fov, device=True))
    print("   alpha lookup alone: %.3f ms" % t)
