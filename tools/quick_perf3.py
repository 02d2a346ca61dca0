"""Kernel timings by arithmetic mode (0 strict, 1 fma, 4 hybrid) + hybrid/strict frame parity."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_path_tracer_b200 import _device as dev, _lib, image_lens as il
from light_path_tracer_b200.metrics import Schwarzschild
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(fn, n=15, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.mean(ts))
e = _lib.ext()
sink = torch.empty(148*32*256, dtype=torch.float64, device='cuda')
t, _ = timeit(lambda: e.bench_dfma(148*32, 256, 20000, sink), n=5)
peak = 148*32*256*20000*16/(t*1e-3)/1e12
print("fp64 peak %.2f TF" % peak)
cfgs = [(2160,3840,100.0,(0.,0.)), (1080,1920,100.0,(0.,0.)), (2160,3840,15.0,(0.,0.)), (2160,3840,100.0,(0.1,0.2)), (4320,7680,100.0,(0.,0.))]
for (H, W, r_obs, psi) in cfgs:
    vfov = np.radians(40.0); fov = (2*np.arctan(np.tan(vfov/2)*W/H), vfov)
    m = Schwarzschild(1.0)
    a = il.build_alpha_lookup((H,W), fov, psi=psi, device=True)
    stats = dev.new_stats()
    fa, w = m.trace_alpha_table(a, r_obs, stats=stats)
    s = dev.read_stats(stats)
    flops = 43*s['sum_steps'] + 40*s['n_rays']
    out = "%dx%d r=%g psi=%s eff=%.3f maxsteps=%d |" % (W,H,r_obs,psi, s['lane_efficiency'], s['max_steps'])
    for flags in (0, 1, 4):
        t, tm = timeit(lambda: m.trace_alpha_table(a, r_obs, flags=flags))
        out += " trace f%d %.3f/%.3f ms %.1f%% |" % (flags, t, tm, 100*flops/tm/1e9/peak)
    src = torch.rand(H,W,3,device='cuda')
    frames = {}
    for flags in (0, 1, 4):
        t, tm = timeit(lambda: il.render_frame(src, fov, r_obs, m, psi=psi, flags=flags))
        out += " render f%d %.3f/%.3f ms %.1f%% |" % (flags, t, tm, 100*flops/tm/1e9/peak)
        frames[flags] = il.render_frame(src, fov, r_obs, m, psi=psi, flags=flags, return_lookups=True)
    t, tm = timeit(lambda: il.render_lensed_image(src, a, fa, w, 0.0, fov, psi=psi))
    out += " remap %.3f/%.3f ms %.0f GB/s" % (t, tm, H*W*30/tm/1e6)
    for flags in (1, 4):
        px = (frames[0][0] != frames[flags][0]).any(dim=-1).sum().item()
        fa0, fa1 = frames[0][1], frames[flags][1]
        nanmis = (torch.isnan(fa0) != torch.isnan(fa1)).sum().item()
        ok = ~torch.isnan(fa0) & ~torch.isnan(fa1)
        fad = (fa0[ok] != fa1[ok]).sum().item()
        wd = (frames[0][2] != frames[flags][2]).sum().item()
        out += " | f%d vs strict: px!= %d, class!= %d, fa32!= %d, w!= %d" % (flags, px, nanmis, fad, wd)
    print(out, flush=True)
