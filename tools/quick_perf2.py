import sys, os
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
sink = torch.empty(148*32*256, dtype=torch.float64, device='cuda')
t, _ = timeit(lambda: e.bench_dfma(148*32, 256, 20000, sink))
peak = 148*32*256*20000*16/(t*1e-3)/1e12
cfgs = [(2160,3840,100.0)] if len(sys.argv) < 2 else [(2160,3840,100.0),(2160,3840,15.0),(1080,1920,100.0),(4320,7680,100.0)]
for (H, W, r_obs) in cfgs:
    vfov = np.radians(40.0); fov = (2*np.arctan(np.tan(vfov/2)*W/H), vfov)
    m = Schwarzschild(1.0)
    a = il.build_alpha_lookup((H,W), fov, device=True)
    stats = dev.new_stats()
    fa, w = m.trace_alpha_table(a, r_obs, stats=stats)
    s = dev.read_stats(stats)
    flops = 43*s['sum_steps'] + 40*s['n_rays']
    out = "block=%s %dx%d r=%g peak=%.1fTF eff=%.3f |" % (os.environ.get('LP_TRACE_BLOCK','def'), W,H,r_obs, peak, s['lane_efficiency'])
    for flags in (0, 1):
        t, tm = timeit(lambda: m.trace_alpha_table(a, r_obs, flags=flags))
        out += " flags=%d: %.3f ms %.2f Grays/s %.2f TF (%.1f%%) |" % (flags, t, H*W/t/1e6, flops/t/1e9, 100*flops/t/1e9/peak)
    t, tm = timeit(lambda: m.trace_alpha_table(a, r_obs, stats=stats))
    out += " +stats %.3f ms" % t
    src = torch.rand(H,W,3,device='cuda')
    t, tm = timeit(lambda: il.render_frame(src, fov, r_obs, m))
    out += " | fused frame %.3f ms" % t
    print(out)
