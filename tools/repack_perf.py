"""Lane re-packing vs one ray per thread for the fused frame kernel, over frames of different
divergence: kernel time (CUDA events, best of 5 after warm-up), lane efficiency of both schedules,
the per-launch predictor's choice.  LP_REPACK_REFILL (read once per process) is swept by the caller.

    python tools/repack_perf.py            # prints one line per frame"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_path_tracer_b200 import image_lens as il, _device as dev, _lib  # noqa: E402
from light_path_tracer_b200.metrics import Schwarzschild  # noqa: E402

metric = Schwarzschild(1.0)
e = _lib.ext()
FRAMES = [("bench 4K r100 v40", 2160, 3840, 40.0, 100.0, (0.0, 0.0)),
          ("1080p r100 v40", 1080, 1920, 40.0, 100.0, (0.0, 0.0)),
          ("1024 r15 v40", 1024, 1024, 40.0, 15.0, (0.0, 0.0)),
          ("1024 r30 v40", 1024, 1024, 40.0, 30.0, (0.0, 0.0)),
          ("1024 r1000 v40", 1024, 1024, 40.0, 1000.0, (0.0, 0.0)),
          ("1024 r100 v12 zoom", 1024, 1024, 12.0, 100.0, (0.0, 0.0)),
          ("2048 r100 v12 zoom", 2048, 2048, 12.0, 100.0, (0.0, 0.0)),
          ("1024 r100 v6 zoom", 1024, 1024, 6.0, 100.0, (0.0, 0.0)),
          ("1024 r300 v4 zoom", 1024, 1024, 4.0, 300.0, (0.0, 0.0)),
          ("smoke 96x128 v12", 96, 128, 12.0, 100.0, (0.0, 0.0)),
          ("512 r100 v20 off", 512, 512, 20.0, 100.0, (0.05, -0.08))]


def best(fn, k=5):
    fn(); fn()
    ts = []
    for _ in range(k):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


print("LP_REPACK_REFILL=%s LP_RENDER_TILE_H=%s" % (os.environ.get("LP_REPACK_REFILL", "default"), os.environ.get("LP_RENDER_TILE_H", "default")))
for name, H, W, vdeg, r_obs, psi in FRAMES:
    vfov = np.radians(vdeg)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    src = torch.rand(H, W, 3, device="cuda")
    out = torch.empty_like(src)
    res = {}
    for tag, fl in (("one", 0), ("repack", dev.TRACE_REPACK)):
        st = dev.new_stats()
        il.render_frame(src, fov, r_obs, metric, psi=psi, out=out, stats=st, flags=dev.TRACE_HYBRID | fl)
        s = dev.read_stats(st)
        ms = best(lambda: il.render_frame(src, fov, r_obs, metric, psi=psi, out=out, flags=dev.TRACE_HYBRID | fl))
        res[tag] = (ms, s["lane_efficiency"], s["sum_steps"] / max(s["n_rays"], 1), s["max_steps"])
    print("%-22s one %8.4f ms (lane_eff %.3f)  repack %8.4f ms (lane_eff %.3f)  ratio %.3f  steps/ray %.1f max %d"
          % (name, res["one"][0], res["one"][1], res["repack"][0], res["repack"][1], res["one"][0] / res["repack"][0],
             res["one"][2], res["one"][3]), flush=True)
