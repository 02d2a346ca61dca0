"""Frame-pipeline parity fuzz: random frame shapes / FOV / camera offsets / observer radii / source
dtypes.  Stage-isolated against the oracle: alpha table (float32 steps), lookups
from the oracle's alpha (exact), remap from the oracle's lookups (exact INTEGER source index, incl.
loop_around, channels 1/3, uint8/float32/float64, with and without winding)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lp_oracle as O
from light_path_tracer_b200 import image_lens as il
from light_path_tracer_b200.metrics import Schwarzschild

O.build()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
bad = 0
tot_px = tot_alpha = tot_fa = tot_remap = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    H = int(rng.integers(40, 400)); W = int(rng.integers(40, 600))
    if rng.random() < 0.5:
        W -= W % 4
    vfov = np.radians(rng.uniform(3, 110))
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    psi = (float(rng.normal(0, 0.3)), float(rng.normal(0, 0.3))) if rng.random() < 0.7 else (0.0, 0.0)
    M = float(rng.choice([1.0, 1.0, 0.4, 3.0]))
    r_obs = float(M * 10 ** rng.uniform(0.7, 3))
    metric = Schwarzschild(M)
    a_ref = O.build_alpha_lookup((H, W), fov, psi=psi)
    a = il.build_alpha_lookup((H, W), fov, psi=psi)
    d = np.abs(a.view(np.int32).astype(np.int64) - a_ref.view(np.int32))
    n_alpha = int((d > 0).sum())
    fa_ref, w_ref, _, _ = O.precompute_final_alpha_lookup(a_ref, M, r_obs)
    fa, w, _, _ = il.precompute_final_alpha_lookup(a_ref, 0.0, r_obs, metric)
    cls = int((np.isnan(fa) != np.isnan(fa_ref)).sum()) + int((w != w_ref).sum())
    both = np.isfinite(fa) & np.isfinite(fa_ref)
    dfa = np.abs(fa[both].view(np.int32).astype(np.int64) - fa_ref[both].view(np.int32))
    n_fa = int((dfa > 0).sum())
    C = int(rng.choice([1, 3]))
    dt = rng.choice(["f32", "u8", "f64"])
    shape = (H, W, 3) if C == 3 else (H, W)
    src = rng.random(shape)
    src = {"f32": src.astype(np.float32), "f64": src, "u8": (src * 255).astype(np.uint8)}[dt]
    loop = bool(rng.random() < 0.3)
    wl = w_ref if rng.random() < 0.8 else None
    ref = O.render_lensed_image(src, fa_ref, wl, fov, loop, psi)
    out = il.render_lensed_image(src, a_ref, fa_ref, wl, 0.0, fov, loop, psi)
    n_remap = int((out != ref).reshape(H, W, -1).any(-1).sum())
    ok = d.max() <= 1 and n_alpha <= max(4, 2e-5 * H * W) and cls == 0 and (dfa.max() if dfa.size else 0) <= 1 \
        and n_fa <= max(4, 2e-5 * H * W) and n_remap == 0
    bad += 0 if ok else 1
    tot_px += H * W; tot_alpha += n_alpha; tot_fa += n_fa; tot_remap += n_remap
    print("%s %dx%d vfov=%.1f psi=(%.2f,%.2f) M=%g r_obs=%.4g %s C=%d loop=%d wind=%d | alpha!= %d, class/wind!= %d, fa32!= %d, remap px!= %d"
          % ("ok  " if ok else "FAIL", W, H, np.degrees(vfov), psi[0], psi[1], M, r_obs, dt, C, loop, wl is not None,
             n_alpha, cls, n_fa, n_remap), flush=True)
print("pixels %d: alpha off-by-one-float32-step %d, final_alpha off-by-one %d, remapped pixels different %d; violations: %d"
      % (tot_px, tot_alpha, tot_fa, tot_remap, bad))
sys.exit(1 if bad else 0)
