#!/usr/bin/env python
"""Kerr golden fixtures FROM THE UNMODIFIED REFERENCE (build container only; needs /root/reference):

    python tests/golden/make_golden_kerr.py  ->  tests/golden/kerr_rays.npz, kerr_frames.npz

Same rules as make_golden.py: the reference is imported through oracle/ref_harness.py and every
array that went through a platform-dependent numpy ufunc is stored so that later stages can be
fed the reference's own upstream values.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_harness  # noqa: E402

R = ref_harness.load()
MM, IL = R.metrics, R.image_lens


def rays():
    """metrics.py:419-567 (_kerr_trace_ray_numba) through Kerr.trace_rays_batch, plus the scalar
    helpers of the class (metrics.py:840-944)."""
    rng = np.random.default_rng(20261019)
    out = {}
    cfgs = [(1.0, 0.9, 100.0, np.pi / 2), (1.0, 0.5, 50.0, np.pi / 3), (1.0, 0.998, 30.0, 1.2),
            (2.0, -1.4, 80.0, np.pi / 2), (1.0, 0.9, 20.0, 0.3)]
    meta = []
    for k, (M, a, r_obs, th_obs) in enumerate(cfgs):
        m = MM.Kerr(M, a)
        ac = float(m.alpha_crit(r_obs, th_obs))
        n = 1500
        alpha = np.concatenate([rng.uniform(0, 4 * ac, n), ac * (1 + rng.normal(0, 0.05, n)),
                                rng.uniform(0, np.pi, 300), [0.0, 1e-9, ac, np.pi / 2, 3.0]])
        theta = rng.uniform(-np.pi, np.pi, alpha.size)
        theta[-5:] = [0.0, 0.5, np.pi / 2, -np.pi / 2, np.pi]
        refine = rng.random(alpha.size) < 0.2
        fa = np.full(alpha.size, -1.0)
        w = np.full(alpha.size, -1, dtype=np.int64)
        m.trace_rays_batch(r_obs, alpha, theta, th_obs, refine, fa, w)
        # status per ray from the scalar API (same numba kernel)
        st = np.empty(alpha.size, dtype=np.int8)
        for i in range(alpha.size):
            o = m.trace_ray(r_obs, float(alpha[i]), float(theta[i]), th_obs, axis_refine=bool(refine[i]))[2]
            st[i] = {"escaped": 1, "captured": -1, "invalid": 0}[o]
        p = "c%d_" % k
        out.update({p + "alpha": alpha, p + "theta": theta, p + "refine": refine, p + "fa": fa, p + "w": w,
                    p + "status": st})
        b = np.array([m.viewing_angle_to_impact_parameter(x, r_obs, th_obs) for x in (0.01, 0.1, 1.0)])
        ic = np.array(m.initial_conditions(r_obs, 0.07, 0.4, th_obs))
        rhs = np.array(m.geodesic_equations(0.0, ic))
        meta.append([M, a, r_obs, th_obs, ac, m.r_plus, m.capture_radius(), *m._unstable_photon_r(), *b])
        out[p + "ic"] = ic
        out[p + "rhs"] = rhs
    out["cfg"] = np.array(meta)
    np.savez_compressed(os.path.join(HERE, "kerr_rays.npz"), **out)
    return {k: int(v.size) for k, v in out.items() if k.endswith("alpha")}


def frames():
    """image_lens.precompute_final_alpha_lookup_2d (image_lens.py:185-280) + render on small frames."""
    out = {}
    cfgs = {"eq": dict(H=40, W=64, vfov_deg=14.0, psi=(0.0, 0.0), M=1.0, a=0.9, r_obs=100.0, theta_obs=np.pi / 2),
            "incl": dict(H=36, W=48, vfov_deg=20.0, psi=(0.02, -0.03), M=1.0, a=0.6, r_obs=60.0, theta_obs=1.0),
            "odd": dict(H=33, W=50, vfov_deg=16.0, psi=(0.0, 0.04), M=1.0, a=-0.8, r_obs=100.0, theta_obs=np.pi / 2)}
    for tag, c in cfgs.items():
        H, W = c["H"], c["W"]
        vfov = np.radians(c["vfov_deg"])
        fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
        m = MM.Kerr(c["M"], c["a"])
        ac = m.alpha_crit(c["r_obs"], c["theta_obs"])
        alpha = IL.build_alpha_lookup((H, W), fov, psi=c["psi"])
        fa, w, n_total, n_traced = IL.precompute_final_alpha_lookup_2d(
            alpha, fov, ac, c["r_obs"], m, theta_obs=c["theta_obs"], psi=c["psi"])
        yy, xx = np.mgrid[0:H, 0:W]
        src = np.stack([((yy // 4 + xx // 4) & 1), 1 - ((yy // 4 + xx // 4) & 1), xx / W], -1).astype(np.float32)
        img = IL.render_lensed_image(src, alpha, fa, w, ac, fov, False, psi=c["psi"])
        out.update({tag + "_alpha32": alpha, tag + "_fa32": fa, tag + "_w16": w, tag + "_src": src, tag + "_img": img,
                    tag + "_cfg": np.array([H, W, fov[0], fov[1], c["psi"][0], c["psi"][1], c["M"], c["a"],
                                            c["r_obs"], c["theta_obs"], ac, n_total, n_traced])})
    np.savez_compressed(os.path.join(HERE, "kerr_frames.npz"), **out)
    return {t: [int(np.isfinite(out[t + "_fa32"]).sum()), int(out[t + "_cfg"][-1])] for t in cfgs}


if __name__ == "__main__":
    print(rays())
    print(frames())


def rk45_kerr():
    """geodesic_tracer.trace_ray with a Kerr metric (scipy RK45 on Kerr.geodesic_equations)."""
    GT = R.geodesic_tracer
    rows, y0s, yfs = [], [], []
    ts, rs, phis = [], [], []
    for M, a, r_obs, angles_deg in [(1.0, 0.9, 50.0, [0.5, 3, 5, 5.8, 6.2, 7, 10, 30, 100]),
                                    (1.0, -0.6, 80.0, [1, 3.6, 3.9, 5, 20]),
                                    (2.0, 1.99, 40.0, [8, 14, 15, 17, 25, 60])]:
        m = MM.Kerr(M, a)
        for ad in angles_deg:
            al = float(np.radians(ad))
            sol, outcome = GT.trace_ray(m, r_obs, al)
            rows.append([M, a, r_obs, al, 1 if outcome == "escaped" else -1, sol.t.size, sol.nfev, sol.status, sol.t[-1]])
            y0s.append(np.array(m.initial_conditions(r_obs, al), dtype=np.float64))
            yfs.append(sol.y[:, -1])
            ts.append(sol.t); rs.append(sol.y[1]); phis.append(sol.y[3])
    rows = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "kerr_rk45_rays.npz"), rows=rows, state0=np.stack(y0s), y_final=np.stack(yfs),
                        traj_offsets=np.cumsum([0] + [t.size for t in ts]).astype(np.int64),
                        traj_t=np.concatenate(ts), traj_r=np.concatenate(rs), traj_phi=np.concatenate(phis))
    return rows.shape[0]


if __name__ == "__main__":
    print("kerr rk45 rays:", rk45_kerr())
