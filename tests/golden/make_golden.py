#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (dhg14n9/Light-path-tracer) ships no tests or golden files, so
parity is anchored on what the reference itself computes here: this script
imports it through oracle/ref_harness.py (matplotlib import stub, numba cache
redirected) and stores inputs + outputs as small .npz files.  The GPU box has no
/root/reference; tests there compare against these files and against the
oracle (oracle/lp_oracle.*), which tests/test_oracle_golden.py pins to the same
files.

Host: results of numpy's SIMD arccos/arctan2 are platform dependent at the ulp
level; every array that passed through them is stored, so
each later stage can be fed the reference's own upstream array.
"""
import json
import os
import platform
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_harness  # noqa: E402

R = ref_harness.load()
MM, IL, GT, BS = R.metrics, R.image_lens, R.geodesic_tracer, R.black_hole_shadow


def separatrix(M, r_obs):
    """Adjacent doubles (last captured, first escaped) of the RK4 scheme, by bisection."""
    m = MM.Schwarzschild(M)
    ac = float(m.alpha_crit(r_obs))
    lo, hi = ac * (1 - 1e-4), ac * (1 + 1e-4)
    if not (m.trace_ray(r_obs, lo)[2] == "captured" and m.trace_ray(r_obs, hi)[2] == "escaped"):
        return None
    while True:
        mid = 0.5 * (lo + hi)
        if mid <= lo or mid >= hi:
            break
        if m.trace_ray(r_obs, mid)[2] == "captured":
            lo = mid
        else:
            hi = mid
    assert np.nextafter(lo, 1.0) == hi
    return lo, hi


def known_answers():
    rng = np.random.default_rng(20261018)
    rows = []
    for M, r_obs in [(1.0, 50.0), (1.0, 100.0), (1.0, 25.0), (1.0, 1000.0), (1.0, 15.0),
                     (2.5, 40.0), (0.5, 7.0), (1.0, 3.0)]:
        m = MM.Schwarzschild(M)
        ac = float(m.alpha_crit(r_obs))
        sep = separatrix(M, r_obs)
        lo, hi = sep if sep is not None else (0.9 * ac, 0.95 * ac)
        alphas = [0.0, 0.01, 0.5 * ac, 0.98 * ac, ac * (1 - 1e-3), ac * (1 - 1e-6), ac,
                  ac * (1 + 1e-6), ac * (1 + 1e-3), 1.02 * ac, 1.2 * ac, np.radians(8.0),
                  0.3, 0.7, 1.5, np.pi / 2, 2.0, 3.0, np.pi, -0.2, 4.0, 7.5,
                  np.nextafter(lo, 0.0), lo, hi, np.nextafter(hi, 1.0)]
        alphas += list(rng.uniform(0.0, np.pi, 40))
        alphas += list(ac * (1 + rng.normal(0, 1e-2, 40)))
        alphas += list(np.float64(np.float32(rng.uniform(0.0, 0.7, 40))))
        for a in alphas:
            a = float(a)
            so, phi, u, w = MM._schwarzschild_trace_orbit_numba(M, m.R_S, r_obs, a, 50.0, 0.05)
            sr, fa, nh = MM._schwarzschild_trace_ray_numba(M, m.R_S, r_obs, a, 50.0, 0.05)
            fa_api, nh_api, outcome = m.trace_ray(r_obs, a)
            rows.append((M, r_obs, a, so, phi, u, w, sr, fa, nh,
                         {"escaped": 1, "captured": -1, "invalid": 0}[outcome]))
    arr = np.array(rows, dtype=np.float64)
    np.savez_compressed(
        os.path.join(HERE, "binet_known_answers.npz"),
        M=arr[:, 0], r_obs=arr[:, 1], alpha=arr[:, 2],
        orbit_status=arr[:, 3].astype(np.int8), phi_f=arr[:, 4], u_f=arr[:, 5], w_f=arr[:, 6],
        ray_status=arr[:, 7].astype(np.int8), final_alpha=arr[:, 8],
        n_half=arr[:, 9].astype(np.int64), api_outcome=arr[:, 10].astype(np.int8))
    seps = {}
    for M, r_obs in [(1.0, 25.0), (1.0, 50.0), (1.0, 100.0), (1.0, 1000.0)]:
        lo, hi = separatrix(M, r_obs)
        seps["M%g_r%g" % (M, r_obs)] = [lo.hex() if hasattr(lo, "hex") else float(lo).hex(),
                                        float(hi).hex()]
    return len(rows), seps


def batch():
    rng = np.random.default_rng(7)
    out = {}
    for tag, M, r_obs in [("r100", 1.0, 100.0), ("r15", 1.0, 15.0), ("m2p5_r40", 2.5, 40.0)]:
        m = MM.Schwarzschild(M)
        ac = float(m.alpha_crit(r_obs))
        a = np.concatenate([
            rng.uniform(0.0, 0.7, 2048),
            ac * (1 + rng.normal(0, 3e-3, 1024)),
            rng.uniform(0.0, np.pi, 1024),
            np.float64(np.float32(rng.uniform(0.0, 0.5, 1024))),
            [0.0, np.pi, np.pi / 2, -0.3, 4.0],
        ])
        fa = np.full(a.size, np.nan)
        w = np.zeros(a.size, np.int64)
        m.trace_rays_batch(r_obs, a, fa, w)
        out[tag + "_alpha"] = a
        out[tag + "_fa"] = fa
        out[tag + "_w"] = w
        out[tag + "_M"] = np.float64(M)
        out[tag + "_r_obs"] = np.float64(r_obs)
    np.savez_compressed(os.path.join(HERE, "binet_batch.npz"), **out)


def checkerboard(H, W):
    y = np.arange(H)[:, None]
    x = np.arange(W)[None, :]
    r = (((y // 8) + (x // 8)) & 1).astype(np.float32)
    img = np.empty((H, W, 3), np.float32)
    img[..., 0] = r
    img[..., 1] = 1.0 - r
    img[..., 2] = (x / W).astype(np.float32) * np.ones((H, 1), np.float32)
    return img


def frames():
    cases = {
        # tag: (H, W, vfov_deg, r_obs, M, psi)
        "wide": (54, 96, 40.0, 100.0, 1.0, (0.0, 0.0)),
        "zoom": (60, 80, 9.0, 100.0, 1.0, (0.0, 0.0)),
        "offset": (48, 64, 12.0, 30.0, 1.0, (np.radians(1.5), np.radians(-2.0))),
        "odd": (37, 53, 25.0, 50.0, 2.0, (np.radians(-3.0), np.radians(4.0))),
        "bigpsi": (40, 56, 60.0, 20.0, 1.0, (np.radians(35.0), np.radians(50.0))),
    }
    out = {}
    meta = {}
    for tag, (H, W, vfov_deg, r_obs, M, psi) in cases.items():
        m = MM.Schwarzschild(M)
        vfov = np.radians(vfov_deg)
        hfov = 2 * np.arctan(np.tan(vfov / 2) * W / H)  # image_lens.py:461-463
        fov = (hfov, vfov)
        ac = m.alpha_crit(r_obs)
        alpha = IL.build_alpha_lookup((H, W), fov, psi=psi)
        fa, wnd, n_tot, n_tr = IL.precompute_final_alpha_lookup(alpha, ac, r_obs, m)
        src = checkerboard(H, W)
        src_u8 = np.floor(255 * src).astype(np.uint8)
        src_gray = src[..., 2].copy()
        src_f64 = src.astype(np.float64)
        out[tag + "_alpha32"] = alpha
        out[tag + "_fa32"] = fa
        out[tag + "_w16"] = wnd
        out[tag + "_src"] = src
        for name, s in [("rgb32", src), ("rgb8", src_u8), ("gray32", src_gray), ("rgb64", src_f64)]:
            out[tag + "_render_" + name] = IL.render_lensed_image(
                s, alpha, fa, wnd, ac, fov, False, psi=psi)
        out[tag + "_render_rgb32_loop"] = IL.render_lensed_image(
            src, alpha, fa, wnd, ac, fov, True, psi=psi)
        out[tag + "_render_rgb32_nowind"] = IL.render_lensed_image(
            src, alpha, fa, None, ac, fov, False, psi=psi)
        meta[tag] = dict(H=H, W=W, vfov_deg=vfov_deg, r_obs=r_obs, M=M, psi=list(psi),
                         hfov=float(hfov), vfov=float(vfov), alpha_crit=float(ac),
                         n_total=int(n_tot), n_traced=int(n_tr),
                         escaped=int(np.isfinite(fa).sum()),
                         winding=int((np.isfinite(fa) & (fa > np.pi / 2)).sum()))
    np.savez_compressed(os.path.join(HERE, "frames_small.npz"), **out)
    # scalar helpers (pixel_to_angles / angles_to_pixel), API parity only
    helpers = []
    rng = np.random.default_rng(3)
    H, W = 48, 64
    fov = (np.radians(50.0), np.radians(38.0))
    for psi in [(0.0, 0.0), (0.1, -0.2)]:
        for _ in range(20):
            px = (int(rng.integers(0, H)), int(rng.integers(0, W)))
            a, th = IL.pixel_to_angles(px, (H, W), fov, psi=psi)
            back = IL.angles_to_pixel((a, th), (H, W), fov, psi=psi)
            back_clip = IL.angles_to_pixel((a * 3, th), (H, W), fov, clip=True, psi=psi)
            helpers.append(dict(psi=list(psi), pixel=list(px), alpha=a, theta=th,
                                back=[int(back[0]), int(back[1])],
                                back_clip=[int(back_clip[0]), int(back_clip[1])]))
    meta["_helpers"] = dict(H=H, W=W, fov=list(fov), cases=helpers)
    return meta


def aggregates():
    """Frame-level aggregates of the default pipeline at 256x256."""
    H = W = 256
    M, r_obs = 1.0, 100.0
    m = MM.Schwarzschild(M)
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    alpha = IL.build_alpha_lookup((H, W), fov)
    fa, wnd, _, _ = IL.precompute_final_alpha_lookup(alpha, m.alpha_crit(r_obs), r_obs, m)
    a64 = alpha.ravel().astype(np.float64)
    status = np.array([MM._schwarzschild_trace_ray_numba(M, 2.0, r_obs, a, 50.0, 0.05)[0]
                       for a in a64], dtype=np.int8)
    np.savez_compressed(os.path.join(HERE, "frame_256.npz"), alpha32=alpha, fa32=fa, w16=wnd,
                        status=status.reshape(H, W))
    return dict(H=H, W=W, M=M, r_obs=r_obs, vfov_deg=40.0,
                escaped=int((status == 1).sum()), captured=int((status == -1).sum()),
                invalid=int((status == 0).sum()),
                winding=int((np.isfinite(fa) & (fa > np.pi / 2)).sum()),
                max_winding=int(wnd.max()))


def shadow():
    """black_hole_shadow.main's pixel loop (black_hole_shadow.py:30-37) at 64x48."""
    m = MM.Schwarzschild(1.0)
    width, height = 64, 48
    fov = np.radians(40)
    r_obs = 50.0 * m.M
    ac = m.alpha_crit(r_obs)
    image = np.zeros((width, height))
    for j in range(height):
        for i in range(width):
            ax = BS.pixel_to_viewing_angle(i, width, fov)
            ay = BS.pixel_to_viewing_angle(j, height, fov)
            alpha = np.arccos(np.cos(ax) * np.cos(ay))
            image[i, j] = BS.get_pixel_color(m, r_obs, alpha, ac)
    # a second, zoomed case so that the disc covers many pixels
    fov2 = np.radians(14)
    image2 = np.zeros((80, 80))
    for j in range(80):
        for i in range(80):
            ax = BS.pixel_to_viewing_angle(i, 80, fov2)
            ay = BS.pixel_to_viewing_angle(j, 80, fov2)
            image2[i, j] = BS.get_pixel_color(
                m, r_obs, np.arccos(np.cos(ax) * np.cos(ay)), ac)
    np.savez_compressed(os.path.join(HERE, "shadow.npz"),
                        image_64x48=image, fov_64x48=np.float64(fov),
                        image_80x80=image2, fov_80x80=np.float64(fov2),
                        alpha_crit=np.float64(ac), r_obs=np.float64(r_obs))


def rk45():
    """geodesic_tracer.trace_ray (scipy RK45 on the 8-D Hamiltonian, geodesic_tracer.py:22-82)."""
    out = {}
    idx = []
    ts, rs, phis = [], [], []
    k = 0
    for M, r_obs, angles_deg in [
            (1.0, 50.0, [0, 2, 4, 5, 5.5, 5.84, 5.85, 5.97, 6.5, 8, 10, 15, 45, 90, 120, 170]),
            (1.0, 100.0, [1, 2.9, 2.95, 3.0, 3.5, 8, 20, 60]),
            (2.0, 30.0, [5, 15, 19, 21, 25, 40])]:
        m = MM.Schwarzschild(M)
        for ad in angles_deg:
            a = float(np.radians(ad))
            sol, outcome = GT.trace_ray(m, r_obs, a)
            state0 = m.initial_conditions(r_obs, a)
            idx.append((M, r_obs, a, 1 if outcome == "escaped" else -1, sol.t.size, sol.nfev,
                        sol.status, sol.t[-1]))
            out["y_final_%d" % k] = sol.y[:, -1]
            out["state0_%d" % k] = np.array(state0, dtype=np.float64)
            ts.append(sol.t)
            rs.append(sol.y[1])
            phis.append(sol.y[3])
            k += 1
    idx = np.array(idx, dtype=np.float64)
    np.savez_compressed(
        os.path.join(HERE, "rk45_rays.npz"),
        M=idx[:, 0], r_obs=idx[:, 1], alpha=idx[:, 2], outcome=idx[:, 3].astype(np.int8),
        n_points=idx[:, 4].astype(np.int32), nfev=idx[:, 5].astype(np.int32),
        status=idx[:, 6].astype(np.int8), t_final=idx[:, 7],
        y_final=np.stack([out["y_final_%d" % i] for i in range(k)]),
        state0=np.stack([out["state0_%d" % i] for i in range(k)]),
        traj_offsets=np.cumsum([0] + [t.size for t in ts]).astype(np.int64),
        traj_t=np.concatenate(ts), traj_r=np.concatenate(rs), traj_phi=np.concatenate(phis))
    return k


def main():
    import numba
    import scipy
    n_ka, seps = known_answers()
    batch()
    meta = frames()
    agg = aggregates()
    shadow()
    n_rk = rk45()
    info = dict(
        generated_by="tests/golden/make_golden.py",
        reference="dhg14n9/Light-path-tracer (unmodified, /root/reference)",
        host=dict(machine=platform.machine(), python=platform.python_version(),
                  numpy=np.__version__, scipy=scipy.__version__, numba=numba.__version__,
                  libc=" ".join(platform.libc_ver())),
        known_answer_rows=n_ka, separatrix_hex=seps, frames=meta, frame_256=agg, rk45_rays=n_rk)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(info, f, indent=1, sort_keys=True)
    print(json.dumps({k: info[k] for k in ("known_answer_rows", "separatrix_hex", "frame_256",
                                           "rk45_rays")}, indent=1))


if __name__ == "__main__":
    main()
