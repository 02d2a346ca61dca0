"""CPU: pin the oracle (oracle/lp_oracle.*) to the fixtures generated from the
UNMODIFIED reference (tests/golden/make_golden.py).  Bit-exact everywhere."""
import numpy as np

from conftest import bits_equal


def test_binet_known_answers(oracle, golden):
    g = golden("binet_known_answers.npz")
    n = g["alpha"].size
    assert n > 1000
    for i in range(n):
        M, r_obs, a = float(g["M"][i]), float(g["r_obs"][i]), float(g["alpha"][i])
        so, phi, u, w, _ = oracle.binet_orbit(M, 2 * M, r_obs, a)
        assert so == g["orbit_status"][i]
        assert bits_equal([phi, u, w], [g["phi_f"][i], g["u_f"][i], g["w_f"][i]]), (M, r_obs, a)
        sr, fa, nh, _ = oracle.binet_ray(M, 2 * M, r_obs, a)
        assert sr == g["ray_status"][i] == g["api_outcome"][i]
        assert bits_equal([fa], [g["final_alpha"][i]])
        assert nh == g["n_half"][i]


def test_separatrix_adjacent_doubles(oracle, golden):
    meta = golden("golden_meta.json")["separatrix_hex"]
    for key, (lo, hi) in meta.items():
        M = float(key.split("_")[0][1:])
        r_obs = float(key.split("_r")[1])
        lo, hi = float.fromhex(lo), float.fromhex(hi)
        assert np.nextafter(lo, 1.0) == hi
        assert oracle.binet_ray(M, 2 * M, r_obs, lo)[0] == -1
        assert oracle.binet_ray(M, 2 * M, r_obs, hi)[0] == 1


def test_batch(oracle, golden):
    g = golden("binet_batch.npz")
    for tag in ("r100", "r15", "m2p5_r40"):
        fa, w, st, steps = oracle.trace_rays_batch(float(g[tag + "_M"]), float(g[tag + "_r_obs"]),
                                                    g[tag + "_alpha"])
        assert bits_equal(fa, g[tag + "_fa"])
        assert np.array_equal(w, g[tag + "_w"])
        assert ((st == 1) == np.isfinite(g[tag + "_fa"])).all()


def test_frames(oracle, golden):
    g = golden("frames_small.npz")
    meta = golden("golden_meta.json")["frames"]
    for tag in ("wide", "zoom", "offset", "odd", "bigpsi"):
        m = meta[tag]
        dim, fov, psi = (m["H"], m["W"]), (m["hfov"], m["vfov"]), tuple(m["psi"])
        assert np.array_equal(oracle.build_alpha_lookup(dim, fov, psi=psi), g[tag + "_alpha32"])
        fa, w, n, n2 = oracle.precompute_final_alpha_lookup(g[tag + "_alpha32"], m["M"], m["r_obs"])
        assert bits_equal(fa, g[tag + "_fa32"]) and np.array_equal(w, g[tag + "_w16"])
        assert n == n2 == m["n_total"]
        src = g[tag + "_src"]
        variants = {"rgb32": src, "rgb8": np.floor(255 * src).astype(np.uint8),
                    "gray32": src[..., 2].copy(), "rgb64": src.astype(np.float64)}
        for name, s in variants.items():
            r = oracle.render_lensed_image(s, g[tag + "_fa32"], g[tag + "_w16"], fov, False, psi)
            ref = g[tag + "_render_" + name]
            assert r.dtype == ref.dtype and np.array_equal(r, ref), (tag, name)
        assert np.array_equal(oracle.render_lensed_image(src, g[tag + "_fa32"], g[tag + "_w16"], fov, True, psi),
                              g[tag + "_render_rgb32_loop"])
        assert np.array_equal(oracle.render_lensed_image(src, g[tag + "_fa32"], None, fov, False, psi),
                              g[tag + "_render_rgb32_nowind"])


def test_frame_256_aggregates(oracle, golden):
    g = golden("frame_256.npz")
    agg = golden("golden_meta.json")["frame_256"]
    fa, w, n, _, st, steps = oracle.precompute_final_alpha_lookup(g["alpha32"], agg["M"], agg["r_obs"],
                                                                   want_status=True)
    assert bits_equal(fa, g["fa32"]) and np.array_equal(w, g["w16"]) and np.array_equal(st, g["status"])
    s = oracle.frame_stats(fa, w, st, steps)
    for k in ("escaped", "captured", "invalid", "winding", "max_winding"):
        assert s[k] == agg[k], k
    # SURVEY.md Appendix A (survey session, same reference): 256x256 default frame
    assert (s["escaped"], s["captured"], s["invalid"], s["winding"]) == (64495, 1040, 1, 384)
    assert s["sum_steps"] == 3896693 and s["max_steps"] == 200


def test_shadow(oracle, golden):
    g = golden("shadow.npz")
    ac = float(g["alpha_crit"])
    assert np.array_equal(oracle.shadow_image(64, 48, float(g["fov_64x48"]), ac), g["image_64x48"])
    assert np.array_equal(oracle.shadow_image(80, 80, float(g["fov_80x80"]), ac), g["image_80x80"])


def test_alpha_crit(oracle):
    # SURVEY.md Appendix A
    assert oracle.alpha_crit(1.0, 50.0) == 0.10200015330371326
    assert oracle.alpha_crit(1.0, 100.0) == 0.051461996376274736


def test_rk45_golden_rays(oracle, golden):
    """The scipy-RK45 restatement (oracle/lp_oracle_rk45.c) against the reference's
    geodesic_tracer.trace_ray (scipy 1.18.1): accepted points and nfev EXACT for all 30 rays
    (i.e. the whole accept/reject sequence and the event step are reproduced), event point and
    final state to 1e-11, every accepted point of the trajectory to 1e-8."""
    g = golden("rk45_rays.npz")
    off = g["traj_offsets"]
    for i in range(g["alpha"].size):
        r = oracle.rk45_trace_ray(float(g["M"][i]), float(g["r_obs"][i]), float(g["alpha"][i]))
        assert r["outcome"] == g["outcome"][i] and r["status"] == g["status"][i]
        assert r["n_points"] == g["n_points"][i] and r["nfev"] == g["nfev"][i], i
        assert np.array_equal(r["state0"], g["state0"][i])
        assert abs(r["t_final"] - g["t_final"][i]) <= 1e-11 * max(1.0, g["t_final"][i])
        assert (np.abs(r["y_final"] - g["y_final"][i]) <= 1e-11 * np.maximum(np.abs(g["y_final"][i]), 1e-3)).all()
        sl = slice(off[i], off[i + 1])
        assert np.abs(r["t"] - g["traj_t"][sl]).max() <= 1e-8
        assert (np.abs(r["y"][1] - g["traj_r"][sl]) / g["traj_r"][sl]).max() <= 1e-8
        assert np.abs(r["y"][3] - g["traj_phi"][sl]).max() <= 1e-8


def test_rk45_batch_matches_single(oracle):
    rng = np.random.default_rng(4)
    alpha = np.concatenate([rng.uniform(0, 3.1, 200), [0.0, np.nan]])
    state, lam, oc, ns, st = oracle.rk45_trace_batch(1.0, 60.0, alpha)
    for i in (0, 17, 101, 200):
        r = oracle.rk45_trace_ray(1.0, 60.0, float(alpha[i]))
        assert np.array_equal(state[i], r["y_final"]) and lam[i] == r["t_final"] and oc[i] == r["outcome"]
        assert tuple(ns[i]) == (r["n_points"], r["nfev"]) and st[i] == r["status"]
    assert oc[-1] == 0 and np.isnan(state[-1]).all() and st[-1] == -2


def test_kerr_golden_rays(oracle, golden):
    """The Kerr restatement (oracle/lp_oracle_kerr.c: DP45 on the reduced 5-D state) against the
    reference's Kerr.trace_rays_batch — bit for bit (final_alpha, winding, status) for 16 525
    rays over five (M, a, r_obs, theta_obs) configurations, axis_refine on and off."""
    g = golden("kerr_rays.npz")
    for k, row in enumerate(g["cfg"]):
        M, a, r_obs, th_obs = (float(x) for x in row[:4])
        p = "c%d_" % k
        fa, w, st, _ = oracle.kerr_trace_rays_batch(M, a, r_obs, g[p + "alpha"], g[p + "theta"], th_obs, g[p + "refine"])
        assert bits_equal(fa, g[p + "fa"]), k
        assert np.array_equal(w, g[p + "w"]) and np.array_equal(st, g[p + "status"]), k
        assert abs(oracle.kerr_r_plus(M, a) - row[5]) == 0.0


def test_rk45_kerr_golden_rays(oracle, golden):
    """scipy-RK45 restatement with Kerr.geodesic_equations (metrics.py:946-1029) against the
    reference's geodesic_tracer.trace_ray(Kerr(...)): accepted points and nfev exact, event
    point to 1e-11."""
    g = golden("kerr_rk45_rays.npz")
    for i, row in enumerate(g["rows"]):
        M, a, r_obs, al, oc, npts, nfev, status, tf = row
        r = oracle.rk45_integrate_kerr(float(M), float(a), g["state0"][i])
        assert (r["outcome"], r["n_points"], r["nfev"], r["status"]) == (int(oc), int(npts), int(nfev), int(status)), i
        assert abs(r["t_final"] - tf) <= 1e-11 * max(1.0, tf)
        assert (np.abs(r["y_final"] - g["y_final"][i]) <= 1e-11 * np.maximum(np.abs(g["y_final"][i]), 1e-3)).all()


def test_acos_restatement_accuracy(tmp_path):
    """lp_acos_unit (csrc/lp_internal.cuh) stands in for np.arccos on the device.  tools/acos_study.c runs
    the same operation sequence on the host (coefficient table tools/acos_coef.h, the MUFU seed modelled
    as a ~20-bit reciprocal square root) against long double arithmetic: worst error below 0.75 ulp, and
    the end points / NaN behave like the library's."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "acos_study")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", os.path.join(root, "tools", "acos_study.c"), "-lm", "-o", exe],
                   check=True, cwd=os.path.join(root, "tools"))
    out = subprocess.run([exe, "4000000"], check=True, capture_output=True, text=True).stdout
    worst = float(re.search(r"worst mine=([0-9.]+) ulp", out).group(1))
    assert worst < 0.75, out
    lines = out.strip().splitlines()[1:]
    for ln in lines:                                   # "x -> mine (libm ref)"
        mine, ref = re.search(r"-> (\S+) \(libm (\S+)\)", ln).groups()
        assert mine == ref, ln
    # the device table is the same table
    cuh = open(os.path.join(root, "light_path_tracer_b200", "csrc", "lp_internal.cuh")).read()
    coef = open(os.path.join(root, "tools", "acos_coef.h")).read()
    for c in re.findall(r"(-?0x1\.[0-9a-f]+p[-+]\d+)", coef):
        assert c in cuh, c
