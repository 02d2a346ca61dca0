"""GPU parity of kernel (1a), the Schwarzschild Binet RK4 tracer, through the
reference-facing API (light_path_tracer_b200.metrics -> torch extension -> C ABI).

Bar (BASELINE.json north_star): escape/capture classification bit-exact except for
rays whose impact parameter lies within 1e-9 of 3*sqrt(3)*M; final direction within
1e-9 relative; winding (integer) exact.  The checker is the oracle (oracle/lp_oracle.*,
pinned to the reference in test_oracle_golden.py) and the golden fixtures.
"""
import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9


def _metrics():
    from light_path_tracer_b200 import metrics
    return metrics


def _impact_parameter(M, r_obs, alpha):
    return r_obs * np.sin(alpha) / np.sqrt(1 - 2 * M / r_obs)


def _check_batch(M, r_obs, alpha, fa_ref, w_ref, fa, w, what):
    """Assert the north-star bar; returns (#rays exempted by the critical band, max rel err)."""
    esc_ref, esc = np.isfinite(fa_ref), np.isfinite(fa)
    b = _impact_parameter(M, r_obs, alpha)
    in_band = np.abs(b - 3 * np.sqrt(3) * M) <= 1e-9
    flips = (esc_ref != esc)
    assert not (flips & ~in_band).any(), "%s: %d classification flips outside the 1e-9 band" % (
        what, int((flips & ~in_band).sum()))
    same = esc_ref & esc
    assert np.array_equal(w[~flips], w_ref[~flips]), "%s: winding differs" % what
    err = np.abs(fa[same] - fa_ref[same]) / np.abs(fa_ref[same])
    worst = float(err.max()) if err.size else 0.0
    assert worst <= REL_TOL, "%s: final_alpha rel err %.3e > 1e-9 (at alpha=%r)" % (
        what, worst, alpha[same][np.argmax(err)])
    return int((flips & in_band).sum()), worst


def test_known_answers(native, golden):
    """Every row of the known-answer fixture (reference scalars incl. the adjacent doubles
    either side of the RK4 scheme's own separatrix) through Schwarzschild.trace_rays_batch."""
    m = _metrics()
    g = golden("binet_known_answers.npz")
    for M in np.unique(g["M"]):
        for r_obs in np.unique(g["r_obs"][g["M"] == M]):
            sel = (g["M"] == M) & (g["r_obs"] == r_obs)
            alpha = g["alpha"][sel]
            fa = np.full(alpha.size, 123.0)
            w = np.full(alpha.size, -7, dtype=np.int64)
            st = np.full(alpha.size, 9, dtype=np.int8)
            m.Schwarzschild(float(M)).trace_rays_batch(float(r_obs), alpha, fa, w, status=st)
            assert np.array_equal(st, g["ray_status"][sel]), (M, r_obs)
            fa_ref = np.where(g["ray_status"][sel] == 1, g["final_alpha"][sel], np.nan)
            _check_batch(float(M), float(r_obs), alpha, fa_ref, g["n_half"][sel], fa, w,
                         "known answers M=%g r_obs=%g" % (M, r_obs))


def test_scalar_trace_ray(native, golden):
    """Schwarzschild.trace_ray keeps the reference's (float, int, str) contract (metrics.py:817-829)."""
    m = _metrics()
    g = golden("binet_known_answers.npz")
    names = {1: "escaped", -1: "captured", 0: "invalid"}
    idx = np.where((g["M"] == 1.0) & (g["r_obs"] == 50.0))[0][:30]
    metric = m.Schwarzschild(1.0)
    for i in idx:
        fa, nh, outcome = metric.trace_ray(50.0, float(g["alpha"][i]))
        assert outcome == names[int(g["ray_status"][i])]
        assert isinstance(nh, int) and isinstance(outcome, str)
        if outcome == "escaped":
            assert abs(fa - g["final_alpha"][i]) <= REL_TOL * abs(g["final_alpha"][i])
            assert nh == g["n_half"][i]
        elif outcome == "captured":
            assert np.isnan(fa) and nh == g["n_half"][i]
        else:
            assert np.isnan(fa) and nh == 0


def test_trace_ray_honours_phi_max(native, oracle):
    """metrics.py:817-824: the scalar API passes phi_max through (h stays 0.05)."""
    m = _metrics()
    metric = m.Schwarzschild(1.0)
    for phi_max in (0.3, 1.0, 2.02, 3.14, 7.0):
        for alpha in (0.02, 0.3, 1.2):
            s, fa_o, nh_o, _ = oracle.binet_ray(1.0, 2.0, 30.0, alpha, phi_max=phi_max)
            fa, nh, outcome = metric.trace_ray(30.0, alpha, phi_max=phi_max)
            assert outcome == {1: "escaped", -1: "captured", 0: "invalid"}[s]
            assert nh == nh_o
            if s == 1:
                assert abs(fa - fa_o) <= REL_TOL * abs(fa_o)


def test_batch_golden(native, golden):
    m = _metrics()
    g = golden("binet_batch.npz")
    for tag in ("r100", "r15", "m2p5_r40"):
        M, r_obs, alpha = float(g[tag + "_M"]), float(g[tag + "_r_obs"]), g[tag + "_alpha"]
        fa = np.empty(alpha.size)
        w = np.empty(alpha.size, dtype=np.int64)
        m.Schwarzschild(M).trace_rays_batch(r_obs, alpha, fa, w)
        _check_batch(M, r_obs, alpha, g[tag + "_fa"], g[tag + "_w"], fa, w, tag)


def test_batch_views_and_empty(native, oracle):
    """The reference hands trace_rays_batch slices of larger arrays (image_lens.py:172-174)
    and may pass nothing at all; outputs are written in place, untouched elsewhere."""
    m = _metrics()
    metric = m.Schwarzschild(1.0)
    rng = np.random.default_rng(5)
    big_a = rng.uniform(0, 0.6, 1000)
    big_fa = np.full(1000, -1.0)
    big_w = np.full(1000, -1, dtype=np.int64)
    metric.trace_rays_batch(100.0, big_a[100:357], big_fa[100:357], big_w[100:357])
    assert (big_fa[:100] == -1).all() and (big_fa[357:] == -1).all() and (big_w[357:] == -1).all()
    fa_o, w_o, _, _ = oracle.trace_rays_batch(1.0, 100.0, big_a[100:357])
    _check_batch(1.0, 100.0, big_a[100:357], fa_o, w_o, big_fa[100:357], big_w[100:357], "slice")
    # strided views
    fa2 = np.full(600, -1.0)
    w2 = np.full(600, -1, dtype=np.int64)
    metric.trace_rays_batch(100.0, big_a[:600:2], fa2[::2], w2[::2])
    assert (fa2[1::2] == -1).all()
    fa_o, w_o, _, _ = oracle.trace_rays_batch(1.0, 100.0, big_a[:600:2])
    _check_batch(1.0, 100.0, big_a[:600:2], fa_o, w_o, fa2[::2], w2[::2], "strided")
    # empty
    metric.trace_rays_batch(100.0, np.empty(0), np.empty(0), np.empty(0, dtype=np.int64))


@pytest.mark.parametrize("M,r_obs", [(1.0, 100.0), (1.0, 15.0), (1.0, 1000.0), (2.5, 40.0),
                                     (1.0, 3.0), (1.0, 2.01), (1.0, 1.9)])
def test_batch_random_vs_oracle(native, oracle, M, r_obs):
    """Seeded random rays incl. a cluster around the critical angle, odd angles (negative,
    > pi, NaN) and observers at / inside the photon sphere and the horizon."""
    m = _metrics()
    rng = np.random.default_rng(int(r_obs * 10))
    ac = float(oracle.alpha_crit(M, r_obs)) if r_obs > 2 * M else 0.3
    alpha = np.concatenate([
        rng.uniform(0, np.pi, 60000), ac + rng.normal(0, 1e-3, 30000),
        ac * (1 + rng.normal(0, 1e-6, 20000)), rng.uniform(0, 2 * ac, 20000),
        [0.0, np.pi, np.pi / 2, -0.3, 4.0, -2.0, 7.0, np.nan, 1e-12, 3.14159]])
    fa = np.empty(alpha.size)
    w = np.empty(alpha.size, dtype=np.int64)
    st = np.empty(alpha.size, dtype=np.int8)
    steps = np.empty(alpha.size, dtype=np.int32)
    m.Schwarzschild(M).trace_rays_batch(r_obs, alpha, fa, w, status=st, steps=steps)
    fa_o, w_o, st_o, steps_o = oracle.trace_rays_batch(M, r_obs, alpha)
    exempt, worst = _check_batch(M, r_obs, alpha, fa_o, w_o, fa, w, "random M=%g r=%g" % (M, r_obs))
    same = st == st_o
    assert same.sum() >= alpha.size - exempt
    # identical step counts == identical discrete trajectories
    assert np.array_equal(steps[same], steps_o[same])


def test_device_resident_tensors(native, oracle):
    """CUDA tensors in -> results written into the caller's CUDA tensors, no host copies."""
    import torch
    m = _metrics()
    rng = np.random.default_rng(11)
    alpha = rng.uniform(0, 0.5, 5000)
    d_a = torch.from_numpy(alpha).cuda()
    d_fa = torch.empty(5000, dtype=torch.float64, device="cuda")
    d_w = torch.empty(5000, dtype=torch.int64, device="cuda")
    m.Schwarzschild(1.0).trace_rays_batch(100.0, d_a, d_fa, d_w)
    fa_o, w_o, _, _ = oracle.trace_rays_batch(1.0, 100.0, alpha)
    _check_batch(1.0, 100.0, alpha, fa_o, w_o, d_fa.cpu().numpy(), d_w.cpu().numpy(), "tensors")


def test_fused_mode_within_tolerance(native, oracle):
    """LP_TRACE_FUSED (FMA contraction inside the RK4 step) is not bit-identical but must
    stay inside the same north-star bar on a frame-like sample."""
    import torch
    m = _metrics()
    rng = np.random.default_rng(12)
    alpha = np.float64(np.float32(rng.uniform(0, 0.45, 200000)))
    d_a = torch.from_numpy(alpha).cuda()
    d_fa = torch.empty(alpha.size, dtype=torch.float64, device="cuda")
    d_w = torch.empty(alpha.size, dtype=torch.int64, device="cuda")
    m.Schwarzschild(1.0).trace_rays_batch(100.0, d_a, d_fa, d_w, flags=1)
    fa_o, w_o, _, _ = oracle.trace_rays_batch(1.0, 100.0, alpha)
    _check_batch(1.0, 100.0, alpha, fa_o, w_o, d_fa.cpu().numpy(), d_w.cpu().numpy(), "fused")
