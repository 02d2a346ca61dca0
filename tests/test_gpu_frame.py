"""GPU parity of the frame pipeline: build_alpha_lookup (lp_build_alpha_lookup),
precompute_final_alpha_lookup (lp_schw_trace_alpha32 / lp_schw_trace_frame), render_lensed_image
(lp_remap), the fully fused lp_render_frame, the shadow classifier and the frame reductions.

Stage-isolated: each GPU stage is fed the REFERENCE's own upstream array
from tests/golden/frames_small.npz and compared with the reference's output of that stage;
end-to-end frames are compared at the north-star pixel tolerance (1/255 in 8-bit).
"""
import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu

TAGS = ("wide", "zoom", "offset", "odd", "bigpsi")


def _il():
    from light_path_tracer_b200 import image_lens
    return image_lens


def _metric(M):
    from light_path_tracer_b200.metrics import Schwarzschild
    return Schwarzschild(M)


def _f32_ulp_diff(a, b):
    """Number of float32 steps between a and b (NaN == NaN -> 0)."""
    ai = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    d = np.abs(ai - bi)
    d[np.isnan(a) & np.isnan(b)] = 0
    return d


def _f32_boundary_distance(x64):
    """For float64 values x: distance from x to the nearest float32 ROUNDING BOUNDARY (the midpoint
    of two adjacent float32 values), in units of x.  A float64 value that two correct
    implementations compute a few fp64 ulps apart can round to different float32 neighbours only
    if it sits this close to a boundary."""
    x64 = np.asarray(x64, np.float64)
    f = x64.astype(np.float32)
    up = np.nextafter(f, np.float32(np.inf)).astype(np.float64)
    dn = np.nextafter(f, np.float32(-np.inf)).astype(np.float64)
    f = f.astype(np.float64)
    d = np.minimum(np.abs(x64 - 0.5 * (f + up)), np.abs(x64 - 0.5 * (f + dn)))
    return d / np.maximum(np.abs(x64), 1e-300)


def _attributed_frame_check(il, oracle, metric, H, W, fov, r_obs, tag):
    """Full-size frame against the oracle with EVERY difference attributed (north star: pixels within
    1/255; there is no allowance): (1) alpha table — a float32 entry differs from numpy's only where
    the fp64 arccos sits on a float32 rounding boundary (numpy's SIMD arccos and
    CUDA's differ by an fp64 ulp); (2) lookups traced from the oracle's OWN alpha table — same
    criterion for final_alpha, classification and winding identical; (3) remap fed the oracle's
    lookups — identical; (4) fused frame — every pixel whose float32 lookups equal the oracle's is
    IDENTICAL (not just within 1/255), and the others are exactly the boundary cases of (1)/(2)."""
    import torch
    from conftest import record_parity
    M = float(metric.M)
    a_ref = oracle.build_alpha_lookup((H, W), fov)
    a = il.build_alpha_lookup((H, W), fov)
    da = _f32_ulp_diff(a, a_ref)
    assert da.max() <= 1
    # (1) the fp64 alpha of the differing pixels (numpy, the reference's expression: image_lens.py:141-149)
    ys, xs = np.nonzero(da)
    if ys.size:
        fx = (W / 2) / np.tan(fov[0] / 2)
        fy = (H / 2) / np.tan(fov[1] / 2)
        xc, yc = (xs - W / 2) / fx, (ys - H / 2) / fy
        a64 = np.arccos(np.clip(1.0 / np.sqrt(1.0 + xc * xc + yc * yc), -1.0, 1.0))     # psi = (0, 0): d = (0, 0, 1)
        assert _f32_boundary_distance(a64).max() <= 4 * 2.2e-16, "alpha differs away from a float32 rounding boundary"
    fa_ref, w_ref, _, _ = oracle.precompute_final_alpha_lookup(a_ref, M, r_obs)
    d_a = torch.from_numpy(a_ref).cuda()
    n_fa = {}
    for flags in (0, 4):
        fa, w = metric.trace_alpha_table(d_a, r_obs, flags=flags)
        fa, w = fa.cpu().numpy(), w.cpu().numpy()
        assert np.array_equal(np.isnan(fa), np.isnan(fa_ref)), "classification differs (flags=%d)" % flags
        assert np.array_equal(w, w_ref), "winding differs (flags=%d)" % flags
        d = _f32_ulp_diff(fa, fa_ref)
        assert d.max() <= 1
        idx = np.nonzero(d.ravel())[0]
        n_fa[flags] = idx.size
        if idx.size:
            # (2) the reference's fp64 final_alpha of those rays sits on a float32 rounding boundary
            # to within the fp64 parity bar
            fa64, _ = oracle.trace_rays_batch(M, r_obs, a_ref.ravel()[idx].astype(np.float64))[:2]
            assert _f32_boundary_distance(fa64).max() <= 1e-9, "final_alpha differs away from a float32 rounding boundary"
    src = oracle.checkerboard(H, W)
    ref = oracle.render_lensed_image(src, fa_ref, w_ref, fov)
    out = il.render_lensed_image(torch.from_numpy(src).cuda(), None, torch.from_numpy(fa_ref).cuda(),
                                 torch.from_numpy(w_ref).cuda(), 0.0, fov).cpu().numpy()
    assert np.array_equal(out, ref), "%d remapped pixels differ" % int((out != ref).any(-1).sum())
    fused, fa_f, w_f = il.render_frame(torch.from_numpy(src).cuda(), fov, r_obs, metric, return_lookups=True)
    fused, fa_f, w_f = fused.cpu().numpy(), fa_f.cpu().numpy(), w_f.cpu().numpy()
    same_lookup = (_f32_ulp_diff(fa_f, fa_ref) == 0) & (w_f == w_ref)
    differs = (fused != ref).any(-1)
    assert not (differs & same_lookup).any(), "pixels differ although their lookups equal the reference's"
    # the pixels with a different lookup are the boundary cases established above (alpha one step
    # off, or final_alpha one step off from the same alpha): nothing else may differ
    unexplained = (da == 0) & ((_f32_ulp_diff(fa_f, fa_ref) > 1) | (w_f != w_ref))
    assert not unexplained.any()
    beyond = (np.abs(np.floor(fused * 255) - np.floor(ref * 255)) > 1).any(-1)
    assert not (beyond & same_lookup).any()
    record_parity("frame/" + tag, pixels=H * W, alpha_f32_one_step=int((da > 0).sum()),
                  final_alpha_f32_one_step_strict=n_fa[0], final_alpha_f32_one_step_hybrid=n_fa[4],
                  fused_pixels_with_a_boundary_lookup=int((~same_lookup).sum()),
                  fused_pixels_differing=int(differs.sum()), fused_pixels_beyond_1_255=int(beyond.sum()),
                  fused_pixels_beyond_1_255_with_reference_lookups=int((beyond & same_lookup).sum()))
    print("%s: alpha one float32 step off on %d pixels, final_alpha on %d (strict) / %d (hybrid); fused frame: %d "
          "pixels carry a boundary lookup, %d of them differ (%d by more than 1/255); 0 pixels differ otherwise"
          % (tag, int((da > 0).sum()), n_fa[0], n_fa[4], int((~same_lookup).sum()), int(differs.sum()),
             int(beyond.sum())))
    return int((~same_lookup).sum())



@pytest.mark.parametrize("tag", TAGS)
def test_alpha_lookup_golden(native, golden, tag):
    il = _il()
    g, m = golden("frames_small.npz"), golden("golden_meta.json")["frames"][tag]
    a = il.build_alpha_lookup((m["H"], m["W"]), (m["hfov"], m["vfov"]), psi=tuple(m["psi"]))
    ref = g[tag + "_alpha32"]
    assert a.dtype == np.float32 and a.shape == ref.shape
    # fp64 arccos then one rounding to float32: a device/host ulp difference in arccos can
    # only show where the fp64 value sits on a float32 rounding boundary
    d = _f32_ulp_diff(a, ref)
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


@pytest.mark.parametrize("decimals", [0, 2, 3, 6])
def test_alpha_lookup_decimals(native, oracle, decimals):
    """build_alpha_lookup(decimals=d) bins alpha with np.round before the float32 cast
    (image_lens.py:150-151): rint(alpha * 10^d) / 10^d on the device.  A 1-ulp arccos difference
    can move a value across a bin edge only where alpha * 10^d sits on a half-integer."""
    il = _il()
    H, W = 360, 640
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    for psi in [(0.0, 0.0), (0.21, -0.13)]:
        ref = oracle.build_alpha_lookup((H, W), fov, decimals=decimals, psi=psi)
        a = il.build_alpha_lookup((H, W), fov, decimals=decimals, psi=psi)
        assert a.dtype == np.float32 and a.shape == ref.shape
        d = _f32_ulp_diff(a, ref)
        moved = d > 1
        assert moved.sum() <= 2, "%d values landed in another bin" % int(moved.sum())
        assert np.all(np.abs(a[moved].astype(np.float64) - ref[moved]) <= 1.0001 * 10.0 ** -decimals)
        assert (d > 0).mean() < 1e-3
        # the binned table really is binned: at most ~pi * 10^d + 1 distinct values
        assert np.unique(a).size <= int(np.pi * 10 ** decimals) + 2


@pytest.mark.parametrize("tag", TAGS)
def test_trace_alpha_table_golden(native, golden, tag):
    """Reference alpha table in -> final_alpha float32 / winding uint16 out (image_lens.py:155-178)."""
    il = _il()
    g, m = golden("frames_small.npz"), golden("golden_meta.json")["frames"][tag]
    fa, w, n, n_tr = il.precompute_final_alpha_lookup(g[tag + "_alpha32"], m["alpha_crit"], m["r_obs"],
                                                      _metric(m["M"]))
    assert fa.dtype == np.float32 and w.dtype == np.uint16 and n == n_tr == m["n_total"]
    ref_fa, ref_w = g[tag + "_fa32"], g[tag + "_w16"]
    assert np.array_equal(np.isnan(fa), np.isnan(ref_fa)), "escape/capture pattern differs"
    assert np.array_equal(w, ref_w)
    # fp64 agreement to 1e-9 relative means the float32 roundings differ by at most one step
    assert _f32_ulp_diff(fa, ref_fa).max() <= 1
    assert (_f32_ulp_diff(fa, ref_fa) > 0).mean() < 1e-3


@pytest.mark.parametrize("tag", TAGS)
def test_remap_golden(native, golden, tag):
    """Reference lookups in -> rendered frame out, every source layout the reference supports;
    integer source index => exact pixel equality."""
    il = _il()
    g, m = golden("frames_small.npz"), golden("golden_meta.json")["frames"][tag]
    fov, psi = (m["hfov"], m["vfov"]), tuple(m["psi"])
    src = g[tag + "_src"]
    variants = {"rgb32": src, "rgb8": np.floor(255 * src).astype(np.uint8),
                "gray32": src[..., 2].copy(), "rgb64": src.astype(np.float64)}
    for name, s in variants.items():
        out = il.render_lensed_image(s, g[tag + "_alpha32"], g[tag + "_fa32"], g[tag + "_w16"],
                                     m["alpha_crit"], fov, False, psi=psi)
        ref = g[tag + "_render_" + name]
        assert out.dtype == ref.dtype and out.shape == ref.shape
        assert np.array_equal(out, ref), "%s/%s: %d pixels differ" % (
            tag, name, int((out != ref).reshape(ref.shape[0], ref.shape[1], -1).any(-1).sum()))
    out = il.render_lensed_image(src, g[tag + "_alpha32"], g[tag + "_fa32"], g[tag + "_w16"],
                                 m["alpha_crit"], fov, True, psi=psi)
    assert np.array_equal(out, g[tag + "_render_rgb32_loop"])
    out = il.render_lensed_image(src, g[tag + "_alpha32"], g[tag + "_fa32"], None,
                                 m["alpha_crit"], fov, False, psi=psi)
    assert np.array_equal(out, g[tag + "_render_rgb32_nowind"])


@pytest.mark.parametrize("tag", TAGS)
def test_fused_render_equals_staged(native, golden, tag):
    """lp_render_frame (one launch) == build_alpha_lookup -> trace -> remap run separately,
    bit for bit, and within 1/255 of the reference's frame end to end."""
    import torch
    il = _il()
    g, m = golden("frames_small.npz"), golden("golden_meta.json")["frames"][tag]
    dim, fov, psi = (m["H"], m["W"]), (m["hfov"], m["vfov"]), tuple(m["psi"])
    metric = _metric(m["M"])
    src = torch.from_numpy(g[tag + "_src"]).cuda()
    a = il.build_alpha_lookup(dim, fov, psi=psi, device=True)
    fa, w, _, _ = il.precompute_final_alpha_lookup(a, m["alpha_crit"], m["r_obs"], metric)
    staged = il.render_lensed_image(src, a, fa, w, m["alpha_crit"], fov, False, psi=psi)
    fused, fa2, w2 = il.render_frame(src, fov, m["r_obs"], metric, psi=psi, return_lookups=True)
    assert torch.equal(staged, fused)
    assert torch.equal(fa.view(torch.int32), fa2.view(torch.int32))
    assert torch.equal(w.view(torch.int16), w2.view(torch.int16))
    ref8 = np.floor(g[tag + "_render_rgb32"] * 255)
    out8 = np.floor(fused.cpu().numpy() * 255)
    bad = (np.abs(ref8 - out8) > 1).any(-1)
    # an alpha that differs by one float32 step can move a pixel across a checker edge;
    # none is expected at this size
    assert bad.sum() == 0, "%d pixels off by more than 1/255" % int(bad.sum())


def test_row_tiles_equal_full_frame(native):
    """Row-tile sharding (what each GPU gets) reproduces the full frame bit for bit."""
    import torch
    il = _il()
    H, W = 270, 480
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    metric = _metric(1.0)
    src = torch.rand(H, W, 3, device="cuda")
    full = il.render_frame(src, fov, 100.0, metric, psi=(0.02, -0.03))
    for parts in (2, 4, 7, 8):
        bounds = [H * k // parts for k in range(parts + 1)]
        tiles = [il.render_frame(src, fov, 100.0, metric, psi=(0.02, -0.03), rows=(bounds[k], bounds[k + 1] - bounds[k]))
                 for k in range(parts)]
        assert torch.equal(torch.cat(tiles, 0), full)


def test_frame_256_stats(native, golden):
    """Frame reductions (kernel 3) against the reference's 256x256 default frame
    (escaped 64495 / captured 1040 / invalid 1 / winding 384,
    sum of steps 3896693, max 200)."""
    import torch
    from light_path_tracer_b200 import _device as dev
    g = golden("frame_256.npz")
    metric = _metric(1.0)
    stats = dev.new_stats()
    a = torch.from_numpy(g["alpha32"]).cuda()
    st = torch.empty(a.shape, dtype=torch.int8, device="cuda")
    steps = torch.empty(a.shape, dtype=torch.int32, device="cuda")
    fa, w = metric.trace_alpha_table(a, 100.0, status=st, steps=steps, stats=stats)
    s = dev.read_stats(stats)
    assert np.array_equal(st.cpu().numpy(), g["status"])
    assert (s["n_rays"], s["n_escaped"], s["n_captured"], s["n_invalid"], s["n_winding"]) == \
        (65536, 64495, 1040, 1, 384)
    assert s["sum_steps"] == 3896693 and s["max_steps"] == 200 and s["max_winding"] == 3
    assert s["sum_steps"] == int(steps.sum().item())
    assert 0.5 < s["lane_efficiency"] <= 1.0
    fa_np = fa.cpu().numpy()
    assert abs(s["min_final_alpha"] - np.nanmin(fa_np)) <= 1e-6 * np.nanmin(fa_np) + 1e-12
    assert abs(s["max_final_alpha"] - np.nanmax(fa_np)) <= 1e-6
    # stand-alone reduction over the finished lookups gives the same counts
    stats2 = dev.new_stats()
    from light_path_tracer_b200 import _lib
    _lib.ext().stats_reduce(fa, w, st, steps, stats2)
    s2 = dev.read_stats(stats2)
    for k in ("n_rays", "n_escaped", "n_captured", "n_invalid", "n_winding", "sum_steps", "max_steps",
              "max_winding"):
        assert s[k] == s2[k], k


def test_frame_symmetry_4k(native):
    """Full-size property (no oracle needed): at psi = 0 rows y and H-y (and columns x, W-x)
    see the same alpha, so the lookups must be mirror images bit for bit; and the frame
    counts must match the analytic shadow: captured pixels == pixels with alpha < alpha_sep."""
    import torch
    from light_path_tracer_b200 import _device as dev
    il = _il()
    H, W = 2160, 3840
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    metric = _metric(1.0)
    stats = dev.new_stats()
    a = il.build_alpha_lookup((H, W), fov, device=True)
    fa, w = metric.trace_alpha_table(a, 100.0, stats=stats)
    fa_i = fa.view(torch.int32)
    assert torch.equal(fa_i[1:], fa_i[1:].flip(0))
    assert torch.equal(fa_i[:, 1:], fa_i[:, 1:].flip(1))
    assert torch.equal(w.view(torch.int16)[1:], w.view(torch.int16)[1:].flip(0))
    s = dev.read_stats(stats)
    # SURVEY.md Appendix A, 2160x3840 default frame (alpha from the survey host's numpy;
    # +-few pixels allowed for arccos ulp differences at float32 rounding boundaries)
    assert s["n_rays"] == H * W and s["n_invalid"] == 1
    assert abs(s["n_escaped"] - 8221031) <= 8 and abs(s["n_captured"] - 73368) <= 8
    assert abs(s["n_winding"] - 26872) <= 8
    assert abs(s["sum_steps"] - 459558513) <= 5000 and s["max_steps"] >= 200


def test_frame_8k_config4_properties(native):
    """BASELINE config 4 at its full size (7680x4320, 33 177 600 rays), through properties that
    need no oracle: uneven row tiles (what 2 / 4 / 8 GPUs render) concatenate to the one-launch
    frame bit for bit; the lookups are mirror images at psi = 0; every ray is classified; the
    shadow's pixel count scales with the pixel area of the 4K frame's (same field of view)."""
    import torch
    from light_path_tracer_b200 import _device as dev
    il = _il()
    H, W = 4320, 7680
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    metric = _metric(1.0)
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.rand(H, W, 3, device="cuda", generator=g)
    full = il.render_frame(src, fov, 100.0, metric)
    for parts in (2, 8):
        bounds = [H * k // parts for k in range(parts + 1)]
        out = torch.empty_like(full)
        for k in range(parts):
            il.render_frame(src, fov, 100.0, metric, rows=(bounds[k], bounds[k + 1] - bounds[k]),
                            out=out[bounds[k]:bounds[k + 1]])
        assert torch.equal(out, full), parts
    bounds = [0, 1, 1000, 1001, 3333, H]                     # ragged tiles incl. single rows
    out = torch.empty_like(full)
    for k in range(len(bounds) - 1):
        il.render_frame(src, fov, 100.0, metric, rows=(bounds[k], bounds[k + 1] - bounds[k]),
                        out=out[bounds[k]:bounds[k + 1]])
    assert torch.equal(out, full)
    del out, full
    stats = dev.new_stats()
    a = il.build_alpha_lookup((H, W), fov, device=True)
    fa, w = metric.trace_alpha_table(a, 100.0, stats=stats)
    fa_i = fa.view(torch.int32)
    assert torch.equal(fa_i[1:], fa_i[1:].flip(0)) and torch.equal(fa_i[:, 1:], fa_i[:, 1:].flip(1))
    assert torch.equal(w.view(torch.int16)[1:], w.view(torch.int16)[1:].flip(0))
    s = dev.read_stats(stats)
    assert s["n_rays"] == H * W and s["n_invalid"] == 1
    assert s["n_escaped"] + s["n_captured"] + s["n_invalid"] == H * W
    assert abs(s["n_captured"] / 4.0 - 73368) < 0.01 * 73368   # 4K frame: 73 368 captured
    assert int(torch.isnan(fa).sum().item()) == s["n_captured"] + s["n_invalid"]


def test_sweep_512_config5_properties(native):
    """BASELINE config 5 at its full size (512 frames of 1024x1024: 32 distances x 16 camera
    pitches, dist.sweep_grid), through properties that need no oracle: every ray of every frame
    is classified; for a fixed pitch the visible shadow shrinks with the distance (same centre,
    smaller disc); the lookups of pitch +p and -p are row-mirrored images bit for bit; and the
    round-robin frame shards of 8 ranks cover the sweep exactly once."""
    import torch
    from light_path_tracer_b200 import _device as dev, dist as lpdist
    il = _il()
    H = W = 1024
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    metric = _metric(1.0)
    params = lpdist.sweep_grid()
    assert len(params) == 512
    shards = [lpdist.frame_shard(len(params), r, 8) for r in range(8)]
    assert sorted(sum(shards, [])) == list(range(512)) and all(len(sh) == 64 for sh in shards)
    g = torch.Generator(device="cuda").manual_seed(9)
    pipe = il.LensPipeline(torch.rand(H, W, 3, device="cuda", generator=g), 40.0, metric)
    out = torch.empty(H, W, 3, device="cuda")
    captured = np.zeros((32, 16), dtype=np.int64)
    for k, (r_obs, psi) in enumerate(params):
        stats = dev.new_stats()
        pipe.render(r_obs, psi=psi, stats=stats, out=out)
        s = dev.read_stats(stats)
        assert s["n_rays"] == H * W and s["n_escaped"] + s["n_captured"] + s["n_invalid"] == H * W, k
        captured[k // 16, k % 16] = s["n_captured"]
    assert (np.diff(captured, axis=0) <= 0).all(), "the visible shadow must shrink with r_obs"
    assert captured[0].min() > 50 * captured[-1].max() > 0
    for r_obs, (p, _) in (params[3], params[16 * 13 + 1], params[16 * 31 + 6]):
        fa_p, w_p = metric.trace_alpha_table(il.build_alpha_lookup((H, W), fov, psi=(p, 0.0), device=True), r_obs)
        fa_m, w_m = metric.trace_alpha_table(il.build_alpha_lookup((H, W), fov, psi=(-p, 0.0), device=True), r_obs)
        assert torch.equal(fa_p.view(torch.int32)[1:], fa_m.view(torch.int32)[1:].flip(0))
        assert torch.equal(w_p.view(torch.int16)[1:], w_m.view(torch.int16)[1:].flip(0))


@pytest.mark.parametrize("H,W,r_obs,psi", [(2160, 3840, 100.0, (0.0, 0.0)), (2160, 3840, 15.0, (0.05, -0.1)),
                                           (1080, 1920, 1000.0, (0.0, 0.0))])
def test_hybrid_frame_equals_strict_frame(native, H, W, r_obs, psi):
    """Full-size check of the image pipeline's default arithmetic (LP_TRACE_HYBRID) against
    LP_TRACE_STRICT (the mode that is bit-identical to the reference's trajectories): same
    escape/capture classification and winding for every pixel, float32 final_alpha equal up
    to a handful of one-step float32 rounding flips, rendered pixels equal wherever
    final_alpha is."""
    import torch
    il = _il()
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    metric = _metric(1.0)
    src = torch.rand(H, W, 3, device="cuda")
    out_s, fa_s, w_s = il.render_frame(src, fov, r_obs, metric, psi=psi, flags=0, return_lookups=True)
    out_h, fa_h, w_h = il.render_frame(src, fov, r_obs, metric, psi=psi, flags=4, return_lookups=True)
    assert torch.equal(torch.isnan(fa_s), torch.isnan(fa_h))
    assert torch.equal(w_s.view(torch.int16), w_h.view(torch.int16))
    d = (fa_s.view(torch.int32) - fa_h.view(torch.int32)).abs()
    d[torch.isnan(fa_s)] = 0
    assert int(d.max()) <= 1 and int((d > 0).sum()) <= 16
    same = d == 0
    assert torch.equal(out_s[same], out_h[same])


def test_shadow_golden(native, golden):
    from light_path_tracer_b200 import black_hole_shadow as bs
    g = golden("shadow.npz")
    metric = _metric(1.0)
    img, n_dark = bs.shadow_image(metric, 64, 48, float(g["fov_64x48"]), float(g["r_obs"]), return_count=True)
    assert img.dtype == np.float64 and img.shape == (64, 48)
    assert np.array_equal(img, g["image_64x48"]) and n_dark == int((g["image_64x48"] == 0).sum())
    img = bs.shadow_image(metric, 80, 80, float(g["fov_80x80"]), float(g["r_obs"]))
    assert np.array_equal(img, g["image_80x80"])


def test_shadow_256_vs_oracle(native, oracle):
    """BASELINE.json config 1: the black_hole_shadow computation at 256x256 (fov 40 deg, r_obs 50 M)."""
    from light_path_tracer_b200 import black_hole_shadow as bs
    metric = _metric(1.0)
    fov = np.radians(40)
    img = bs.shadow_image(metric, 256, 256, fov, 50.0)
    ref = oracle.shadow_image(256, 256, fov, float(metric.alpha_crit(50.0)))
    assert np.array_equal(img, ref)
    assert set(np.unique(img)) <= {0.0, 1.0}


def test_frame_1080p_vs_oracle(native, oracle):
    """BASELINE.json config 2 (1920x1080 checkerboard) against the oracle, stage by stage, every
    difference attributed; plus the reference-facing numpy calls on the oracle's tables."""
    il = _il()
    H, W = 1080, 1920
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    metric = _metric(1.0)
    n = _attributed_frame_check(il, oracle, metric, H, W, fov, 100.0, "1080p_r100")
    assert n <= 16
    a_ref = oracle.build_alpha_lookup((H, W), fov)
    fa_ref, w_ref, _, _ = oracle.precompute_final_alpha_lookup(a_ref, 1.0, 100.0)
    fa, w, _, _ = il.precompute_final_alpha_lookup(a_ref, metric.alpha_crit(100.0), 100.0, metric)
    assert np.array_equal(np.isnan(fa), np.isnan(fa_ref)) and np.array_equal(w, w_ref)
    src = oracle.checkerboard(H, W)
    assert np.array_equal(il.render_lensed_image(src, a_ref, fa_ref, w_ref, 0.0, fov),
                          oracle.render_lensed_image(src, fa_ref, w_ref, fov))


def test_staged_stores_same_frame(native):
    """LP_RENDER_STAGED_STORES (16-byte pixel stores through shared memory, used for peer-memory
    tiles) writes exactly the frame the plain stores write — full frame, a row tile whose pixel
    count is not a multiple of 32, and a misaligned tile (falls back to plain stores)."""
    import torch
    from light_path_tracer_b200 import _device as dev
    il = _il()
    metric = _metric(1.0)
    for H, W in ((270, 480), (101, 250)):
        vfov = np.radians(40.0)
        fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
        src = torch.rand(H, W, 3, device="cuda")
        ref = il.render_frame(src, fov, 100.0, metric)
        out = il.render_frame(src, fov, 100.0, metric, flags=dev.TRACE_HYBRID | dev.RENDER_STAGED_STORES)
        assert torch.equal(ref, out)
        big = torch.full((H * W * 3 + 8,), -1.0, device="cuda")
        for shift, rows in ((0, (7, H - 20)), (1, (0, H)), (4, (3, 50))):
            n = rows[1] * W * 3
            view = big[shift:shift + n].view(rows[1], W, 3)
            big.fill_(-1.0)
            il.render_frame(src, fov, 100.0, metric, rows=rows, out=view,
                            flags=dev.TRACE_HYBRID | dev.RENDER_STAGED_STORES)
            assert torch.equal(view, ref[rows[0]:rows[0] + rows[1]])
            assert (big[:shift] == -1).all() and (big[shift + n:] == -1).all()


def test_host_frame_pipeline(native, oracle):
    """HostFramePipeline (pinned host source -> H2D -> fused kernel -> D2H, double-buffered over
    two streams): every returned frame equals the device-resident render of the same inputs,
    including when the source changes from frame to frame and for row tiles."""
    import torch
    il = _il()
    metric = _metric(1.0)
    H, W = 240, 320
    pipe = il.HostFramePipeline((H, W, 3), torch.float32, 40.0, metric, depth=2)
    g = torch.Generator().manual_seed(9)
    sources = [torch.rand(H, W, 3, generator=g).pin_memory() for _ in range(5)]
    params = [(100.0, (0.0, 0.0)), (30.0, (0.1, 0.0)), (15.0, (0.0, -0.2)), (300.0, (0.05, 0.05)), (100.0, (0.0, 0.0))]
    outs = [pipe.submit(s, r, psi=p) for s, (r, p) in zip(sources, params)]
    tile = pipe.submit(sources[0], 100.0, rows=(60, 100))
    pipe.synchronize()
    for s, (r, p), o in zip(sources, params, outs):
        ref = il.render_frame(s.cuda(), pipe.fov, r, metric, psi=p).cpu()
        assert torch.equal(o, ref)
    ref = il.render_frame(sources[0].cuda(), pipe.fov, 100.0, metric, rows=(60, 100)).cpu()
    assert torch.equal(tile, ref)


@pytest.mark.parametrize("tag", ("wide", "offset"))
def test_unit_u8_boundary(native, golden, oracle, tag):
    """LP_DTYPE_U8_UNIT: the 8-bit boundary of image_lens.main (imread uint8 -> float32/255 ->
    pipeline -> imsave 8 bit).  The byte frame must equal trunc(255 * v) of the reference's
    float32 pipeline on source/255 — exactly, for the remap alone (reference lookups in) and
    for the fused kernel wherever its float32 final_alpha equals the reference's."""
    import torch
    il = _il()
    g = golden("frames_small.npz")
    meta = golden("golden_meta.json")["frames"][tag]
    H, W = meta["H"], meta["W"]
    fov, psi = (meta["hfov"], meta["vfov"]), tuple(meta["psi"])
    rng = np.random.default_rng(17)
    src8 = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    src_f = src8.astype(np.float32) / 255.0                       # image_lens.py:450
    fa, w = g[tag + "_fa32"], g[tag + "_w16"]
    ref = (oracle.render_lensed_image(src_f, fa, w, fov, False, psi) * 255).astype(np.uint8)
    out = il.render_lensed_image(src8, None, fa, w, 0.0, fov, False, psi, unit_u8=True)
    assert out.dtype == np.uint8 and np.array_equal(out, ref)
    # plain uint8 keeps the reference's own (0/1 colour) behaviour
    ref_plain = oracle.render_lensed_image(src8, fa, w, fov, False, psi)
    assert np.array_equal(il.render_lensed_image(src8, None, fa, w, 0.0, fov, False, psi), ref_plain)
    # fused
    metric = _metric(float(meta["M"]))
    frame, fa_g, w_g = il.render_frame(torch.from_numpy(src8).cuda(), fov, float(meta["r_obs"]), metric, psi=psi,
                                       unit_u8=True, return_lookups=True)
    fa_g, w_g = fa_g.cpu().numpy(), w_g.cpu().numpy()
    ref2 = (oracle.render_lensed_image(src_f, fa_g, w_g, fov, False, psi) * 255).astype(np.uint8)
    assert np.array_equal(frame.cpu().numpy(), ref2)


def test_frame_4k_vs_oracle(native, oracle):
    """The bench workload itself (3840x2160, r_obs = 100 M, 40 deg, float32 RGB checkerboard)
    against the oracle at FULL size, every difference attributed (see _attributed_frame_check):
    no pixel differs unless one of its float32 lookups sits on a rounding boundary."""
    il = _il()
    H, W = 2160, 3840
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    n = _attributed_frame_check(il, oracle, _metric(1.0), H, W, fov, 100.0, "4k_r100")
    assert n <= 64            # 8.3e6 pixels x P(fp64 value within a few ulp of a float32 boundary) ~ 1e-6


def test_capture_sweep_graph(native):
    """LensPipeline.capture_sweep: the CUDA-graph replay of a sweep writes the frames the
    one-by-one renders write (and again after the buffers were cleared)."""
    import torch
    il = _il()
    metric = _metric(1.0)
    src = torch.rand(96, 128, 3, device="cuda")
    pipe = il.LensPipeline(src, 30.0, metric)
    params = [(15.0, (0.0, 0.0)), (40.0, (0.05, 0.0)), (100.0, (0.0, -0.1)), (600.0, (-0.1, 0.1))]
    ref = torch.stack([pipe.render(r, psi=p) for r, p in params])
    out = torch.empty_like(ref)
    graph = pipe.capture_sweep(params, out)
    out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)


def test_bilinear_sampling_extension(native, oracle, golden):
    """LP_SAMPLE_BILINEAR (opt-in; the reference samples nearest): same in/out-of-frame decision as
    nearest (taken from the rounded coordinate), 4-tap blend around the continuous source
    coordinate with clamped taps.  Checked against a numpy evaluation of the same definition built
    on the oracle's geometry; nearest pixels that are magenta / winding / black stay so."""
    il = _il()
    g = golden("frames_small.npz")
    meta = golden("golden_meta.json")["frames"]["offset"]
    H, W = meta["H"], meta["W"]
    fov, psi = (meta["hfov"], meta["vfov"]), tuple(meta["psi"])
    fa, w = g["offset_fa32"], g["offset_w16"]
    yy, xx = np.mgrid[0:H, 0:W]
    src = np.stack([np.sin(xx / 7.0) * 0.5 + 0.5, np.cos(yy / 5.0) * 0.5 + 0.5, (xx + yy) / (H + W)], -1).astype(np.float32)
    near, sy_map, sx_map = oracle.render_lensed_image(src, fa, w, fov, False, psi, return_index=True)
    out = il.render_lensed_image(src, None, fa, w, 0.0, fov, False, psi, sampling=il.SAMPLE_BILINEAR)
    sampled = sy_map >= 0
    assert np.array_equal(out[~sampled], near[~sampled])            # captured / winding / magenta untouched
    # continuous source coordinates, as the oracle forms them (image_lens.py:310-375)
    fx, fy = oracle.focal((H, W), fov)
    d, e_x, e_y, _ = oracle.psi_frame(psi)
    xc = (np.arange(W) - W / 2) / fx
    yc = (np.arange(H) - H / 2) / fy
    norm = np.sqrt(1.0 + xc[None, :] ** 2 + yc[:, None] ** 2)
    vx, vy, vz = xc[None, :] / norm, yc[:, None] / norm, 1.0 / norm
    th = np.arctan2(vx * e_x[0] + vy * e_x[1] + vz * e_x[2], vx * e_y[0] + vy * e_y[1] + vz * e_y[2])
    f = fa.astype(np.float64)
    with np.errstate(invalid="ignore"):
        s = [np.cos(f) * d[k] + np.sin(f) * (np.sin(th) * e_x[k] + np.cos(th) * e_y[k]) for k in range(3)]
        px = s[0] / s[2] * fx + W / 2
        py = s[1] / s[2] * fy + H / 2
    x0 = np.floor(px[sampled]); y0 = np.floor(py[sampled])
    tx = (px[sampled] - x0)[:, None]; ty = (py[sampled] - y0)[:, None]
    cl = lambda v, n: np.clip(v.astype(np.int64), 0, n - 1)
    X0, X1, Y0, Y1 = cl(x0, W), cl(x0 + 1, W), cl(y0, H), cl(y0 + 1, H)
    srcd = src.astype(np.float64)
    ref = ((1 - tx) * (1 - ty) * srcd[Y0, X0] + tx * (1 - ty) * srcd[Y0, X1]
           + (1 - tx) * ty * srcd[Y1, X0] + tx * ty * srcd[Y1, X1])
    assert np.abs(out[sampled] - ref).max() <= 2e-6
    assert np.abs(out[sampled] - near[sampled]).max() > 1e-3        # it really interpolates


def test_unique_alpha_mode_bit_identical(native):
    """precompute_final_alpha_lookup(..., unique=True): every distinct float32 alpha traced once and
    scattered back — the lookups are bit-identical to the per-pixel path (on-axis and off-axis
    cameras, binned tables, numpy and device input, a reused index across observer distances)."""
    import torch
    il = _il()
    metric = _metric(1.0)
    H, W = 540, 960
    vfov = np.radians(40.0)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    for psi, decimals in (((0.0, 0.0), None), ((0.07, -0.04), None), ((0.0, 0.0), 3)):
        a = il.build_alpha_lookup((H, W), fov, decimals=decimals, psi=psi, device=True)
        fa0, w0, n0, t0 = il.precompute_final_alpha_lookup(a, 0.0, 100.0, metric)
        fa1, w1, n1, t1 = il.precompute_final_alpha_lookup(a, 0.0, 100.0, metric, unique=True)
        assert torch.equal(fa0.view(torch.int32), fa1.view(torch.int32)) and torch.equal(w0.view(torch.int16), w1.view(torch.int16))
        assert n0 == n1 == H * W and t0 == H * W and 0 < t1 < H * W
        print("psi=%r decimals=%r: %d distinct alphas for %d pixels (%.1fx fewer rays)" % (psi, decimals, t1, H * W, H * W / t1))
        idx = il.UniqueAlphaIndex(a)
        assert idx.n_unique == t1
        for r_obs in (15.0, 400.0):
            fa0, w0, _, _ = il.precompute_final_alpha_lookup(a, 0.0, r_obs, metric)
            fa1, w1, _, _ = il.precompute_final_alpha_lookup(a, 0.0, r_obs, metric, unique=idx)
            assert torch.equal(fa0.view(torch.int32), fa1.view(torch.int32)) and torch.equal(w0.view(torch.int16), w1.view(torch.int16))
    a_np = il.build_alpha_lookup((H, W), fov)
    fa0, w0, _, _ = il.precompute_final_alpha_lookup(a_np, 0.0, 100.0, metric)
    fa1, w1, _, t1 = il.precompute_final_alpha_lookup(a_np, 0.0, 100.0, metric, unique=True)
    assert bits_equal(fa0, fa1) and np.array_equal(w0, w1) and t1 < H * W
