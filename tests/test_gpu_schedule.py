"""GPU parity of the schedules added for multi-GPU frames and divergent frames: interleaved row
bands (lp_render_frame_bands), 16-byte staged stores for 8-bit tiles, the lane re-packing frame
kernel (lp_render_repack_kernel, LP_TRACE_REPACK) and the peer completion flags.  Everything here
must reproduce the one-launch, one-ray-per-thread frame BIT FOR BIT: the schedules only move work
between lanes / ranks, the per-ray arithmetic is the same device code (reference: metrics.py:49-145,
image_lens.py:133-178, :296-397)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(H, W, vfov_deg=40.0):
    from light_path_tracer_b200 import image_lens as il, _device as dev
    from light_path_tracer_b200.metrics import Schwarzschild
    vfov = np.radians(vfov_deg)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    return il, dev, Schwarzschild(1.0), fov


@pytest.mark.parametrize("dtype", ["float32", "uint8"])
def test_interleaved_bands_equal_full_frame(native, dtype):
    """Rank g's interleaved bands, as a compact tile and stored frame-addressed into a full frame
    (what PeerFrame does), for widths that are and are not multiples of 32."""
    import torch
    from light_path_tracer_b200.dist import band_layout, band_rows_of
    for H, W, G, target in ((216, 480, 4, 9), (216, 250, 2, 27), (96, 128, 3, 5)):
        il, dev, metric, fov = _setup(H, W)
        src = torch.rand(H, W, 3, device="cuda")
        u8 = dtype == "uint8"
        if u8:
            src = (src * 255).to(torch.uint8)
        full = il.render_frame(src, fov, 30.0, metric, psi=(0.03, -0.02), unit_u8=u8, flags=dev.TRACE_HYBRID)
        b = band_layout(H, G, target)
        assert b is not None
        frame = torch.zeros_like(full)
        for g in range(G):
            row0, rows, bands, frame_rows = band_rows_of(H, g, G, b)
            idx = torch.from_numpy(frame_rows).cuda()
            tile = il.render_frame(src, fov, 30.0, metric, psi=(0.03, -0.02), rows=(row0, rows), bands=bands,
                                   unit_u8=u8)
            assert torch.equal(tile, full[idx])
            # (RENDER_ROW_RUNS: the 32 x 1 warp tiles PeerFrame asks for when more than four ranks store 8-bit
            # bands into the root — same pixels, and a no-op for float32)
            for extra in (0, dev.RENDER_STAGED_STORES, dev.TRACE_REPACK, dev.TRACE_REPACK | dev.RENDER_STAGED_STORES,
                          dev.RENDER_ROW_RUNS, dev.RENDER_ROW_RUNS | dev.RENDER_STAGED_STORES):
                il.render_frame(src, fov, 30.0, metric, psi=(0.03, -0.02), rows=(row0, rows), bands=bands,
                                unit_u8=u8, out=frame[row0:],
                                flags=dev.TRACE_HYBRID | dev.RENDER_OUT_FRAME_ROWS | extra)
                assert torch.equal(frame[idx], full[idx])
        assert torch.equal(frame, full)


def test_staged_stores_u8(native):
    """16-byte staged stores of 8-bit RGB tiles: full frame, ragged tile, misaligned tile."""
    import torch
    for H, W in ((270, 480), (101, 250)):
        il, dev, metric, fov = _setup(H, W)
        src = (torch.rand(H, W, 3, device="cuda") * 255).to(torch.uint8)
        for unit in (False, True):
            ref = il.render_frame(src, fov, 100.0, metric, unit_u8=unit)
            out = il.render_frame(src, fov, 100.0, metric, unit_u8=unit,
                                  flags=dev.TRACE_HYBRID | dev.RENDER_STAGED_STORES)
            assert torch.equal(ref, out)
            out = il.render_frame(src, fov, 100.0, metric, unit_u8=unit,
                                  flags=dev.TRACE_HYBRID | dev.RENDER_STAGED_STORES | dev.RENDER_ROW_RUNS)
            assert torch.equal(ref, out)
            big = torch.full((H * W * 3 + 32,), 77, device="cuda", dtype=torch.uint8)
            for shift, rows in ((0, (7, H - 20)), (1, (0, H)), (16, (3, 50))):
                n = rows[1] * W * 3
                view = big[shift:shift + n].view(rows[1], W, 3)
                big.fill_(77)
                il.render_frame(src, fov, 100.0, metric, rows=rows, out=view, unit_u8=unit,
                                flags=dev.TRACE_HYBRID | dev.RENDER_STAGED_STORES)
                assert torch.equal(view, ref[rows[0]:rows[0] + rows[1]])
                assert (big[:shift] == 77).all() and (big[shift + n:] == 77).all()


@pytest.mark.parametrize("mode", ["hybrid", "strict"])
def test_repack_frame_bit_identical(native, mode):
    """LP_TRACE_REPACK against the one-ray-per-thread schedule: frames, lookups and statistics,
    over image layouts (float32 / uint8 / float64, 1 / 3 / 4 channels), zoomed and wide cameras,
    observers near and far, ragged tiles (chunks of 256 rays that end mid-row)."""
    import torch
    flags0 = 4 if mode == "hybrid" else 0
    cases = [(96, 128, 12.0, 100.0, (0.0, 0.0)),      # the smoke frame: lane efficiency 0.75 without re-packing
             (270, 480, 40.0, 100.0, (0.02, -0.03)),
             (301, 250, 40.0, 15.0, (0.1, 0.2)),
             (64, 100, 3.0, 30.0, (0.0, 0.01)),       # inside the shadow edge: long, near-critical rays
             (128, 128, 60.0, 2.9, (0.0, 0.0))]       # observer inside the photon sphere
    for H, W, vfov_deg, r_obs, psi in cases:
        il, dev, metric, fov = _setup(H, W, vfov_deg)
        base = torch.rand(H, W, 4, device="cuda", dtype=torch.float64)
        for src in (base[..., :3].float().contiguous(), (base[..., :3] * 255).to(torch.uint8).contiguous(),
                    base[..., 0].float().contiguous(), base.contiguous(), (base * 255).to(torch.uint8).contiguous()):
            s0, s1 = dev.new_stats(), dev.new_stats()
            ref, fa0, w0 = il.render_frame(src, fov, r_obs, metric, psi=psi, return_lookups=True, stats=s0,
                                           flags=flags0)
            out, fa1, w1 = il.render_frame(src, fov, r_obs, metric, psi=psi, return_lookups=True, stats=s1,
                                           flags=flags0 | dev.TRACE_REPACK)
            assert torch.equal(ref, out)
            assert torch.equal(fa0.view(torch.int32), fa1.view(torch.int32)) and torch.equal(w0.view(torch.int16), w1.view(torch.int16))
            a, b = dev.read_stats(s0), dev.read_stats(s1)
            for k in ("n_rays", "n_escaped", "n_captured", "n_invalid", "n_winding", "sum_steps", "max_steps",
                      "max_winding", "min_final_alpha", "max_final_alpha"):
                assert a[k] == b[k], (k, a[k], b[k])
            assert b["sum_warp_steps"] >= b["sum_steps"]
        # row tiles whose pixel count is not a multiple of the chunk
        src = base[..., :3].float().contiguous()
        ref = il.render_frame(src, fov, r_obs, metric, psi=psi, flags=flags0)
        for rows in ((0, 1), (3, H - 5), (H - 1, 1)):
            out = il.render_frame(src, fov, r_obs, metric, psi=psi, rows=rows, flags=flags0 | dev.TRACE_REPACK)
            assert torch.equal(out, ref[rows[0]:rows[0] + rows[1]])


def test_repack_1080p(native):
    """A full 1080p frame through the opt-in re-packing kernel equals the default schedule's frame."""
    import torch
    il, dev, metric, fov = _setup(1080, 1920)
    src = torch.rand(1080, 1920, 3, device="cuda")
    ref = il.render_frame(src, fov, 100.0, metric)
    assert torch.equal(ref, il.render_frame(src, fov, 100.0, metric, flags=dev.TRACE_HYBRID | dev.TRACE_REPACK))


_TILE_SCRIPT = r"""
import hashlib, sys, numpy as np, torch
sys.path.insert(0, %r)
from light_path_tracer_b200 import image_lens as il, _device as dev
from light_path_tracer_b200.metrics import Schwarzschild
metric = Schwarzschild(1.0)
h = hashlib.sha1()
g = torch.Generator().manual_seed(11)
for H, W, vdeg, r_obs, psi in ((216, 480, 40.0, 30.0, (0.03, -0.02)), (101, 250, 40.0, 100.0, (0.0, 0.0)),
                               (96, 128, 12.0, 100.0, (0.0, 0.0)), (64, 72, 20.0, 15.0, (0.1, 0.0))):
    vfov = np.radians(vdeg)
    fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
    base = torch.rand(H, W, 3, generator=g)
    for src in (base.cuda(), (base * 255).to(torch.uint8).cuda()):
        for fl in (dev.TRACE_HYBRID, dev.TRACE_HYBRID | dev.RENDER_STAGED_STORES, 0):
            st = dev.new_stats()
            out, fa, w = il.render_frame(src, fov, r_obs, metric, psi=psi, flags=fl, return_lookups=True, stats=st)
            s = dev.read_stats(st)
            h.update(out.cpu().numpy().tobytes()); h.update(fa.cpu().numpy().tobytes()); h.update(w.cpu().numpy().tobytes())
            h.update(repr((s["n_rays"], s["n_escaped"], s["n_captured"], s["sum_steps"], s["max_steps"])).encode())
        for rows in ((8, 40), (3, 17)):
            h.update(il.render_frame(src, fov, r_obs, metric, psi=psi, rows=rows).cpu().numpy().tobytes())
print(h.hexdigest())
"""


def test_warp_tile_shapes_same_frames(native):
    """The warp tile of the frame kernel (LP_RENDER_TILE_H = 1: 32x1 strip, 2: 16x2, 4: 8x4 pixels,
    the default) only changes which lane traces which pixel: frames, lookups and statistics are
    identical for every shape, with and without staged stores, float32 and 8-bit, full frames and
    row tiles (odd sizes fall back to narrower tiles)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = {}
    for th in ("1", "2", "4"):
        env = dict(os.environ, LP_RENDER_TILE_H=th)
        res = subprocess.run([sys.executable, "-c", _TILE_SCRIPT % root], env=env, capture_output=True, text=True,
                             timeout=600)
        assert res.returncode == 0, res.stderr[-2000:]
        digests[th] = res.stdout.strip().splitlines()[-1]
    assert digests["1"] == digests["2"] == digests["4"], digests


def test_peer_flags_single_device(native):
    """lp_peer_signal / lp_peer_wait on one device.  No kernel here waits for ANOTHER kernel (two
    launches on one GPU are not guaranteed to run at the same time; the cross-GPU protocol is
    exercised by tests/test_gpu_multi.py): a wait issued after its signal returns at once, a wait
    for an epoch already passed returns at once, and a wait that nobody will satisfy gives up
    after its timeout instead of hanging the GPU."""
    import time
    import torch
    from light_path_tracer_b200 import _lib
    e = _lib.ext()
    flags = torch.zeros(8, dtype=torch.int64, device="cuda")
    to = torch.zeros(1, dtype=torch.int32, device="cuda")
    e.peer_signal([flags.data_ptr(), flags.data_ptr() + 8, flags.data_ptr() + 16], 5, flags)
    e.peer_wait(flags[0:3], 3, 5, 5000, to)
    e.peer_wait(flags[0:3], 3, 4, 5000, to)           # epochs only grow: an older one is satisfied too
    torch.cuda.synchronize()
    assert int(to.item()) == 0 and flags.tolist() == [5, 5, 5, 0, 0, 0, 0, 0]
    t0 = time.perf_counter()
    e.peer_wait(flags[3:4], 1, 1, 200, to)            # nobody signals flag 3
    torch.cuda.synchronize()
    assert 0.15 < time.perf_counter() - t0 < 3.0 and int(to.item()) == 1
