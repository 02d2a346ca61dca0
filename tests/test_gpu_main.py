"""The reference's script entry points end to end on the GPU:
image_lens.main reads image.jpg, renders, writes lensed_image.png (image_lens.py:432-514);
black_hole_shadow.main writes black_hole_shadow.png (black_hole_shadow.py:18-42)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_image(path, H, W):
    from PIL import Image
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.stack([((yy // 8 + xx // 8) & 1) * 255, (1 - ((yy // 8 + xx // 8) & 1)) * 200, xx * 255 // W], -1)
    Image.fromarray(img.astype(np.uint8)).save(path, quality=95)
    return np.asarray(Image.open(path))                       # what main() will read (JPEG is lossy)


@pytest.mark.parametrize("spin", [0.0, 0.7])
def test_image_lens_main_end_to_end(native, oracle, tmp_path, monkeypatch, capsys, spin):
    from PIL import Image
    from light_path_tracer_b200 import image_lens as il
    H, W = 90, 160
    monkeypatch.chdir(tmp_path)
    src8 = _write_image(tmp_path / "image.jpg", H, W)
    il.main(M=1.0, a=spin, r_obs_mult=100.0, psi=(0.0, 0.0), vertical_fov_deg=14.0)
    out = capsys.readouterr().out
    assert ("Kerr" if spin else "Schwarzschild") in out and "MPix/s" in out and "total" in out
    png = np.asarray(Image.open(tmp_path / "lensed_image.png"))[..., :3]
    assert png.shape == (H, W, 3)
    if spin == 0.0:
        # the reference's pipeline on the same file, through the oracle, in 8 bit: within 1/255
        src = src8.astype(np.float32) / 255.0
        vfov = np.radians(14.0)
        fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
        a = oracle.build_alpha_lookup((H, W), fov)
        fa, w, _, _ = oracle.precompute_final_alpha_lookup(a, 1.0, 100.0)
        ref = oracle.render_lensed_image(src, fa, w, fov)
        ref8 = (np.clip(ref, 0, 1) * 255).astype(np.uint8)
        bad = (np.abs(png.astype(int) - ref8.astype(int)) > 1).any(-1)
        assert bad.sum() <= 2, "%d pixels differ by more than 1/255" % int(bad.sum())
        assert (png == 0).all(-1).sum() >= 10          # the shadow is there
    else:
        assert (png == 0).all(-1).sum() >= 10


def test_black_hole_shadow_main(native, tmp_path, monkeypatch):
    from light_path_tracer_b200 import black_hole_shadow as bs
    monkeypatch.chdir(tmp_path)
    bs.main()
    assert os.path.getsize(tmp_path / "black_hole_shadow.png") > 1000
