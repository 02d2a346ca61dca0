"""Analytic black-hole shadow — drop-in for the reference's ``black_hole_shadow`` module
(reference: black_hole_shadow.py:1-46).  The per-pixel classification loop runs in
lp_shadow_classify (one launch) instead of a Python double loop."""
import numpy as np

from . import _device as dev
from . import _lib
from .metrics import Schwarzschild


def pixel_to_viewing_angle(i, n, fov):
    """Viewing angle of pixel i of n along one axis (black_hole_shadow.py:7-9)."""
    i_unit = (i - n / 2) / (n / 2)
    return np.arctan(i_unit * np.tan(fov / 2))


def get_pixel_color(metric, r_obs, alpha, alpha_crit):
    """0.0 inside the shadow, 1.0 outside (black_hole_shadow.py:12-15)."""
    if alpha < alpha_crit:
        return 0.0
    return 1.0


def shadow_image(metric, width, height, fov, r_obs, *, device=False, return_count=False):
    """float64[width, height] of {0., 1.}, indexed [x, y] like the reference's ``image``
    (black_hole_shadow.py:30-37)."""
    t = dev.torch()
    alpha_crit = float(metric.alpha_crit(r_obs))
    image = t.empty((width, height), dtype=t.float64, device=dev.device())
    count = t.zeros(1, dtype=t.int64, device=image.device) if return_count else None
    _lib.ext().shadow_classify(int(width), int(height), float(fov), alpha_crit, image, count)
    img = image if device else dev.d2h(image, "shadow")
    if return_count:
        return img, int(count.item())
    return img


def main(metric=None, width=800, height=800, fov_deg=40, save=True):
    if metric is None:
        metric = Schwarzschild(M=1.0)
    fov = np.radians(fov_deg)
    r_obs = 50.0 * metric.M
    image = shadow_image(metric, width, height, fov, r_obs)
    if save:
        try:
            import matplotlib.pyplot as plt
            plt.imshow(image, cmap="gray", origin="lower")
            plt.axis("off")
            plt.savefig("black_hole_shadow.png", dpi=200, bbox_inches="tight")
            plt.close()
        except ImportError:
            from PIL import Image
            # imshow(image, origin='lower') draws image[row, col] with row 0 at the bottom
            Image.fromarray((image[::-1] * 255).astype(np.uint8)).save("black_hole_shadow.png")


if __name__ == "__main__":
    main()
