"""Synthetic inputs of the benchmark / smoke workloads (no image files, no network).

``checkerboard`` is the deterministic source image of SURVEY.md 8(d) config 2, standing in for
the reference's ``imread('image.jpg')`` (image_lens.py:447-450): R = ((y//32 + x//32) & 1),
G = 1 - R, B = x / W as float32 RGB; the uint8 variant is floor(255 v), i.e. the 8-bit image
whose ``/ 255`` the reference would hand to its pipeline."""
import numpy as np


def checkerboard(H, W, dtype=np.float32):
    y = np.arange(H)[:, None]
    x = np.arange(W)[None, :]
    r = (((y // 32) + (x // 32)) & 1).astype(np.float32)
    img = np.empty((H, W, 3), np.float32)
    img[..., 0] = r
    img[..., 1] = 1.0 - r
    img[..., 2] = (x / W).astype(np.float32) * np.ones((H, 1), np.float32)
    if np.dtype(dtype) == np.uint8:
        return np.floor(255 * img).astype(np.uint8)
    return img.astype(dtype)
