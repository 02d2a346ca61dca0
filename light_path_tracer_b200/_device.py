"""Device plumbing shared by the reference-facing modules: host<->device staging through
pinned memory, camera packing, frame-statistics decoding.  PyTorch is used for memory
and streams only; all compute goes through the C ABI (``_lib.ext()``)."""
import ctypes

import numpy as np

from . import _lib

PHI_MAX = 50.0   # metrics.py:833
H_MAX = 0.05     # metrics.py:833 (and :824 for the scalar path)

TRACE_STRICT = 0
TRACE_FUSED = 1
RENDER_STAGED_STORES = 8   # lp_render_frame: 16-byte staged pixel stores (peer-memory tiles)
TRACE_HYBRID = 4   # FMA loop + strict re-trace of rays longer than 240 steps (lightpath.h)

_pinned = {}


def torch():
    return _lib.require_cuda()


def device():
    t = torch()
    return t.device("cuda", t.cuda.current_device())


def pinned_buffer(tag, nbytes):
    """Grow-only pinned staging buffer (uint8 tensor) per (tag, device)."""
    t = torch()
    key = (tag, t.cuda.current_device())
    buf = _pinned.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = t.empty(max(int(nbytes), 1 << 16), dtype=t.uint8, pin_memory=True)
        _pinned[key] = buf
    return buf


def h2d(arr, tag="h2d"):
    """numpy array (any layout) -> contiguous CUDA tensor of the same dtype/shape."""
    t = torch()
    arr = np.asarray(arr)
    if arr.size == 0:
        return t.empty(arr.shape, dtype=_torch_dtype(arr.dtype), device=device())
    stage = pinned_buffer(tag, arr.nbytes)[:arr.nbytes].view(_torch_dtype(arr.dtype)).view(arr.shape)
    np.copyto(stage.numpy(), arr)
    return stage.to(device(), non_blocking=True)


def d2h_into(tensor, out, tag="d2h"):
    """CUDA tensor -> caller-owned numpy array (possibly a strided view), in place."""
    t = torch()
    if tensor.numel() == 0:
        return
    nbytes = tensor.numel() * tensor.element_size()
    stage = pinned_buffer(tag, nbytes)[:nbytes].view(tensor.dtype).view(tensor.shape)
    stage.copy_(tensor, non_blocking=True)
    t.cuda.current_stream().synchronize()
    np.copyto(out, stage.numpy().reshape(out.shape), casting="same_kind")


def d2h(tensor, tag="d2h"):
    out = np.empty(tuple(tensor.shape), dtype=_numpy_dtype(tensor.dtype))
    d2h_into(tensor, out, tag)
    return out


def _torch_dtype(dt):
    t = torch()
    return {np.dtype(np.float64): t.float64, np.dtype(np.float32): t.float32,
            np.dtype(np.int64): t.int64, np.dtype(np.int32): t.int32, np.dtype(np.int8): t.int8,
            np.dtype(np.uint8): t.uint8, np.dtype(np.uint16): t.uint16}[np.dtype(dt)]


def _numpy_dtype(dt):
    t = torch()
    return {t.float64: np.float64, t.float32: np.float32, t.int64: np.int64, t.int32: np.int32,
            t.int8: np.int8, t.uint8: np.uint8, t.uint16: np.uint16}[dt]


def camera_vector(image_dimension, fov, psi, frame):
    """(H, W, fx, fy, d, e_x, e_y) as the flat list the extension expects.  `frame` is
    image_lens._psi_frame (host numpy, exactly the reference's arithmetic)."""
    height, width = image_dimension
    hfov, vfov = fov
    fx = (width / 2) / np.tan(hfov / 2)      # image_lens.py:138
    fy = (height / 2) / np.tan(vfov / 2)     # image_lens.py:139
    d, e_x, e_y, _ = frame(psi)
    return [float(height), float(width), float(fx), float(fy)] + [float(v) for v in d] + \
        [float(v) for v in e_x] + [float(v) for v in e_y]


def new_stats():
    """Zeroed lp_frame_stats in device memory (as an int64 tensor)."""
    t = torch()
    e = _lib.ext()
    s = t.zeros(int(e.STATS_WORDS), dtype=t.int64, device=device())
    e.stats_reset(s)
    return s


def read_stats(stats_tensor):
    """Decode a device lp_frame_stats into a dict (one small D2H copy)."""
    raw = stats_tensor.cpu().numpy().tobytes()
    s = _lib.lp_frame_stats.from_buffer_copy(raw)
    out = {name: getattr(s, name) for name, _ in _lib.lp_frame_stats._fields_}
    out["lane_efficiency"] = (out["sum_steps"] / out["sum_warp_steps"]) if out["sum_warp_steps"] else 1.0
    return out


def sizeof_stats():
    return ctypes.sizeof(_lib.lp_frame_stats)
