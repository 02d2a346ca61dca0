"""Lensed-image pipeline — drop-in for the reference's ``image_lens`` module
(reference: image_lens.py:21-535) with the per-pixel work on the GPU.

Every set of coordinates is written as (y, x); FOV pairs are (horizontal, vertical).

Reference-facing functions keep their names, positional order, defaults, return
shapes and dtypes:

    build_alpha_lookup            -> float32[H, W]                    (image_lens.py:133-152)
    precompute_final_alpha_lookup -> (float32[H,W], uint16[H,W], n, n) (image_lens.py:155-178)
    render_lensed_image           -> array like source_image          (image_lens.py:296-397)

They accept and return numpy arrays like the reference (host<->device staging inside),
or CUDA tensors in / CUDA tensors out when handed tensors.  ``render_frame`` and
``LensPipeline`` are the device-resident, fully fused form (pixel -> alpha -> geodesic ->
remap in one launch); they are additions, not part of the reference API.
"""
from time import perf_counter

import numpy as np

from . import _device as dev
from . import _lib
from .metrics import Schwarzschild, Kerr, _is_tensor

WINDING_DTYPE = np.uint16
WINDING_MAX = np.iinfo(WINDING_DTYPE).max
Y_AXIS_REFINE_FRAC = 0.07

SAMPLE_NEAREST = 0
SAMPLE_BILINEAR = 1

WINDING_COLORS = np.array([
    [0.0, 0.2, 1.0],   # blue
    [0.0, 0.7, 1.0],   # sky blue
    [0.0, 1.0, 0.4],   # green
    [1.0, 1.0, 0.0],   # yellow
    [1.0, 0.4, 0.0],   # orange
], dtype=np.float32)


# ============================================================================
# Camera geometry (host, three-vectors; reference: image_lens.py:21-126)
# ============================================================================

def _psi_to_bh_direction(psi):
    """psi=(pitch_up, yaw_right) [rad] -> unit vector towards the BH in camera axes
    (+x right, +y down, +z forward)."""
    psi_y, psi_x = psi
    cp = np.cos(psi_y)
    return np.array([np.sin(psi_x) * cp, -np.sin(psi_y), np.cos(psi_x) * cp], dtype=np.float64)


def _psi_frame(psi):
    """(d, e_x, e_y, in_front): BH direction and the tangent basis around it; e_x/e_y line
    up with the image axes at psi = 0."""
    d = _psi_to_bh_direction(psi)
    in_front = bool(d[2] > 1e-12)
    axes = (np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0]))

    e_x = axes[0] - np.dot(axes[0], d) * d
    n = np.linalg.norm(e_x)
    if n < 1e-12:
        e_x = axes[1] - np.dot(axes[1], d) * d
        n = np.linalg.norm(e_x)
    e_x /= max(n, 1e-12)

    e_y = axes[1] - np.dot(axes[1], d) * d - np.dot(axes[1], e_x) * e_x
    n = np.linalg.norm(e_y)
    if n < 1e-12:
        e_y = np.cross(d, e_x)
        n = np.linalg.norm(e_y)
    e_y /= max(n, 1e-12)
    return d, e_x, e_y, in_front


def _psi_to_cam_projection(psi):
    """BH direction on the pinhole plane: (y_cam, x_cam, in_front)."""
    d, _, _, in_front = _psi_frame(psi)
    if not in_front:
        return (np.nan, np.nan, False)
    return (float(d[1] / d[2]), float(d[0] / d[2]), True)


def _focal(image_dimension, fov):
    height, width = image_dimension
    horizontal_fov, vertical_fov = fov
    return (width / 2) / np.tan(horizontal_fov / 2), (height / 2) / np.tan(vertical_fov / 2)


def pixel_to_angles(pixel, image_dimension, fov, psi=(0.0, 0.0)):
    height, width = image_dimension
    fx, fy = _focal(image_dimension, fov)
    d, e_x, e_y, _ = _psi_frame(psi)
    ray = np.array([(pixel[1] - width / 2) / fx, (pixel[0] - height / 2) / fy, 1.0],
                   dtype=np.float64)
    ray /= np.linalg.norm(ray)
    alpha = float(np.arccos(np.clip(np.dot(ray, d), -1.0, 1.0)))
    theta = float(np.arctan2(np.dot(ray, e_x), np.dot(ray, e_y)))
    return (alpha, theta)


def angles_to_pixel(angles, image_dimension, fov, clip=False, psi=(0.0, 0.0)):
    alpha, theta = angles
    height, width = image_dimension
    fx, fy = _focal(image_dimension, fov)
    d, e_x, e_y, _ = _psi_frame(psi)
    ray = (np.cos(alpha) * d + np.sin(alpha) * (np.sin(theta) * e_x + np.cos(theta) * e_y))
    if ray[2] <= 1e-12:
        return (0, 0) if clip else (-1, -1)
    px = int(np.rint(ray[0] / ray[2] * fx + width / 2))
    py = int(np.rint(ray[1] / ray[2] * fy + height / 2))
    if clip:
        px = int(np.clip(px, 0, width - 1))
        py = int(np.clip(py, 0, height - 1))
    return (py, px)


# ============================================================================
# Alpha lookup (1-D, spherically symmetric metrics)
# ============================================================================

def build_alpha_lookup(image_dimension, fov, decimals=None, psi=(0.0, 0.0), *, device=False):
    """Per-pixel viewing angle, float32[H, W] (image_lens.py:133-152), computed by
    lp_build_alpha_lookup.  ``device=True`` returns the CUDA tensor instead of numpy."""
    height, width = image_dimension
    t = dev.torch()
    if decimals is not None and decimals < 0:
        raise NotImplementedError("negative `decimals` is not supported on the GPU path")
    out = t.empty((height, width), dtype=t.float32, device=dev.device())
    cam = dev.camera_vector(image_dimension, fov, psi, _psi_frame)
    _lib.ext().build_alpha_lookup(cam, 0, int(height), -1 if decimals is None else int(decimals), out)
    return out if device else dev.d2h(out, "alpha")


class UniqueAlphaIndex:
    """The distinct float32 viewing angles of an alpha table and, per pixel, which of them it holds
    (``np.unique(alpha_lookup, return_inverse=True)``, the idea of the reference's older driver,
    debugging_image_lense.py:634-640).  It depends on the camera only, so a sweep over observer
    distances builds it once: ``idx = UniqueAlphaIndex(alpha_lookup)`` and then
    ``precompute_final_alpha_lookup(alpha_lookup, ac, r_obs, metric, unique=idx)`` per frame."""

    def __init__(self, alpha_lookup):
        t = dev.torch()
        a32 = alpha_lookup if _is_tensor(alpha_lookup) else dev.h2d(np.asarray(alpha_lookup, dtype=np.float32), "alpha")
        if a32.dtype != t.float32:
            raise TypeError("UniqueAlphaIndex takes the float32 alpha table")
        self.shape = tuple(a32.shape)
        # bit patterns, so that -0.0 / 0.0 and NaN payloads stay distinct rays like in the per-pixel path
        bits = a32.contiguous().view(t.int32).reshape(-1)
        u, inv = t.unique(bits, sorted=True, return_inverse=True)
        self.alpha = u.view(t.float32)
        self.inverse = inv
        self.n_unique = int(u.numel())


def precompute_final_alpha_lookup(alpha_lookup, alpha_crit, r_obs, metric, *, unique=False):
    """One ray per pixel (image_lens.py:155-178) -> (final_alpha float32[H,W],
    winding uint16[H,W], total_rays, traced_rays).  ``alpha_crit`` is unused, as in the
    reference: shadow pixels are integrated too.

    ``unique=True`` (or a ``UniqueAlphaIndex``; opt-in, Schwarzschild only): for a spherically
    symmetric metric final_alpha and the winding number are functions of the viewing angle alone,
    so every DISTINCT float32 alpha is traced once and the result is scattered back to the pixels
    that hold it.  No interpolation: the lookups are bit-identical to the per-pixel path (pixel
    symmetry alone makes most alphas of an on-axis frame occur 4-8 times; with
    ``build_alpha_lookup(decimals=...)`` bins, thousands of times); ``traced_rays`` reports the
    number of distinct angles.  Off by default because finding the distinct values (a device sort)
    costs about as much as tracing a 4K frame; it pays when the index is reused across frames.

    A ``Schwarzschild`` metric takes the single-launch GPU path on the float32 table; any
    other ``Metric`` goes through its own ``trace_rays_batch`` in 50 000-ray chunks
    exactly as the reference drives it."""
    tensor_in = _is_tensor(alpha_lookup)
    n = int(alpha_lookup.numel() if tensor_in else alpha_lookup.size)
    shape = tuple(alpha_lookup.shape)
    if n == 0:
        if tensor_in:
            t = dev.torch()
            return (t.full(shape, float("nan"), dtype=t.float32, device=alpha_lookup.device),
                    t.zeros(shape, dtype=t.uint16, device=alpha_lookup.device), n, 0)
        return (np.full(shape, np.nan, dtype=np.float32), np.zeros(shape, dtype=WINDING_DTYPE), n, 0)

    if isinstance(metric, Schwarzschild) and type(metric).trace_rays_batch is Schwarzschild.trace_rays_batch:
        t = dev.torch()
        is_f32 = (alpha_lookup.dtype == t.float32) if tensor_in else (np.asarray(alpha_lookup).dtype == np.float32)
        if unique is not False and unique is not None:
            if not is_f32:
                raise TypeError("unique-alpha tracing works on the float32 alpha table")
            index = unique if isinstance(unique, UniqueAlphaIndex) else UniqueAlphaIndex(alpha_lookup)
            if index.shape != shape:
                raise ValueError("UniqueAlphaIndex was built for a table of shape %r" % (index.shape,))
            fa_u, w_u = metric.trace_alpha_table(index.alpha, r_obs)
            fa = fa_u[index.inverse].reshape(shape)
            w = w_u.view(t.int16)[index.inverse].view(t.uint16).reshape(shape)
            if tensor_in:
                return fa, w, n, index.n_unique
            return dev.d2h(fa, "fa32"), dev.d2h(w, "w16"), n, index.n_unique
        if is_f32:
            a32 = alpha_lookup.contiguous() if tensor_in else dev.h2d(np.asarray(alpha_lookup), "alpha")
            fa, w = metric.trace_alpha_table(a32, r_obs)
        else:
            # the reference traces alpha_lookup.ravel().astype(float64) (image_lens.py:157): a table that
            # is not float32 keeps its precision through the fp64 kernel and only the RESULTS are
            # narrowed (image_lens.py:176-177)
            a64 = (alpha_lookup.to(t.float64).contiguous() if tensor_in
                   else dev.h2d(np.asarray(alpha_lookup, dtype=np.float64), "alpha"))
            fa64 = t.empty(shape, dtype=t.float64, device=a64.device)
            w64 = t.empty(shape, dtype=t.int64, device=a64.device)
            metric.trace_rays_batch(r_obs, a64.view(-1), fa64.view(-1), w64.view(-1), flags=dev.TRACE_HYBRID)
            fa = fa64.to(t.float32)
            w = w64.clamp(0, WINDING_MAX).to(t.uint16)
        if tensor_in:
            return fa, w, n, n
        return dev.d2h(fa, "fa32"), dev.d2h(w, "w16"), n, n

    # generic plug-in metric: the reference's own chunked driver
    a_np = dev.d2h(alpha_lookup) if tensor_in else np.asarray(alpha_lookup)
    alpha_flat = a_np.ravel().astype(np.float64)
    final_alpha_flat = np.full(n, np.nan, dtype=np.float64)
    winding_flat = np.zeros(n, dtype=np.int64)
    chunk = 50_000
    for start in range(0, n, chunk):
        end = min(start + chunk, n)
        metric.trace_rays_batch(r_obs, alpha_flat[start:end], final_alpha_flat[start:end],
                                winding_flat[start:end])
    fa_out = final_alpha_flat.astype(np.float32).reshape(shape)
    w_out = np.clip(winding_flat, 0, WINDING_MAX).astype(WINDING_DTYPE).reshape(shape)
    return fa_out, w_out, n, n


def _axis_refine_columns(width, fx, psi):
    """Columns within Y_AXIS_REFINE_FRAC of the BH's screen column get the tighter integrator
    tolerances (image_lens.py:210-216)."""
    x_cam = (np.arange(width) - width / 2) / fx
    _, bh_x_cam, in_front = _psi_to_cam_projection(psi)
    if not in_front:
        return np.zeros(width, dtype=bool)
    x_rel = x_cam - bh_x_cam
    x_abs_max = max(float(np.max(np.abs(x_rel))), 1e-12)
    return np.abs(x_rel) <= (Y_AXIS_REFINE_FRAC * x_abs_max)


def _axis_refine_columns_device(width, fx, psi):
    """_axis_refine_columns as a uint8 CUDA tensor, built on the device (no host staging, so it can
    be captured in a CUDA graph): the same IEEE subtractions / divisions column by column; the
    maximum of |x_rel| is attained at the first or the last column and is taken on the host."""
    t = dev.torch()
    _, bh_x_cam, in_front = _psi_to_cam_projection(psi)
    if not in_front or width == 0:
        return t.zeros(width, dtype=t.uint8, device=dev.device())
    ends = (np.array([0.0, float(width - 1)]) - width / 2) / fx - bh_x_cam
    x_abs_max = max(float(np.max(np.abs(ends))), 1e-12)
    x_rel = (t.arange(width, dtype=t.float64, device=dev.device()) - width / 2) / fx - bh_x_cam
    return (x_rel.abs() <= (Y_AXIS_REFINE_FRAC * x_abs_max)).to(t.uint8)


def _theta_pixel(image_dimension, fov, psi, rows):
    """Screen angle of every pixel of the first `rows` rows (image_lens.py:194-208), host numpy —
    only used for plug-in metrics that are not served by the CUDA Kerr kernel."""
    height, width = image_dimension
    fx, fy = _focal(image_dimension, fov)
    x_cam = (np.arange(width) - width / 2) / fx
    y_cam = (np.arange(rows) - height / 2) / fy
    _, e_x, e_y, _ = _psi_frame(psi)
    denom = np.sqrt(1.0 + x_cam[None, :]**2 + y_cam[:, None]**2)
    vx, vy, vz = x_cam[None, :] / denom, y_cam[:, None] / denom, 1.0 / denom
    return np.arctan2(vx * e_x[0] + vy * e_x[1] + vz * e_x[2], vx * e_y[0] + vy * e_y[1] + vz * e_y[2])


def precompute_final_alpha_lookup_2d(alpha_lookup, fov, alpha_crit, r_obs, metric,
                                     theta_obs=np.pi / 2, psi=(0.0, 0.0)):
    """One ray per pixel for metrics without spherical symmetry (image_lens.py:185-280):
    -> (final_alpha float32[H,W], winding uint16[H,W], total_rays, traced_rays).

    Per pixel the ray is launched at viewing angle alpha and screen angle theta_pixel; the
    columns around the BH's screen column are traced with tighter tolerances (axis_refine); an
    equatorial observer with psi_y = 0 traces the top half only and mirrors it.  A ``Kerr``
    metric takes the single-launch GPU path (lp_kerr_trace_alpha32: theta_pixel is evaluated on
    the device); any other metric is driven through its own 7-argument ``trace_rays_batch`` in
    50 000-ray chunks exactly like the reference."""
    tensor_in = _is_tensor(alpha_lookup)
    shape = tuple(alpha_lookup.shape)
    height, width = shape
    fx, _ = _focal(shape, fov)
    refine_cols = _axis_refine_columns(width, fx, psi)
    use_tb_symmetry = bool(np.isclose(theta_obs, np.pi / 2) and np.isclose(psi[0], 0.0))
    trace_rows = (height + 1) // 2 if use_tb_symmetry else height
    n_traced = trace_rows * width
    if use_tb_symmetry:
        print(f"  tracing {n_traced:,} rays with top/bottom symmetry ({height * width:,} pixels total)")
    else:
        print(f"  tracing {n_traced:,} rays ({height * width:,} pixels total)")

    if isinstance(metric, Kerr) and type(metric).trace_rays_batch is Kerr.trace_rays_batch:
        t = dev.torch()
        if tensor_in:
            a32 = alpha_lookup[:trace_rows].to(t.float32).contiguous()
        else:
            a32 = dev.h2d(np.ascontiguousarray(np.asarray(alpha_lookup, dtype=np.float32)[:trace_rows]), "alpha")
        fa_out = t.full(shape, float("nan"), dtype=t.float32, device=a32.device)
        w_out = t.zeros(shape, dtype=t.uint16, device=a32.device)
        if n_traced:
            cam = dev.camera_vector(shape, fov, psi, _psi_frame)
            d_cols = _axis_refine_columns_device(width, fx, psi)
            fa, w = metric.trace_alpha_table_2d(a32, cam, r_obs, theta_obs, row0=0, refine_cols=d_cols)
            fa_out[:trace_rows] = fa
            w_out[:trace_rows] = w
        if use_tb_symmetry:
            top_half = height // 2
            if top_half > 0:                                 # image_lens.py:272-276
                fa_out[height - top_half:] = fa_out[:top_half].flip(0)
                w_out.view(t.int16)[height - top_half:] = w_out.view(t.int16)[:top_half].flip(0)
        if tensor_in:
            return fa_out, w_out, height * width, n_traced
        return dev.d2h(fa_out, "fa32"), dev.d2h(w_out, "w16"), height * width, n_traced

    # plug-in metric: the reference's own chunked host driver
    a_np = dev.d2h(alpha_lookup) if tensor_in else np.asarray(alpha_lookup)
    alpha_f64 = a_np[:trace_rows, :].ravel().astype(np.float64)
    theta_f64 = _theta_pixel(shape, fov, psi, trace_rows).ravel().astype(np.float64)
    axis_flat = np.broadcast_to(refine_cols[None, :], (trace_rows, width)).ravel().astype(np.bool_)
    fa_buf = np.full(alpha_f64.size, np.nan, dtype=np.float64)
    w_buf = np.zeros(alpha_f64.size, dtype=np.int64)
    chunk = 50_000
    for start in range(0, alpha_f64.size, chunk):
        end = min(start + chunk, alpha_f64.size)
        metric.trace_rays_batch(r_obs, alpha_f64[start:end], theta_f64[start:end], theta_obs,
                                axis_flat[start:end], fa_buf[start:end], w_buf[start:end])
    fa_out = np.full(shape, np.nan, dtype=np.float32)
    w_out = np.zeros(shape, dtype=WINDING_DTYPE)
    fa_out[:trace_rows] = fa_buf.astype(np.float32).reshape(trace_rows, width)
    w_out[:trace_rows] = np.clip(w_buf, 0, WINDING_MAX).astype(WINDING_DTYPE).reshape(trace_rows, width)
    if use_tb_symmetry:
        top_half = height // 2
        if top_half > 0:
            fa_out[height - top_half:] = fa_out[:top_half][::-1]
            w_out[height - top_half:] = w_out[:top_half][::-1]
    return fa_out, w_out, height * width, n_traced


# ============================================================================
# Rendering
# ============================================================================

def _source_layout(source_image):
    shape = tuple(source_image.shape)
    if len(shape) == 2:
        return shape[0], shape[1], 1
    if len(shape) == 3 and 1 <= shape[2] <= 4:
        return shape
    raise ValueError("source_image must be [H, W] or [H, W, C<=4], got shape %r" % (shape,))


def render_lensed_image(source_image, alpha_lookup, final_alpha_lookup, winding_lookup,
                        alpha_crit, fov, render_loop_around=False, psi=(0.0, 0.0), *,
                        sampling=SAMPLE_NEAREST, unit_u8=False):
    """Deflection -> background remap (image_lens.py:296-397) by lp_remap.  Returns an
    array shaped and typed like ``source_image`` (uint8 / float32 / float64).
    ``alpha_lookup`` and ``alpha_crit`` are accepted and unused, as in the reference.
    numpy in -> numpy out; CUDA tensors in -> CUDA tensor out.

    ``unit_u8=True`` (uint8 images only) treats the bytes as float32 value/255 — the 8-bit
    boundary of ``main()`` (imread -> /255 ... imsave): the result is the byte image
    ``trunc(255 * render(source/255))``.  Without it a uint8 image behaves as in the
    reference (colour constants are cast to 0/1)."""
    t = dev.torch()
    e = _lib.ext()
    height, width, channels = _source_layout(source_image)
    tensor_in = _is_tensor(source_image)
    if not tensor_in:
        src_np = np.asarray(source_image)
        if src_np.dtype not in (np.uint8, np.float32, np.float64):
            raise TypeError("source_image dtype %s is not supported on the GPU path "
                            "(uint8, float32, float64)" % src_np.dtype)
    # the lookups first: when they are the (pinned) arrays this package returned they upload in place,
    # and those DMAs then run while the pageable source image is being staged by the host threads
    fa = final_alpha_lookup if _is_tensor(final_alpha_lookup) else \
        dev.h2d(np.asarray(final_alpha_lookup, dtype=np.float32), "fa32")
    if winding_lookup is None:
        w = None
    elif _is_tensor(winding_lookup):
        w = winding_lookup
    else:
        w_np = np.asarray(winding_lookup)
        if w_np.dtype != WINDING_DTYPE:          # uint16 is already in range: no host pass over it
            w_np = np.clip(w_np, 0, WINDING_MAX).astype(WINDING_DTYPE)
        w = dev.h2d(w_np, "w16")
    src = source_image.contiguous() if tensor_in else dev.h2d(src_np, "src")
    if tuple(fa.shape) != (height, width):
        raise ValueError("final_alpha_lookup shape %r does not match the image %r"
                         % (tuple(fa.shape), (height, width)))
    out = t.empty_like(src)
    cam = dev.camera_vector((height, width), fov, psi, _psi_frame)
    e.remap(src, channels, cam, fa.contiguous(), None if w is None else w.contiguous(),
            bool(render_loop_around), int(sampling), 0, height, out, bool(unit_u8))
    return out if tensor_in else dev.d2h(out, "frame")


def render_frame(source_image, fov, r_obs, metric, psi=(0.0, 0.0), render_loop_around=False, *,
                 sampling=SAMPLE_NEAREST, rows=None, return_lookups=False, stats=None,
                 flags=dev.TRACE_HYBRID, out=None, unit_u8=False, theta_obs=np.pi / 2, bands=None):
    """Fully fused device-resident frame (lp_render_frame): build_alpha_lookup +
    precompute_final_alpha_lookup + render_lensed_image in ONE launch, bit-identical to
    running the three stages back to back.  ``source_image`` is a CUDA tensor [H,W(,C)];
    ``rows=(row0, n_rows)`` renders a row tile (multi-GPU sharding).  Returns the CUDA
    frame tensor (and the float32 / uint16 lookups with ``return_lookups``).

    ``bands=(band_rows, band_stride)`` makes the tile an INTERLEAVED set of rows (tile-local row r
    is frame row ``row0 + (r // band_rows) * band_stride + r % band_rows``; dist.band_layout):
    the black hole sits in the centre rows, so contiguous tiles are unevenly expensive.  With
    ``flags | RENDER_OUT_FRAME_ROWS`` ``out`` is the FULL-frame tensor from row ``row0`` on and
    every pixel is stored at its frame row (peer frames); otherwise ``out`` is a compact tile.

    A ``Kerr`` metric takes the device-resident three-launch path (alpha lookup, Kerr tracer
    with the per-pixel screen angle, remap) for the observer inclination ``theta_obs``; every
    row of the tile is traced (no top/bottom mirror)."""
    t = dev.torch()
    e = _lib.ext()
    height, width, channels = _source_layout(source_image)
    row0, n_rows = (0, height) if rows is None else rows
    src = source_image.contiguous()
    tile_shape = (n_rows, width) + tuple(source_image.shape[2:])
    band_rows, band_stride = (0, 0) if bands is None else (int(bands[0]), int(bands[1]))
    if out is None:
        if flags & dev.RENDER_OUT_FRAME_ROWS:
            raise ValueError("a frame-addressed tile needs the caller's frame tensor as `out`")
        out = t.empty(tile_shape, dtype=src.dtype, device=src.device)
    for base in (Kerr, Schwarzschild):
        if isinstance(metric, base) and type(metric).trace_rays_batch is not base.trace_rays_batch:
            raise NotImplementedError("%s overrides trace_rays_batch: render_frame runs %s's own tracer on the "
                                      "device; use the staged calls (precompute_final_alpha_lookup -> "
                                      "render_lensed_image), which drive the metric's own method"
                                      % (type(metric).__name__, base.__name__))
    if isinstance(metric, Kerr):
        if band_rows:
            raise NotImplementedError("interleaved row bands cover the Schwarzschild fused kernel")
        return _render_frame_kerr(src, channels, fov, r_obs, metric, psi, render_loop_around, sampling,
                                  (row0, n_rows), return_lookups, out, unit_u8, theta_obs)
    if not isinstance(metric, Schwarzschild):
        raise NotImplementedError("render_frame covers Schwarzschild and Kerr metrics")
    fa = w = None
    if return_lookups:
        fa = t.empty((n_rows, width), dtype=t.float32, device=src.device)
        w = t.empty((n_rows, width), dtype=t.uint16, device=src.device)
    cam = dev.camera_vector((height, width), fov, psi, _psi_frame)
    e.render_frame(src, channels, cam, int(row0), int(n_rows), float(metric.M), float(metric.R_S),
                   float(r_obs), dev.PHI_MAX, dev.H_MAX, bool(render_loop_around), int(sampling),
                   out, fa, w, stats, int(flags), bool(unit_u8), band_rows, band_stride)
    if return_lookups:
        return out, fa, w
    return out


def _render_frame_kerr(src, channels, fov, r_obs, metric, psi, loop_around, sampling, rows, return_lookups, out,
                       unit_u8, theta_obs):
    t = dev.torch()
    e = _lib.ext()
    height, width = int(src.shape[0]), int(src.shape[1])
    row0, n_rows = rows
    cam = dev.camera_vector((height, width), fov, psi, _psi_frame)
    a32 = t.empty((n_rows, width), dtype=t.float32, device=src.device)
    e.build_alpha_lookup(cam, int(row0), int(n_rows), -1, a32)
    fx, _ = _focal((height, width), fov)
    d_cols = _axis_refine_columns_device(width, fx, psi)
    fa, w = metric.trace_alpha_table_2d(a32, cam, r_obs, theta_obs, row0=row0, refine_cols=d_cols)
    e.remap(src, channels, cam, fa, w, bool(loop_around), int(sampling), int(row0), int(n_rows), out, bool(unit_u8))
    if return_lookups:
        return out, fa, w
    return out


class LensPipeline:
    """Keeps the source image resident on the GPU and renders frames of it for varying
    observers (parameter sweeps: SURVEY.md §8(d) config 5)."""

    def __init__(self, source_image, vertical_fov_deg=40.0, metric=None):
        t = dev.torch()
        self.metric = metric if metric is not None else Schwarzschild(M=1.0)
        self.src = source_image if _is_tensor(source_image) else dev.h2d(np.asarray(source_image), "src")
        self.height, self.width = int(self.src.shape[0]), int(self.src.shape[1])
        vfov = np.radians(vertical_fov_deg)
        self.fov = (2 * np.arctan(np.tan(vfov / 2) * self.width / self.height), vfov)  # image_lens.py:461-463
        self._t = t

    def render(self, r_obs, psi=(0.0, 0.0), rows=None, stats=None, flags=dev.TRACE_HYBRID, out=None, bands=None,
               unit_u8=False):
        return render_frame(self.src, self.fov, r_obs, self.metric, psi=psi, rows=rows, stats=stats,
                            flags=flags, out=out, bands=bands, unit_u8=unit_u8)

    def capture_sweep(self, params, out, flags=dev.TRACE_HYBRID, lanes=2, unit_u8=False):
        """Capture a whole parameter sweep — ``params`` = [(r_obs, psi), ...], frame j written to
        ``out[j]`` — in ONE CUDA graph and return it (``graph.replay()`` re-renders the sweep).
        Small frames are launch-bound when issued one by one from Python (a 1024x1024 frame is
        ~0.11 ms of kernel time); the graph replays the launches with the per-frame constants
        baked in, as ``lanes`` independent chains (frame j on chain j % lanes) so that one frame's
        last CTAs share the SMs with the next frame's first ones instead of draining alone."""
        t = self._t
        if len(params) != int(out.shape[0]):
            raise ValueError("out must hold one frame per sweep point")
        side = t.cuda.Stream(device=self.src.device)
        side.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(side):                       # warm-up outside the capture (module load)
            for j, (r_obs, psi) in enumerate(params[:2]):
                self.render(r_obs, psi=psi, flags=flags, out=out[j], unit_u8=unit_u8)
        t.cuda.current_stream().wait_stream(side)
        graph = t.cuda.CUDAGraph()
        lanes = max(1, min(int(lanes), len(params)))
        with t.cuda.graph(graph):
            if lanes == 1:
                for j, (r_obs, psi) in enumerate(params):
                    self.render(r_obs, psi=psi, flags=flags, out=out[j], unit_u8=unit_u8)
            else:
                cur = t.cuda.current_stream()
                chains = [t.cuda.Stream(device=self.src.device) for _ in range(lanes)]
                for st in chains:                         # fork
                    st.wait_stream(cur)
                for j, (r_obs, psi) in enumerate(params):
                    with t.cuda.stream(chains[j % lanes]):
                        self.render(r_obs, psi=psi, flags=flags, out=out[j], unit_u8=unit_u8)
                for st in chains:                         # join
                    cur.wait_stream(st)
        return graph


class HostFramePipeline:
    """Host image in -> lensed host frame out, for streams of frames (video, sweeps with a
    changing background): every frame's source is copied from (pinned) host memory, rendered
    by the fused kernel and copied back, on ``depth`` CUDA streams with their own device
    buffers, so frame k+1's host->device copy overlaps frame k's device->host copy (PCIe is
    full duplex) and the render hides under both.  depth = 3: a slot is busy for upload + render
    + download (2.1 + 0.9 + 2.1 ms at 4K float32), so two slots leave a bubble on the copy
    engines (2.5 ms per frame); three reach the duplex PCIe rate (2.1 ms, tools/pcie_probe.py).

        pipe = HostFramePipeline((H, W, 3), torch.float32, 40.0, metric)
        for src, dst in zip(host_sources, host_frames):     # pinned CPU tensors
            pipe.submit(src, r_obs, out=dst)
        pipe.synchronize()

    ``rows=(row0, n)`` renders a row tile of the frame (multi-GPU sharding); ``out`` then
    has n rows."""

    def __init__(self, shape, dtype, vertical_fov_deg=40.0, metric=None, depth=3, unit_u8=False):
        t = dev.torch()
        self.unit_u8 = bool(unit_u8)      # uint8 frames standing for float32/255 (see render_lensed_image)
        self.metric = metric if metric is not None else Schwarzschild(M=1.0)
        self.shape = tuple(shape)
        self.height, self.width = self.shape[0], self.shape[1]
        vfov = np.radians(vertical_fov_deg)
        self.fov = (2 * np.arctan(np.tan(vfov / 2) * self.width / self.height), vfov)
        device = dev.device()
        self._slots = []
        for _ in range(max(1, int(depth))):
            st = t.cuda.Stream(device=device)
            with t.cuda.stream(st):         # the slot's buffers are allocated on (and only ever used on) its stream
                src = t.empty(self.shape, dtype=dtype, device=device)
            self._slots.append(dict(stream=st, src=src, frame=None))
        self._k = 0
        self._t = t

    def submit(self, host_src, r_obs, psi=(0.0, 0.0), out=None, rows=None, flags=dev.TRACE_HYBRID, fov=None):
        """Enqueue one frame; returns the host tensor that will hold it after synchronize()."""
        t = self._t
        slot = self._slots[self._k % len(self._slots)]
        self._k += 1
        n_rows = self.height if rows is None else rows[1]
        tile_shape = (n_rows,) + self.shape[1:]
        if out is None:
            out = t.empty(tile_shape, dtype=slot["src"].dtype).pin_memory()
        st = slot["stream"]
        st.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(st):
            if slot["frame"] is None or tuple(slot["frame"].shape) != tile_shape:
                # allocated on the slot's stream: the caching allocator then orders the reuse of the
                # old block after the work that stream still has queued on it
                slot["frame"] = t.empty(tile_shape, dtype=slot["src"].dtype, device=slot["src"].device)
            slot["src"].copy_(host_src, non_blocking=True)
            render_frame(slot["src"], self.fov if fov is None else fov, r_obs, self.metric, psi=psi, rows=rows,
                         flags=flags, out=slot["frame"], unit_u8=self.unit_u8)
            out.copy_(slot["frame"], non_blocking=True)
        return out

    def synchronize(self):
        """Make the current stream wait for every submitted frame (and block the host)."""
        cur = self._t.cuda.current_stream()
        for slot in self._slots:
            cur.wait_stream(slot["stream"])
        cur.synchronize()


# ============================================================================
# Benchmark
# ============================================================================

def print_benchmark_summary(image_dimension, alpha_crit, total_rays, traced_rays, timings):
    height, width = image_dimension
    pixels = width * height
    render_time = max(timings.get("render", 0.0), 1e-12)
    total_time = max(timings.get("total", 0.0), 1e-12)
    print("\nBenchmark summary")
    print(f"  resolution: {width}x{height} ({pixels:,} pixels)")
    print(f"  alpha_crit: {alpha_crit:.6f} rad")
    print(f"  total rays: {total_rays:,}")
    print(f"  traced rays: {traced_rays:,}")
    for key in ("load_image", "build_lookup", "precompute", "render", "save_image", "total"):
        print(f"  {key:<26}{timings.get(key, 0.0):>10.3f} s")
    print(f"  {'render_throughput':<26}{(pixels / render_time) / 1e6:>10.2f} MPix/s")
    print(f"  {'overall_throughput':<26}{(pixels / total_time) / 1e6:>10.2f} MPix/s")


# ============================================================================
# Image IO either side of the path (matplotlib when present, Pillow otherwise)
# ============================================================================

def _imread(path):
    try:
        import matplotlib.image as mpimg
        return mpimg.imread(path)
    except ImportError:
        from PIL import Image
        return np.asarray(Image.open(path))


def _imsave(path, img):
    try:
        import matplotlib.image as mpimg
        mpimg.imsave(path, img)
    except ImportError:
        from PIL import Image
        arr = np.asarray(img)
        if arr.dtype.kind == "f":
            arr = (np.clip(arr, 0.0, 1.0) * 255).astype(np.uint8)
        Image.fromarray(arr).save(path)


# ============================================================================
# Main
# ============================================================================

def main(metric=None, M=1.0, a=0.0, r_obs_mult=100.0, psi=(0.0, 0.0), vertical_fov_deg=40.0):
    if metric is None:
        metric = Schwarzschild(M=M) if a == 0 else Kerr(M=M, a=a)

    print(f"Metric: {type(metric).__name__} (M={metric.M}, a={getattr(metric, 'a', 0)})")
    timings = {}
    total_start = perf_counter()

    stage_start = perf_counter()
    img = _imread('image.jpg')
    if img.dtype == np.uint8:
        img = img.astype(np.float32) / 255.0
    timings["load_image"] = perf_counter() - stage_start

    height, width = img.shape[:2]
    print(f"Image: {width}x{height}")

    r_obs = r_obs_mult * metric.M
    alpha_crit = metric.alpha_crit(r_obs)
    print(f"r_obs = {r_obs:.1f} M, alpha_crit = {np.degrees(alpha_crit):.4f} deg")

    vertical_fov = np.radians(vertical_fov_deg)
    horizontal_fov = 2 * np.arctan(np.tan(vertical_fov / 2) * width / height)
    fov = (horizontal_fov, vertical_fov)
    psi_y, psi_x = psi
    bh_y_cam, bh_x_cam, bh_in_front = _psi_to_cam_projection(psi)
    bh_in_fov = (bh_in_front and abs(bh_y_cam) <= np.tan(vertical_fov / 2)
                 and abs(bh_x_cam) <= np.tan(horizontal_fov / 2))
    where = "behind observer" if not bh_in_front else ("inside FOV" if bh_in_fov else "outside FOV")
    print(f"BH screen offset: psi_y={np.degrees(psi_y):.4f} deg, "
          f"psi_x={np.degrees(psi_x):.4f} deg ({where})")

    print("Building per-pixel alpha lookup..." if metric.is_spherically_symmetric
          else "Building per-pixel (alpha, theta) lookup...")
    stage_start = perf_counter()
    alpha_lookup = build_alpha_lookup((height, width), fov, psi=psi)
    timings["build_lookup"] = perf_counter() - stage_start

    stage_start = perf_counter()
    if metric.is_spherically_symmetric:                       # image_lens.py:477-498
        final_alpha_lookup, winding_lookup, total_rays, traced_rays = precompute_final_alpha_lookup(
            alpha_lookup, alpha_crit, r_obs, metric)
    else:
        final_alpha_lookup, winding_lookup, total_rays, traced_rays = precompute_final_alpha_lookup_2d(
            alpha_lookup, fov, alpha_crit, r_obs, metric, psi=psi)
    timings["precompute"] = perf_counter() - stage_start

    stage_start = perf_counter()
    lensed_image = render_lensed_image(img, alpha_lookup, final_alpha_lookup, winding_lookup,
                                       alpha_crit, fov, False, psi=psi)
    timings["render"] = perf_counter() - stage_start

    stage_start = perf_counter()
    _imsave('lensed_image.png', lensed_image)
    timings["save_image"] = perf_counter() - stage_start
    timings["total"] = perf_counter() - total_start

    print_benchmark_summary((height, width), alpha_crit, total_rays, traced_rays, timings)


if __name__ == "__main__":
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("--M", type=float, default=1.0, help="BH mass")
    parser.add_argument("--a", type=float, default=0.0, help="BH spin (|a| <= M, 0 = Schwarzschild)")
    parser.add_argument("--r-obs", type=float, default=100.0, help="Observer distance in units of M")
    parser.add_argument("--psi-y", type=float, default=0.0, help="BH vertical offset in deg (+ = top)")
    parser.add_argument("--psi-x", type=float, default=0.0, help="BH horizontal offset in deg (+ = right)")
    parser.add_argument("--fov-v", type=float, default=40.0, help="Vertical field of view in deg")
    args = parser.parse_args()
    main(M=args.M, a=args.a, r_obs_mult=args.r_obs,
         psi=(np.radians(args.psi_y), np.radians(args.psi_x)), vertical_fov_deg=args.fov_v)
