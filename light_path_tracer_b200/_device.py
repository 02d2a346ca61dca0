"""Device plumbing shared by the reference-facing modules: host<->device staging through
pinned memory, camera packing, frame-statistics decoding.  PyTorch is used for memory
and streams only; all compute goes through the C ABI (``_lib.ext()``).

Host <-> device staging (the numpy-in / numpy-out drop-in calls, reference
image_lens.py:480-505):

* results come back in PINNED host memory: ``d2h`` copies the device tensor into a fresh pinned
  block of torch's caching host allocator and returns a numpy view of it (kept alive by the
  array's ``base``), so the DMA writes the caller's array directly — no second host copy, no
  page faults on a fresh 100 MB allocation;
* arrays that are already pinned (e.g. the alpha / final_alpha / winding tables this package
  returned a moment ago) are uploaded straight from where they are;
* pageable arrays are staged through a fresh pinned block in chunks, copied by a small thread
  pool (numpy releases the GIL) while the previous chunks are already on the PCIe link.

Every staging block is single-use: it is allocated per call and handed back to the caching host
allocator, which does not reuse a block while a non_blocking copy recorded on it is in flight.
Host->device staging is refused while a CUDA graph is being captured (the copy node would
re-read whatever the block holds at replay time).
"""
import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib

PHI_MAX = 50.0   # metrics.py:833
H_MAX = 0.05     # metrics.py:833 (and :824 for the scalar path)

TRACE_STRICT = 0
TRACE_FUSED = 1
RENDER_STAGED_STORES = 8   # lp_render_frame: 16-byte staged pixel stores (peer-memory tiles)
TRACE_REPACK = 2            # opt-in lane re-packing schedule (lp_repack.cu); the default is one ray per thread
RENDER_OUT_FRAME_ROWS = 16  # interleaved-band tile stored at its frame rows (peer frames)
RENDER_ROW_RUNS = 32        # 8-bit tiles in 32 x 1 warp tiles: full-sector runs for frames many peers store into (lightpath.h)
TRACE_HYBRID = 4   # FMA loop + strict re-trace of the few rays that linger near the photon sphere (lightpath.h)

_CHUNK_BYTES = 8 << 20          # staging granularity of large pageable arrays
_THREADED_MIN_BYTES = 4 << 20   # below this one np.copyto is cheaper than the pool
_pool = None


def torch():
    return _lib.require_cuda()


def device():
    t = torch()
    return t.device("cuda", t.cuda.current_device())


def bind_to_gpu_numa_node(index=None):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (one process per GPU): pinned
    staging memory allocated afterwards is then local to the GPU's PCIe root port instead of
    wherever the launcher started the process — on a two-socket host, remote pinned memory halves
    the aggregate host<->device bandwidth of an 8-GPU job.  Best effort: returns the node (or None
    when the topology is not visible, e.g. in a restricted container) and never raises."""
    try:
        t = torch()
        index = t.cuda.current_device() if index is None else int(index)
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def _copy_pool():
    global _pool
    if _pool is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        _pool = ThreadPoolExecutor(max_workers=max(1, min(8, n)), thread_name_prefix="lp-stage")
    return _pool


def _pinned_empty(shape, dtype):
    """Fresh pinned tensor from torch's caching host allocator (single-use staging block)."""
    t = torch()
    return t.empty(tuple(shape), dtype=dtype, pin_memory=True)


def _is_pinned_array(arr):
    """numpy array backed by page-locked memory (ours or the caller's)?"""
    t = torch()
    if not _NP2T:
        _torch_dtype(np.float32)
    if not (arr.flags.c_contiguous and arr.flags.writeable and arr.dtype in _NP2T):
        return False
    try:
        return bool(t.from_numpy(arr).is_pinned())
    except Exception:
        return False


def _chunks(nbytes):
    n = max(1, (nbytes + _CHUNK_BYTES - 1) // _CHUNK_BYTES)
    step = -(-nbytes // n)
    step = (step + 4095) & ~4095
    return [(o, min(o + step, nbytes)) for o in range(0, nbytes, step)]


def h2d(arr, tag="h2d"):
    """numpy array (any layout) -> contiguous CUDA tensor of the same dtype/shape.
    Stream-ordered on the current stream; does not synchronise."""
    t = torch()
    arr = np.asarray(arr)
    tdt = _torch_dtype(arr.dtype)
    if arr.size == 0:
        return t.empty(arr.shape, dtype=tdt, device=device())
    if t.cuda.is_current_stream_capturing():
        raise RuntimeError("host->device staging inside CUDA graph capture: upload before capturing")
    if _is_pinned_array(arr):
        return t.from_numpy(arr).to(device(), non_blocking=True)
    out = t.empty(arr.shape, dtype=tdt, device=device())
    stage = _pinned_empty(arr.shape, tdt)
    if arr.nbytes < _THREADED_MIN_BYTES or not arr.flags.c_contiguous:
        np.copyto(stage.numpy(), arr)
        out.copy_(stage, non_blocking=True)
        return out
    src_b = arr.reshape(-1).view(np.uint8)
    stage_b = stage.view(-1).view(t.uint8)
    out_b = out.view(-1).view(t.uint8)
    stage_np = stage_b.numpy()
    parts = _chunks(arr.nbytes)
    futs = [_copy_pool().submit(np.copyto, stage_np[a:b], src_b[a:b]) for a, b in parts]
    for (a, b), f in zip(parts, futs):       # chunk k is on the link while chunk k+1 is still being staged
        f.result()
        out_b[a:b].copy_(stage_b[a:b], non_blocking=True)
    return out


def d2h_into(tensor, out, tag="d2h"):
    """CUDA tensor -> caller-owned numpy array (possibly a strided view), in place."""
    t = torch()
    if tensor.numel() == 0:
        return
    tensor = tensor.contiguous()
    if tuple(out.shape) == tuple(tensor.shape) and out.dtype == _numpy_dtype(tensor.dtype) and _is_pinned_array(out):
        t.from_numpy(out).copy_(tensor, non_blocking=True)
        t.cuda.current_stream().synchronize()
        return
    stage = _pinned_empty(tensor.shape, tensor.dtype)
    nbytes = tensor.numel() * tensor.element_size()
    direct = (out.flags.c_contiguous and tuple(out.shape) == tuple(tensor.shape)
              and out.dtype == _numpy_dtype(tensor.dtype) and nbytes >= _THREADED_MIN_BYTES)
    if not direct:
        stage.copy_(tensor, non_blocking=True)
        t.cuda.current_stream().synchronize()
        np.copyto(out, stage.numpy().reshape(out.shape), casting="same_kind")
        return
    stage_b = stage.view(-1).view(t.uint8)
    src_b = tensor.view(-1).view(t.uint8)
    dst_b = out.reshape(-1).view(np.uint8)
    stage_np = stage_b.numpy()
    parts = _chunks(nbytes)
    events = []
    for a, b in parts:
        stage_b[a:b].copy_(src_b[a:b], non_blocking=True)
        ev = t.cuda.Event()
        ev.record()
        events.append(ev)
    futs = []
    for (a, b), ev in zip(parts, events):    # chunk k is copied out while chunk k+1 is still on the link
        ev.synchronize()
        futs.append(_copy_pool().submit(np.copyto, dst_b[a:b], stage_np[a:b]))
    for f in futs:
        f.result()


def d2h(tensor, tag="d2h"):
    """CUDA tensor -> numpy array in pinned host memory (see the module docstring)."""
    t = torch()
    tensor = tensor.contiguous()
    host = _pinned_empty(tensor.shape, tensor.dtype)
    if tensor.numel():
        host.copy_(tensor, non_blocking=True)
        t.cuda.current_stream().synchronize()
    return host.numpy()


_NP2T = {}


def _torch_dtype(dt):
    t = torch()
    if not _NP2T:
        _NP2T.update({np.dtype(np.float64): t.float64, np.dtype(np.float32): t.float32,
                      np.dtype(np.int64): t.int64, np.dtype(np.int32): t.int32, np.dtype(np.int8): t.int8,
                      np.dtype(np.uint8): t.uint8, np.dtype(np.uint16): t.uint16})
    return _NP2T[np.dtype(dt)]


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
