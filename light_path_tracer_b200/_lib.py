"""Loader for the native code.  There is NO CPU fallback: a missing extension or a
machine without a CUDA device raises, loudly, at the first compute call.

    capi()  -> ctypes handle of _C/liblightpath.so (the C ABI, include/lightpath.h)
    ext()   -> the thin PyTorch C++ extension _C/_lp_torch.so (tensors -> C ABI)
"""
import ctypes
import importlib.machinery
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_C", "liblightpath.so")
EXT_PATH = os.path.join(HERE, "_C", "_lp_torch.so")

_capi = None
_ext = None

_BUILD_HINT = ("native library not built: run `python -m light_path_tracer_b200.build` "
               "(or __graft_entry__.build()) — light_path_tracer_b200 has no CPU fallback")


class lp_camera(ctypes.Structure):
    _fields_ = [("height", ctypes.c_int32), ("width", ctypes.c_int32),
                ("fx", ctypes.c_double), ("fy", ctypes.c_double),
                ("d", ctypes.c_double * 3), ("e_x", ctypes.c_double * 3),
                ("e_y", ctypes.c_double * 3)]


class lp_frame_stats(ctypes.Structure):
    _fields_ = [("n_rays", ctypes.c_uint64), ("n_escaped", ctypes.c_uint64),
                ("n_captured", ctypes.c_uint64), ("n_invalid", ctypes.c_uint64),
                ("n_winding", ctypes.c_uint64), ("sum_steps", ctypes.c_uint64),
                ("sum_warp_steps", ctypes.c_uint64), ("max_steps", ctypes.c_uint32),
                ("max_winding", ctypes.c_uint32), ("min_final_alpha", ctypes.c_double),
                ("max_final_alpha", ctypes.c_double)]


# name -> (restype, argtypes) for every symbol include/lightpath.h declares
_D, _I32, _I64, _U32, _VP = (ctypes.c_double, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32,
                             ctypes.c_void_p)
_CAMP = ctypes.POINTER(lp_camera)
SYMBOLS = {
    "lp_abi_version": (ctypes.c_int, []),
    "lp_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "lp_device_count": (ctypes.c_int, []),
    "lp_device_props": (ctypes.c_int, [_VP, _VP]),
    "lp_camera_init": (ctypes.c_int, [_I32, _I32, _D, _D, _D, _D, _CAMP]),
    "lp_camera_fast_coords": (ctypes.c_int, [_CAMP, _VP, _VP]),
    "lp_schw_trace_batch_f64": (ctypes.c_int, [_VP, _I64, _D, _D, _D, _D, _D, _VP, _VP, _VP, _VP,
                                               _VP, _U32, _VP]),
    "lp_schw_trace_alpha32": (ctypes.c_int, [_VP, _I64, _D, _D, _D, _D, _D, _VP, _VP, _VP, _VP,
                                             _VP, _U32, _VP]),
    "lp_schw_trace_frame": (ctypes.c_int, [_CAMP, _I32, _I32, _D, _D, _D, _D, _D, _VP, _VP, _VP,
                                           _VP, _VP, _VP, _U32, _VP]),
    "lp_build_alpha_lookup": (ctypes.c_int, [_CAMP, _I32, _I32, _I32, _VP, _VP]),
    "lp_remap": (ctypes.c_int, [_VP, _I32, _I32, _CAMP, _VP, _VP, _I32, _I32, _I32, _I32, _VP, _VP]),
    "lp_render_frame": (ctypes.c_int, [_VP, _I32, _I32, _CAMP, _I32, _I32, _D, _D, _D, _D, _D,
                                       _I32, _I32, _VP, _VP, _VP, _VP, _U32, _VP]),
    "lp_render_frame_bands": (ctypes.c_int, [_VP, _I32, _I32, _CAMP, _I32, _I32, _I32, _I32, _D, _D, _D, _D, _D,
                                             _I32, _I32, _VP, _VP, _VP, _VP, _U32, _VP]),
    "lp_peer_signal": (ctypes.c_int, [_VP, _I32, ctypes.c_uint64, _VP]),
    "lp_peer_wait": (ctypes.c_int, [_VP, _I32, ctypes.c_uint64, _U32, _VP, _VP]),
    "lp_shadow_classify": (ctypes.c_int, [_I32, _I32, _D, _D, _VP, _VP, _VP]),
    "lp_frame_stats_reset": (ctypes.c_int, [_VP, _VP]),
    "lp_frame_stats_reduce": (ctypes.c_int, [_VP, _VP, _VP, _VP, _I64, _VP, _VP]),
    "lp_schw_rk45_trace_batch": (ctypes.c_int, [_VP, _I64, _D, _D, _D, _D, _D, _D, _D, _D, _D,
                                                _VP, _VP, _VP, _VP, _VP, _VP]),
    "lp_schw_rk45_trace_paths": (ctypes.c_int, [_VP, _I64, _D, _D, _D, _D, _D, _D, _D, _D, _D, _VP, _I32,
                                                _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "lp_schw_rk45_integrate_paths": (ctypes.c_int, [_VP, _I64, _D, _D, _D, _D, _D, _D, _D, _D, _VP, _I32,
                                                    _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "lp_kerr_rk45_integrate_paths": (ctypes.c_int, [_VP, _I64, _D, _D, _D, _D, _D, _D, _D, _D, _D, _VP, _I32,
                                                    _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "lp_rk45_paths_dense": (ctypes.c_int, [_I32, _VP, _VP, _I64, _D, _D, _D, _D, _D, _D, _D, _D, _D, _D, _VP, _I32,
                                           _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "lp_kerr_trace_batch_f64": (ctypes.c_int, [_VP, _VP, _VP, _I64, _D, _D, _D, _D, _D, _D, _VP, _VP, _VP, _VP, _VP]),
    "lp_kerr_trace_alpha32": (ctypes.c_int, [_VP, _CAMP, _I32, _I32, _VP, _D, _D, _D, _D, _D, _D, _VP, _VP, _VP,
                                             _VP, _VP]),
    "lp_bench_dfma": (ctypes.c_int, [_I32, _I32, _I32, _VP, _VP]),
    "lp_hybrid_retrace_rule": (ctypes.c_int, [_D, _D, _D, _VP, _VP, _VP, _VP]),
}


def capi():
    """ctypes handle with prototypes set for every exported entry point."""
    global _capi
    if _capi is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(_BUILD_HINT + " [missing %s]" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _capi = lib
    return _capi


def ext():
    """The PyTorch extension module (imports torch)."""
    global _ext
    if _ext is None:
        if not os.path.exists(EXT_PATH):
            raise ImportError(_BUILD_HINT + " [missing %s]" % EXT_PATH)
        import torch  # noqa: F401  (must be loaded before the extension's libtorch deps resolve)
        loader = importlib.machinery.ExtensionFileLoader("_lp_torch", EXT_PATH)
        spec = importlib.util.spec_from_loader("_lp_torch", loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _ext = mod
    return _ext


def require_cuda():
    """Raise unless a CUDA device is usable (the product has no CPU path)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("light_path_tracer_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    return torch
