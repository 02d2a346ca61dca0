"""Loader for the UNMODIFIED reference (dhg14n9/Light-path-tracer).

TEST INFRASTRUCTURE — never imported by the product package.  Only
`tests/golden/make_golden.py` and the container-only cross-checks in `tests/`
use it, and only where `/root/reference` exists (it does not on the GPU box).

The reference is a flat set of modules (`metrics`, `image_lens`,
`geodesic_tracer`, `black_hole_shadow`); it needs `matplotlib` only for
plotting / image IO, which we never call, so an import-only stub package
(`oracle/mpl_stub`) is put on `sys.path`.  The reference's numba kernels are
declared `cache=True` (metrics.py:35 …) and the tree is read-only, so
`NUMBA_CACHE_DIR` is pointed at a scratch directory.
"""
import importlib
import os
import sys
import tempfile

REF_DIR = os.environ.get("LP_REFERENCE_DIR", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isfile(os.path.join(REF_DIR, "metrics.py"))


class _Ref:
    pass


_cached = None


def load():
    """Import the reference modules and return them as attributes of one object."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_DIR)
    os.environ.setdefault(
        "NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "lp_ref_numba_cache"))
    stub = os.path.join(_HERE, "mpl_stub")
    added = []
    try:
        import matplotlib  # noqa: F401  (real one, if it ever exists)
    except Exception:
        sys.path.insert(0, stub)
        added.append(stub)
    sys.path.insert(0, REF_DIR)
    added.append(REF_DIR)
    try:
        ref = _Ref()
        for name in ("metrics", "image_lens", "geodesic_tracer", "black_hole_shadow"):
            if name in sys.modules and not getattr(
                    sys.modules[name], "__file__", "").startswith(REF_DIR):
                raise RuntimeError("module name %r already taken by %s" % (
                    name, sys.modules[name].__file__))
            setattr(ref, name, importlib.import_module(name))
    finally:
        for p in added:
            try:
                sys.path.remove(p)
            except ValueError:
                pass
    _cached = ref
    return ref
