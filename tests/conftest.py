import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            path = os.path.join(GOLDEN, name)
            if name.endswith(".json"):
                with open(path) as f:
                    cache[name] = json.load(f)
            else:
                cache[name] = dict(np.load(path))
        return cache[name]
    return load


@pytest.fixture(scope="session")
def oracle():
    import lp_oracle
    lp_oracle.build()
    return lp_oracle


@pytest.fixture(scope="session")
def native():
    """Make sure the in-tree native build exists (compiles without a GPU)."""
    from light_path_tracer_b200 import build as b
    b.build()
    from light_path_tracer_b200 import _lib
    return _lib


def bits_equal(a, b):
    """Bit-for-bit equality of two float arrays, any NaN matching any NaN."""
    a = np.asarray(a)
    b = np.asarray(b)
    return ((a == b) | (np.isnan(a) & np.isnan(b))).all()


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) over entries where b is finite."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = np.isfinite(b)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), floor if floor else 1e-300)))


_PARITY_COUNTS = {}


def record_parity(key, **counts):
    """Per-configuration parity counts (rays that needed a documented clause, pixels that differ and
    why): collected over the session and written to gpurun_out/parity_counts.json (or
    $LP_PARITY_REPORT) so that a regression in a COUNT is visible, not only a crossed bound."""
    _PARITY_COUNTS[key] = {k: (int(v) if isinstance(v, (bool, int, np.integer)) else float(v)) for k, v in counts.items()}


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY_COUNTS:
        return
    path = os.environ.get("LP_PARITY_REPORT") or os.path.join(ROOT, "gpurun_out", "parity_counts.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        old = {}
        if os.path.exists(path):
            with open(path) as f:
                old = json.load(f)
        old.update(_PARITY_COUNTS)
        with open(path, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)
    except OSError:
        pass
