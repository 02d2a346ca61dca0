"""Spacetime metrics — drop-in for the reference's ``metrics`` module on the
Schwarzschild hot path (reference: metrics.py:682-833).

Same class names, method names, positional order, defaults, return types and error
behaviour as the reference's plug-in API (``Metric`` ABC, ``Schwarzschild``); the ray
tracing itself runs in hand-written sm_100a CUDA kernels behind the C ABI
(include/lightpath.h).  No numba, no CPU fallback.

Beyond the reference API (device-resident use, no host round trips):
``Schwarzschild.trace_alpha_table`` and the ``stats=`` keyword.
"""
from abc import ABC, abstractmethod

import numpy as np

from . import _device as dev
from . import _lib


class Metric(ABC):
    """Base class for spacetime metrics (reference: metrics.py:682-728)."""

    is_spherically_symmetric = False

    @abstractmethod
    def geodesic_equations(self, lambda_, state):
        """RHS of Hamilton's equations; state = [t, r, theta, phi, p_t, p_r, p_theta, p_phi]."""

    @abstractmethod
    def initial_conditions(self, r_obs, alpha, theta=0.0, theta_obs=np.pi / 2):
        """Initial 8-D state for a photon at viewing angle alpha, or None."""

    @abstractmethod
    def trace_ray(self, r_obs, alpha, theta=0.0, theta_obs=np.pi / 2, phi_max=50.0,
                  axis_refine=False):
        """-> (final_alpha, n_half_orbits, 'escaped' | 'captured' | 'invalid')."""

    @abstractmethod
    def alpha_crit(self, r_obs, theta_obs=np.pi / 2):
        """Critical viewing angle in radians."""

    @abstractmethod
    def capture_radius(self):
        """Inner stopping radius for integration."""

    def viewing_angle_to_impact_parameter(self, alpha, r_obs, theta_obs=np.pi / 2):
        raise NotImplementedError


def _is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda")


class Schwarzschild(Metric):
    """Non-rotating black hole of mass M (reference: metrics.py:735-833)."""

    is_spherically_symmetric = True

    def __init__(self, M=1.0):
        self.M = M
        self.R_S = 2 * M
        self.R_PHOTON = 3 * M
        self.B_CRIT = 3 * np.sqrt(3) * M

    # -- closed-form scalars: host arithmetic, identical expressions (metrics.py:746-759)
    def _f(self, r):
        return 1 - self.R_S / r

    def capture_radius(self):
        return self.R_S * 1.01

    def alpha_crit(self, r_obs, theta_obs=np.pi / 2):
        arg = self.B_CRIT * np.sqrt(self._f(r_obs)) / r_obs
        return np.arcsin(np.clip(arg, -1.0, 1.0))

    def viewing_angle_to_impact_parameter(self, alpha, r_obs, theta_obs=np.pi / 2):
        return r_obs * np.sin(alpha) / np.sqrt(self._f(r_obs))

    # -- 8-D Hamiltonian system (metrics.py:763-809).  The CUDA RK45 kernel carries its
    #    own copy of this right-hand side; these host versions exist so that code written
    #    against the plug-in API (e.g. a user's own solve_ivp call) keeps working.
    def geodesic_equations(self, lambda_, state):
        t, r, th, phi, p_t, p_r, p_th, p_phi = state
        if r <= self.R_S * 1.001:
            return [0.0] * 8
        f = self._f(r)
        sin_th = np.sin(th)
        s2 = sin_th ** 2
        if s2 < 1e-15:
            s2 = 1e-15
        half = self.R_S / (2 * r**2)
        return [-p_t / f,
                f * p_r,
                p_th / r**2,
                p_phi / (r**2 * s2),
                0.0,
                (-half * (p_t**2 / f**2) - half * p_r**2 + (p_th**2 + p_phi**2 / s2) / r**3),
                np.cos(th) * p_phi**2 / (r**2 * s2 * sin_th),
                0.0]

    def initial_conditions(self, r_obs, alpha, theta=0.0, theta_obs=np.pi / 2):
        b = self.viewing_angle_to_impact_parameter(alpha, r_obs)
        f0 = self._f(r_obs)
        E = 1.0
        L = b * E
        p_r_sq = (E**2 / f0 - L**2 / r_obs**2) / f0
        if p_r_sq < 0:
            return None
        return [0.0, r_obs, np.pi / 2, 0.0, -E, -np.sqrt(p_r_sq), 0.0, L]

    # -- fast path: Binet-equation RK4 on the GPU -------------------------------------
    def trace_ray(self, r_obs, alpha, theta=0.0, theta_obs=np.pi / 2, phi_max=50.0,
                  axis_refine=False):
        """One ray through the CUDA tracer (metrics.py:817-829): honours ``phi_max``,
        fixed step 0.05; ``theta``, ``theta_obs`` and ``axis_refine`` are accepted and
        ignored, as in the reference."""
        t = dev.torch()
        e = _lib.ext()
        a = t.tensor([float(alpha)], dtype=t.float64, device=dev.device())
        fa = t.empty(1, dtype=t.float64, device=a.device)
        w = t.empty(1, dtype=t.int64, device=a.device)
        st = t.empty(1, dtype=t.int8, device=a.device)
        e.trace_batch_f64(a, float(self.M), float(self.R_S), float(r_obs), float(phi_max),
                          dev.H_MAX, fa, w, st, None, None, dev.TRACE_STRICT)
        status = int(st.item())
        if status == 0:
            return np.nan, 0, 'invalid'
        if status == -1:
            return np.nan, int(w.item()), 'captured'
        return float(fa.item()), int(w.item()), 'escaped'

    def trace_rays_batch(self, r_obs, alphas, out_fa, out_w, *, status=None, steps=None,
                         stats=None, flags=dev.TRACE_STRICT):
        """In-place batch trace (metrics.py:831-833): ``out_fa[i]`` = final_alpha or NaN,
        ``out_w[i]`` = n_half_orbits, phi_max=50.0 and h=0.05 hard-coded.

        numpy arrays (any strides; float64 / int64 like the reference's callers pass) are
        staged through pinned memory, traced on the GPU and written back in place.  CUDA
        tensors (contiguous float64 / int64) are used where they are — no copies."""
        e = _lib.ext()
        M, R_S = float(self.M), float(self.R_S)
        if _is_tensor(alphas):
            e.trace_batch_f64(alphas, M, R_S, float(r_obs), dev.PHI_MAX, dev.H_MAX, out_fa, out_w,
                              status, steps, stats, int(flags))
            return
        t = dev.torch()
        a_np = np.asarray(alphas)
        if a_np.dtype != np.float64:
            a_np = a_np.astype(np.float64)
        n = a_np.size
        if n == 0:
            return
        d_a = dev.h2d(a_np.reshape(-1), "alphas")
        d_fa = t.empty(n, dtype=t.float64, device=d_a.device)
        d_w = t.empty(n, dtype=t.int64, device=d_a.device)
        d_st = t.empty(n, dtype=t.int8, device=d_a.device) if status is not None else None
        d_steps = t.empty(n, dtype=t.int32, device=d_a.device) if steps is not None else None
        e.trace_batch_f64(d_a, M, R_S, float(r_obs), dev.PHI_MAX, dev.H_MAX, d_fa, d_w,
                          d_st, d_steps, stats, int(flags))
        dev.d2h_into(d_fa, out_fa, "fa")
        dev.d2h_into(d_w, out_w, "w")
        if status is not None:
            dev.d2h_into(d_st, status, "st")
        if steps is not None:
            dev.d2h_into(d_steps, steps, "steps")

    def trace_alpha_table(self, alpha32, r_obs, *, status=None, steps=None, stats=None,
                          flags=dev.TRACE_HYBRID):
        """Device-resident form of image_lens.precompute_final_alpha_lookup
        (image_lens.py:155-178): float32 CUDA tensor in -> (final_alpha float32,
        winding uint16) CUDA tensors of the same shape, one launch."""
        t = dev.torch()
        e = _lib.ext()
        fa = t.empty(alpha32.shape, dtype=t.float32, device=alpha32.device)
        w = t.empty(alpha32.shape, dtype=t.uint16, device=alpha32.device)
        e.trace_alpha32(alpha32, float(self.M), float(self.R_S), float(r_obs), dev.PHI_MAX,
                        dev.H_MAX, fa, w, status, steps, stats, int(flags))
        return fa, w


class Kerr(Metric):
    """Rotating black hole.  OUT OF SCOPE for this build (next);
    the name exists because the reference's image_lens imports it (image_lens.py:9)."""

    is_spherically_symmetric = False

    def __init__(self, M=1.0, a=0.0):
        if abs(a) > M:
            raise ValueError(f"|a| must be <= M (got a={a}, M={M})")   # metrics.py:849-850
        raise NotImplementedError(
            "Kerr tracing is not part of the B200 hot path yet (Schwarzschild only)")

    def geodesic_equations(self, lambda_, state):
        raise NotImplementedError

    def initial_conditions(self, r_obs, alpha, theta=0.0, theta_obs=np.pi / 2):
        raise NotImplementedError

    def trace_ray(self, r_obs, alpha, theta=0.0, theta_obs=np.pi / 2, phi_max=50.0,
                  axis_refine=False):
        raise NotImplementedError

    def alpha_crit(self, r_obs, theta_obs=np.pi / 2):
        raise NotImplementedError

    def capture_radius(self):
        raise NotImplementedError
