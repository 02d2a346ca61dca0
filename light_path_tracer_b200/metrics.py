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
    """Rotating black hole in Boyer-Lindquist coordinates, spin ``a`` with ``|a| <= M``
    (reference: metrics.py:840-1132).  Ray tracing (``trace_ray``, ``trace_rays_batch``) runs in
    the CUDA kernel lp_kerr_queued_kernel (csrc/lp_kerr.cu: the reference's Dormand-Prince 4(5)
    integrator on the reduced 5-D Hamiltonian state); the closed-form helpers are host
    arithmetic."""

    is_spherically_symmetric = False

    def __init__(self, M=1.0, a=0.0):
        if abs(a) > M:
            raise ValueError(f"|a|={abs(a)} exceeds M={M}")              # metrics.py:849-850
        self.M = M
        self.a = a
        self.r_plus = M + np.sqrt(M**2 - a**2)                            # outer horizon

    def _Sigma(self, r, th):
        return r**2 + self.a**2 * np.cos(th)**2

    def _Delta(self, r):
        return r**2 - 2 * self.M * r + self.a**2

    def capture_radius(self):
        return self.r_plus * 1.01

    # -- spherical photon orbits (metrics.py:864-891) ----------------------------------
    def _unstable_photon_r(self):
        """(prograde, retrograde) circular photon orbit radii — Bardeen's formula."""
        M, a = self.M, self.a
        if a == 0:
            return 3 * M, 3 * M
        return (2 * M * (1 + np.cos(2 / 3 * np.arccos(-a / M))),
                2 * M * (1 + np.cos(2 / 3 * np.arccos(a / M))))

    def _xi_eta(self, r_ph):
        """Conserved (xi, eta) of the spherical photon orbit of radius r_ph."""
        M, a = self.M, self.a
        Delta = self._Delta(r_ph)
        xi = ((r_ph**2 + a**2) / a - 2 * r_ph * Delta / (a * (r_ph - M)))
        eta = (r_ph**3 / (a**2 * (r_ph - M)**2) * (4 * M * Delta - r_ph * (r_ph - M)**2))
        return xi, eta

    def _critical_impact_params(self):
        if self.a == 0:
            raise ValueError("_critical_impact_params undefined for a=0")
        return [self._xi_eta(r_ph) for r_ph in self._unstable_photon_r()]

    def alpha_crit(self, r_obs, theta_obs=np.pi / 2):
        """Conservative shadow envelope: the largest impact parameter over 50 sampled spherical
        photon orbits, never below the Schwarzschild value (metrics.py:893-930)."""
        M, a = self.M, self.a
        if a == 0:
            arg = 3 * np.sqrt(3) * M * np.sqrt(1 - 2 * M / r_obs) / r_obs
            return np.arcsin(np.clip(arg, -1.0, 1.0))
        r_pro, r_ret = self._unstable_photon_r()
        b2_max = 0.0
        for r_ph in np.linspace(r_pro, r_ret, 50):
            xi, eta = self._xi_eta(r_ph)
            b2_max = max(b2_max, xi**2 + max(eta, 0.0))
        b_crit = max(np.sqrt(b2_max), 3 * np.sqrt(3) * M)
        Delta_obs, Sigma_obs = self._Delta(r_obs), self._Sigma(r_obs, theta_obs)
        A = (r_obs**2 + a**2)**2 - a**2 * Delta_obs * np.sin(theta_obs)**2
        arg = b_crit * np.sqrt(Sigma_obs * Delta_obs / A) / r_obs
        return np.arcsin(np.clip(arg, -1.0, 1.0))

    def viewing_angle_to_impact_parameter(self, alpha, r_obs, theta_obs=np.pi / 2):
        if self.a == 0:
            return r_obs * np.sin(alpha) / np.sqrt(1 - 2 * self.M / r_obs)
        Delta, Sigma = self._Delta(r_obs), self._Sigma(r_obs, theta_obs)
        A = (r_obs**2 + self.a**2)**2 - self.a**2 * Delta * np.sin(theta_obs)**2
        return r_obs * np.sin(alpha) * np.sqrt(A / (Sigma * Delta))

    # -- inverse metric and its r / theta derivatives, shared by the two host functions --------
    def _inverse_metric(self, r, th):
        M, a = self.M, self.a
        s, c = np.sin(th), np.cos(th)
        Sigma = r**2 + a**2 * c**2
        Delta = r**2 - 2 * M * r + a**2
        A = (r**2 + a**2)**2 - a**2 * Delta * s**2
        g = dict(tt=-A / (Sigma * Delta), tphi=-2 * M * a * r / (Sigma * Delta), rr=Delta / Sigma,
                 thth=1.0 / Sigma, phiphi=(Delta - a**2 * s**2) / (Sigma * Delta * s**2))
        return g, (s, c, Sigma, Delta, A)

    def geodesic_equations(self, lambda_, state):
        """Hamilton's equations on the 8-D state (metrics.py:946-1029): x' = g^{mu nu} p_nu,
        p' = -(1/2) d g^{ab}/dx p_a p_b, with t and phi cyclic."""
        t, r, th, phi, p_t, p_r, p_th, p_phi = state
        M, a = self.M, self.a
        if r <= self.r_plus * 1.001:
            return [0.0] * 8
        g, (s, c, Sigma, Delta, A) = self._inverse_metric(r, th)
        SD = Sigma * Delta
        dS_r, dD_r = 2 * r, 2 * r - 2 * M
        dA_r = 4 * r * (r**2 + a**2) - a**2 * dD_r * s**2
        dSD_r = dS_r * Delta + Sigma * dD_r
        d_r = dict(tt=-(dA_r * SD - A * dSD_r) / SD**2,
                   tphi=-(2 * M * a * (SD - r * dSD_r)) / SD**2,
                   rr=(dD_r * Sigma - Delta * dS_r) / Sigma**2,
                   thth=-dS_r / Sigma**2,
                   phiphi=(dD_r * SD * s**2 - (Delta - a**2 * s**2) * dSD_r * s**2) / (SD * s**2)**2)
        dS_th = -2 * a**2 * s * c
        dA_th = -a**2 * Delta * 2 * s * c
        num, den = Delta - a**2 * s**2, SD * s**2
        dnum, dden = -a**2 * 2 * s * c, dS_th * Delta * s**2 + SD * 2 * s * c
        d_th = dict(tt=-(dA_th * SD - A * dS_th * Delta) / SD**2,
                    tphi=2 * M * a * r * dS_th / (Sigma**2 * Delta),
                    rr=-Delta * dS_th / Sigma**2,
                    thth=-dS_th / Sigma**2,
                    phiphi=(dnum * den - num * dden) / den**2)

        def contract(d):
            return -0.5 * (d["tt"] * p_t**2 + 2 * d["tphi"] * p_t * p_phi + d["rr"] * p_r**2
                           + d["thth"] * p_th**2 + d["phiphi"] * p_phi**2)
        return [g["tt"] * p_t + g["tphi"] * p_phi, g["rr"] * p_r, g["thth"] * p_th,
                g["tphi"] * p_t + g["phiphi"] * p_phi, 0.0, contract(d_r), contract(d_th), 0.0]

    def initial_conditions(self, r_obs, alpha, theta=0.0, theta_obs=np.pi / 2):
        """Photon at the observer (metrics.py:1033-1109): ``theta`` is the azimuthal screen angle
        (0 = up, pi/2 = right), ``theta_obs`` the observer's inclination; Bardeen's celestial
        coordinates give (L, Q), the null condition gives p_r (inward)."""
        a = self.a
        g, (s, c, Sigma, Delta, A) = self._inverse_metric(r_obs, theta_obs)
        E = 1.0
        rho = r_obs * np.sin(alpha) * np.sqrt(Sigma) / np.sqrt(Delta)
        alpha_screen, beta_screen = -rho * np.sin(theta), -rho * np.cos(theta)
        L = -alpha_screen * s * E
        Q = (beta_screen**2 + c**2 * (alpha_screen**2 - a**2)) * E**2
        p_t, p_phi = -E, L
        Theta = max(Q - c**2 * (L**2 / s**2 - a**2 * E**2), 0.0)
        p_theta = (-1.0 if np.cos(theta) > 0 else 1.0) * np.sqrt(Theta)
        other = (g["tt"] * p_t**2 + 2 * g["tphi"] * p_t * p_phi + g["thth"] * p_theta**2
                 + g["phiphi"] * p_phi**2)
        p_r = -np.sqrt(max(-other / g["rr"], 0.0))
        return [0.0, r_obs, theta_obs, 0.0, p_t, p_r, p_theta, p_phi]

    # -- ray tracing on the GPU ---------------------------------------------------------
    def _lambda_max(self, r_obs):
        return max(5000.0, 6.0 * r_obs)                                   # metrics.py:1120, :1131

    def trace_ray(self, r_obs, alpha, theta=0.0, theta_obs=np.pi / 2, phi_max=50.0,
                  axis_refine=False):
        """-> (final_alpha, n_half_orbits, outcome) (metrics.py:1113-1126); ``phi_max`` is
        accepted and unused, as in the reference."""
        t = dev.torch()
        d = dev.device()
        fa = t.empty(1, dtype=t.float64, device=d)
        w = t.empty(1, dtype=t.int64, device=d)
        st = t.empty(1, dtype=t.int8, device=d)
        _lib.ext().kerr_trace_batch(t.tensor([float(alpha)], dtype=t.float64, device=d),
                                    t.tensor([float(theta)], dtype=t.float64, device=d),
                                    t.tensor([1 if axis_refine else 0], dtype=t.uint8, device=d),
                                    float(self.M), float(self.a), float(self.r_plus), float(r_obs),
                                    float(theta_obs), self._lambda_max(r_obs), fa, w, st, None)
        status = int(st.item())
        if status == 0:
            return np.nan, 0, 'invalid'
        if status == -1:
            return np.nan, int(w.item()), 'captured'
        return float(fa.item()), int(w.item()), 'escaped'

    def trace_rays_batch(self, r_obs, alphas, thetas, theta_obs, axis_refines, out_fa, out_w, *,
                         status=None, steps=None):
        """In-place batch trace (metrics.py:1128-1132): ``out_fa[i]`` = final_alpha or NaN,
        ``out_w[i]`` = n_half_orbits.  numpy arrays are staged through pinned memory and
        written back in place; CUDA tensors (float64 / float64 / uint8 or bool / float64 / int64)
        are used where they are."""
        t = dev.torch()
        e = _lib.ext()
        args = (float(self.M), float(self.a), float(self.r_plus), float(r_obs), float(theta_obs),
                self._lambda_max(r_obs))
        if _is_tensor(alphas):
            ref = None if axis_refines is None else axis_refines.to(t.uint8)
            e.kerr_trace_batch(alphas, thetas, ref, *args, out_fa, out_w, status, steps)
            return
        a_np = np.ascontiguousarray(alphas, dtype=np.float64).reshape(-1)
        n = a_np.size
        if n == 0:
            return
        d_a = dev.h2d(a_np, "kerr_alpha")
        d_t = dev.h2d(np.ascontiguousarray(thetas, dtype=np.float64).reshape(-1), "kerr_theta")
        d_r = None if axis_refines is None else \
            dev.h2d(np.ascontiguousarray(axis_refines).astype(np.uint8).reshape(-1), "kerr_refine")
        d_fa = t.empty(n, dtype=t.float64, device=d_a.device)
        d_w = t.empty(n, dtype=t.int64, device=d_a.device)
        d_st = t.empty(n, dtype=t.int8, device=d_a.device) if status is not None else None
        d_steps = t.empty((n, 2), dtype=t.int32, device=d_a.device) if steps is not None else None
        e.kerr_trace_batch(d_a, d_t, d_r, *args, d_fa, d_w, d_st, d_steps)
        dev.d2h_into(d_fa, out_fa, "fa")
        dev.d2h_into(d_w, out_w, "w")
        if status is not None:
            dev.d2h_into(d_st, status, "st")
        if steps is not None:
            dev.d2h_into(d_steps, steps, "steps")

    def trace_alpha_table_2d(self, alpha32, camera, r_obs, theta_obs, *, row0=0, refine_cols=None,
                             status=None, steps=None):
        """Device-resident tracing stage of precompute_final_alpha_lookup_2d: float32 alpha tile
        [rows, W] (CUDA) -> (final_alpha float32, winding uint16) CUDA tensors, the per-pixel
        screen angle evaluated on the device.  ``camera`` as _device.camera_vector()."""
        t = dev.torch()
        rows = int(alpha32.shape[0])
        fa = t.empty(alpha32.shape, dtype=t.float32, device=alpha32.device)
        w = t.empty(alpha32.shape, dtype=t.uint16, device=alpha32.device)
        _lib.ext().kerr_trace_alpha32(alpha32.contiguous(), camera, int(row0), rows, refine_cols,
                                      float(self.M), float(self.a), float(self.r_plus), float(r_obs),
                                      float(theta_obs), self._lambda_max(r_obs), fa, w, status, steps)
        return fa, w
