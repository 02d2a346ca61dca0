"""Generic geodesic integrator — drop-in for the reference's ``geodesic_tracer`` module
(reference: geodesic_tracer.py:22-142) on Schwarzschild metrics.

``integrate_geodesic`` / ``trace_ray`` keep the reference's signatures and return
``(solution, outcome)`` with ``solution`` shaped like scipy's ``OdeResult`` (``.t``,
``.y`` [8, n_points], ``.t_events``, ``.y_events``, ``.nfev``, ``.status`` ...).  The
integration — scipy's RK45 controller, events and dense-output root finding — runs in the
CUDA kernel lp_rk45_kernel (csrc/lp_rk45.cu) through lp_schw_rk45_* (include/lightpath.h);
no scipy, no CPU fallback.

Beyond the reference API: ``trace_rays`` (batched, device resident — BASELINE config 3's
"full-frame" form of trace_ray) and ``trace_paths`` (several trajectories in one launch,
what ``plot_trajectories`` needs).
"""
import numpy as np

from . import _device as dev
from . import _lib
from .metrics import Schwarzschild, Kerr

# solve_ivp arguments hard-coded by the reference (geodesic_tracer.py:57-67)
RTOL, ATOL, MAX_STEP = 1e-8, 1e-10, 1.0
_MESSAGES = {0: "The solver successfully reached the end of the integration interval.",
             1: "A termination event occurred.",
             -1: "Required step size is less than spacing between numbers."}
_OUTCOME = {1: "escaped", -1: "captured", 0: "invalid"}


class OdeResult(dict):
    """Attribute-access result bunch with the fields of scipy.integrate.OdeResult that the
    reference's callers read, including ``sol`` — the reference calls
    ``solve_ivp(..., dense_output=True)`` (geodesic_tracer.py:57-67)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


def _builtin_rhs(metric, base, need_initial_conditions):
    """The kernel carries `base`'s own right-hand side (and initial conditions): a subclass that
    overrides them would be silently ignored, so it is refused instead."""
    cls = type(metric)
    names = ["geodesic_equations"] + (["initial_conditions"] if need_initial_conditions else [])
    changed = [n for n in names if getattr(cls, n) is not getattr(base, n)]
    if changed:
        raise NotImplementedError(
            "%s overrides %s: the CUDA generic integrator evaluates %s's own %s on the device and has "
            "no host callback path (no CPU fallback)" % (cls.__name__, ", ".join(changed), base.__name__,
                                                         " / ".join(names)))


class DenseSolution:
    """``OdeResult.sol``: the continuous solution scipy's ``OdeSolution`` of ``RkDenseOutput``
    segments gives (scipy ivp/common.py OdeSolution, rk.py:715-737), rebuilt from what the kernel
    recorded per accepted step: ``sol(t) = y_old + h * Q @ (x, x^2, x^3, x^4)``, ``x = (t - t_old)/h``,
    with Q = K^T P formed on the device.  ``sol(t)`` takes a scalar (-> [8]) or an array (-> [8, n]);
    like scipy it extrapolates with the first / last segment outside ``[t_min, t_max]``."""

    def __init__(self, ts, ys, hs, Qs):
        self.ts = np.asarray(ts, dtype=np.float64)           # breakpoints, ascending; the last one is the event time
        self._y_old = np.asarray(ys, dtype=np.float64)[:, :-1].T.copy()    # [n_seg, 8]
        self._h = np.asarray(hs, dtype=np.float64)           # [n_seg] full step sizes
        self._Q = np.asarray(Qs, dtype=np.float64)           # [n_seg, 8, 4]
        self.n_segments = int(self._h.size)
        self.t_min, self.t_max = (float(self.ts[0]), float(self.ts[-1])) if self.ts.size else (0.0, 0.0)
        self.ascending = True

    def _segments(self, t):
        # OdeSolution: searchsorted(ts, t, side='left') for ascending ts, segment = clip(ind - 1, 0, n - 1)
        ind = np.searchsorted(self.ts, t, side="left")
        return np.clip(ind - 1, 0, self.n_segments - 1)

    def __call__(self, t):
        t = np.asarray(t, dtype=np.float64)
        scalar = t.ndim == 0
        tt = np.atleast_1d(t)
        if self.n_segments == 0:
            raise ValueError("no integration step was taken: there is nothing to interpolate")
        seg = self._segments(tt)
        x = (tt - self.ts[seg]) / self._h[seg]
        p = np.cumprod(np.stack([x, x, x, x]), axis=0)        # (x, x^2, x^3, x^4)  [4, n]
        y = np.einsum("nij,jn->in", self._Q[seg], p) * self._h[seg] + self._y_old[seg].T
        return y[:, 0] if scalar else y


def _require_schwarzschild(metric):
    if not isinstance(metric, Schwarzschild):
        raise NotImplementedError(
            "this entry point evaluates Schwarzschild.initial_conditions on the device "
            "(metrics.py:794-809); %s is not supported" % type(metric).__name__)
    _builtin_rhs(metric, Schwarzschild, True)


def _require_known_metric(metric):
    if not isinstance(metric, (Schwarzschild, Kerr)):
        raise NotImplementedError(
            "the CUDA generic integrator carries the Schwarzschild and Kerr right-hand sides "
            "(metrics.py:763-790, :946-1029); %s is not supported" % type(metric).__name__)
    _builtin_rhs(metric, Kerr if isinstance(metric, Kerr) else Schwarzschild, False)


def _alloc(t, n, device):
    return (t.empty((n, 8), dtype=t.float64, device=device), t.empty(n, dtype=t.float64, device=device),
            t.empty(n, dtype=t.int8, device=device), t.empty((n, 2), dtype=t.int32, device=device),
            t.empty(n, dtype=t.int8, device=device))


def _paths(metric, alphas=None, state0=None, r_obs=None, lambda_max=1000.0, r_stop_inner=None,
           r_stop_outer=None, max_points=0, dense=False):
    """Run the kernel with trajectory recording; grows the trajectory buffer if a ray has
    more accepted points than expected.  Returns host arrays."""
    t = dev.torch()
    e = _lib.ext()
    if state0 is None:
        _require_schwarzschild(metric)
    else:
        _require_known_metric(metric)
    r_in = float(metric.capture_radius() if r_stop_inner is None else r_stop_inner)
    r_out = float(0.0 if r_stop_outer is None else r_stop_outer)
    if r_stop_outer is not None and not r_out > 0.0:
        raise ValueError("r_stop_outer must be positive")
    src = np.ascontiguousarray(alphas if state0 is None else state0, dtype=np.float64)
    n = src.size if state0 is None else src.shape[0]
    d_in = dev.h2d(src.reshape(-1), "rk45_in")
    state, lam, outcome, nsteps, status = _alloc(t, n, d_in.device)
    cap = int(max_points) if max_points else int(min(4096, 2 * lambda_max / MAX_STEP + 64))
    d_dense = None
    while True:
        traj = t.empty((n, cap, 9), dtype=t.float64, device=d_in.device)
        npts = t.empty(n, dtype=t.int32, device=d_in.device)
        if dense:
            d_dense = t.zeros((n, cap, 25), dtype=t.float64, device=d_in.device)
            kerr = isinstance(metric, Kerr)
            e.rk45_paths_dense(1 if kerr else 0, d_in if state0 is None else None, None if state0 is None else d_in,
                               float(metric.M), float(getattr(metric, "a", 0.0)),
                               float(metric.r_plus if kerr else metric.R_S), float(0.0 if r_obs is None else r_obs),
                               float(lambda_max), RTOL, ATOL, MAX_STEP, r_in, r_out, traj, cap, npts, d_dense,
                               state, lam, outcome, nsteps, status)
        elif state0 is None:
            e.rk45_trace_paths(d_in, float(metric.M), float(metric.R_S), float(r_obs), float(lambda_max),
                               RTOL, ATOL, MAX_STEP, r_in, r_out, traj, cap, npts, state, lam, outcome,
                               nsteps, status)
        elif isinstance(metric, Kerr):
            e.kerr_rk45_integrate_paths(d_in, float(metric.M), float(metric.a), float(metric.r_plus),
                                        float(lambda_max), RTOL, ATOL, MAX_STEP, r_in, r_out, traj, cap, npts,
                                        state, lam, outcome, nsteps, status)
        else:
            e.rk45_integrate_paths(d_in, float(metric.M), float(metric.R_S), float(lambda_max), RTOL, ATOL,
                                   MAX_STEP, r_in, r_out, traj, cap, npts, state, lam, outcome, nsteps, status)
        need = int(npts.max().item()) if n else 0
        if need <= cap or max_points:
            break
        cap = need
    res = (traj.cpu().numpy(), npts.cpu().numpy(), state.cpu().numpy(), lam.cpu().numpy(),
           outcome.cpu().numpy(), nsteps.cpu().numpy(), status.cpu().numpy(), r_in, r_out)
    if dense:
        return res + (d_dense.cpu().numpy(),)
    return res


def _solution(traj, n_points, state, lam, nsteps, status, r_in, r_out_used, dense=None):
    n = int(n_points)
    ts = traj[:n, 0].copy()
    ys = traj[:n, 1:].T.copy()
    sol = None
    if dense is not None:
        # rows 1 .. n-1 of the kernel's dense buffer: (h, Q[6][4]) over (t, r, theta, phi, p_r, p_theta);
        # the rows of the two constants of the motion (p_t, p_phi) are zero
        hs = dense[1:n, 0].copy()
        Q6 = dense[1:n, 1:].reshape(-1, 6, 4)
        Q8 = np.zeros((Q6.shape[0], 8, 4))
        Q8[:, [0, 1, 2, 3, 5, 6]] = Q6
        sol = DenseSolution(ts, ys, hs, Q8)
    st = int(status)
    t_events = [np.empty(0), np.empty(0)]
    y_events = [np.empty(0), np.empty(0)]
    if st == 1:
        # which terminal event fired: the one whose radius the event point sits on
        k = 0 if abs(state[1] - r_in) <= abs(state[1] - r_out_used) else 1
        t_events[k] = np.array([lam])
        y_events[k] = state[None, :].copy()
    return OdeResult(t=ts, y=ys, sol=sol, t_events=t_events, y_events=y_events, nfev=int(nsteps[1]),
                     njev=0, nlu=0, status=st, message=_MESSAGES.get(st, ""), success=st >= 0)


def integrate_geodesic(metric, state0, lambda_max=1000.0, r_stop_inner=None, r_stop_outer=None):
    """Integrate the geodesic equations from an explicit 8-D initial state
    (geodesic_tracer.py:22-71).  Returns ``(solution, 'captured' | 'escaped')``."""
    s0 = np.asarray(state0, dtype=np.float64).reshape(1, 8)
    traj, npts, state, lam, outcome, nsteps, status, r_in, r_out, dense = _paths(
        metric, state0=s0, lambda_max=lambda_max, r_stop_inner=r_stop_inner, r_stop_outer=r_stop_outer, dense=True)
    r_out_used = r_out if r_out > 0.0 else float(s0[0, 1]) * 2.0
    sol = _solution(traj[0], npts[0], state[0], lam[0], nsteps[0], status[0], r_in, r_out_used, dense=dense[0])
    return sol, _OUTCOME[int(outcome[0])]


def trace_ray(metric, r_obs, alpha, **kwargs):
    """Trace a single ray from its viewing angle with the full Hamiltonian
    (geodesic_tracer.py:74-82).  Returns ``(solution, outcome)`` or ``(None, 'invalid')``."""
    state0 = metric.initial_conditions(r_obs, alpha)
    if state0 is None:
        return None, 'invalid'
    return integrate_geodesic(metric, state0, **kwargs)


def trace_paths(metric, r_obs, alphas, lambda_max=1000.0, r_stop_inner=None, r_stop_outer=None):
    """``[trace_ray(metric, r_obs, a) for a in alphas]`` in one launch (initial conditions
    evaluated on the device)."""
    alphas = np.atleast_1d(np.asarray(alphas, dtype=np.float64))
    if not isinstance(metric, Schwarzschild):
        # any other metric: its own initial_conditions on the host, one launch for the valid rays
        states = [metric.initial_conditions(r_obs, float(al)) for al in alphas]
        good = [i for i, s0 in enumerate(states) if s0 is not None]
        res = [(None, 'invalid')] * alphas.size
        if good:
            s0 = np.array([states[i] for i in good], dtype=np.float64)
            traj, npts, state, lam, outcome, nsteps, status, r_in, r_out = _paths(
                metric, state0=s0, lambda_max=lambda_max, r_stop_inner=r_stop_inner, r_stop_outer=r_stop_outer)
            for j, i in enumerate(good):
                r_out_used = r_out if r_out > 0.0 else float(s0[j, 1]) * 2.0
                res[i] = (_solution(traj[j], npts[j], state[j], lam[j], nsteps[j], status[j], r_in, r_out_used),
                          _OUTCOME[int(outcome[j])])
        return res
    traj, npts, state, lam, outcome, nsteps, status, r_in, r_out = _paths(
        metric, alphas=alphas, r_obs=r_obs, lambda_max=lambda_max, r_stop_inner=r_stop_inner,
        r_stop_outer=r_stop_outer)
    r_out_used = r_out if r_out > 0.0 else float(r_obs) * 2.0
    res = []
    for i in range(alphas.size):
        if outcome[i] == 0:
            res.append((None, 'invalid'))
        else:
            res.append((_solution(traj[i], npts[i], state[i], lam[i], nsteps[i], status[i], r_in, r_out_used),
                        _OUTCOME[int(outcome[i])]))
    return res


def trace_rays(metric, r_obs, alphas, lambda_max=1000.0, r_stop_inner=None, r_stop_outer=None, *,
               return_status=False):
    """Batched ``trace_ray`` without trajectories: what ``solution.y[:, -1]``, ``solution.t[-1]``
    and ``outcome`` would be for every viewing angle.

    ``alphas``: numpy array or CUDA float64 tensor (any shape).  Returns ``(state [..., 8],
    lambda [...], outcome int8 [...] (1 escaped / -1 captured / 0 invalid), nsteps int32
    [..., 2] = (len(solution.t), solution.nfev))`` — CUDA tensors for tensor input, numpy
    arrays otherwise."""
    t = dev.torch()
    e = _lib.ext()
    _require_schwarzschild(metric)
    tensor_in = type(alphas).__module__.startswith("torch")
    if tensor_in:
        d_a = alphas.contiguous().reshape(-1)
        shape = tuple(alphas.shape)
    else:
        a_np = np.ascontiguousarray(alphas, dtype=np.float64)
        shape = a_np.shape
        d_a = dev.h2d(a_np.reshape(-1), "rk45_in") if a_np.size else t.empty(0, dtype=t.float64, device=dev.device())
    n = d_a.numel()
    state, lam, outcome, nsteps, status = _alloc(t, n, d_a.device)
    r_in = float(metric.capture_radius() if r_stop_inner is None else r_stop_inner)
    r_out = float(0.0 if r_stop_outer is None else r_stop_outer)
    if n:
        e.rk45_trace_batch(d_a, float(metric.M), float(metric.R_S), float(r_obs), float(lambda_max), RTOL, ATOL,
                           MAX_STEP, r_in, r_out, state, lam, outcome, nsteps, status)
    out = (state.reshape(shape + (8,)), lam.reshape(shape), outcome.reshape(shape), nsteps.reshape(shape + (2,)))
    if return_status:
        out = out + (status.reshape(shape),)
    if tensor_in:
        return out
    return tuple(x.cpu().numpy() for x in out)


def draw_black_hole(ax, metric, photon_sphere_label='Photon sphere'):
    """Filled capture radius and (when the metric has one) the photon sphere, as the reference's
    two plotting scripts draw them (geodesic_tracer.py:104-113, main.py:38-47)."""
    ring = np.linspace(0, 2 * np.pi, 200)
    unit = np.stack([np.cos(ring), np.sin(ring)])
    ax.fill(*(metric.capture_radius() * unit), 'k', label='Event horizon')
    r_ph = getattr(metric, 'R_PHOTON', None)
    if r_ph is not None:
        ax.plot(*(r_ph * unit), 'r--', linewidth=1.5, label=photon_sphere_label)


def draw_path(ax, solution, outcome, linewidth=1.2, label=None, dashed_if_captured=True):
    """One trajectory in the orbital plane: x = r cos(phi), y = r sin(phi) from solution.y[1], [3]."""
    r, phi = solution.y[1], solution.y[3]
    escaped = outcome == 'escaped'
    ax.plot(r * np.cos(phi), r * np.sin(phi), color='steelblue' if escaped else 'crimson',
            linestyle='-' if (escaped or not dashed_if_captured) else '--', linewidth=linewidth, label=label)


def plot_trajectories(metric, r_obs, angles_deg, ax=None):
    """Plot photon trajectories for several viewing angles (geodesic_tracer.py:89-142); all rays are
    traced in one launch (trace_paths).  Needs matplotlib (drawing only)."""
    import matplotlib.pyplot as plt
    if ax is None:
        _, ax = plt.subplots(figsize=(10, 10))
    draw_black_hole(ax, metric)
    ax.plot(r_obs, 0, 'go', markersize=10, label=f'Observer (r={r_obs}M)')
    traced = trace_paths(metric, r_obs, np.radians(angles_deg))
    for alpha_deg, (solution, outcome) in zip(angles_deg, traced):
        if solution is not None:
            draw_path(ax, solution, outcome, label=f'α={alpha_deg}° ({outcome})')
    ax.set_title(f'Photon trajectories (critical angle ≈ {np.degrees(metric.alpha_crit(r_obs)):.2f}°)')
    ax.set(xlabel='x / M', ylabel='y / M')
    ax.set_aspect('equal')
    ax.legend(loc='upper left', fontsize=8)
    ax.grid(True, alpha=0.3)
    return ax


def outcome_table(metric, r_obs, angles_deg):
    """The alpha -> b -> CAPTURED/ESCAPED table the reference prints from ``__main__``
    (geodesic_tracer.py:164-172) -> [(alpha_deg, b, 'CAPTURED' | 'ESCAPED')]."""
    rows = []
    for alpha_deg, (_, outcome) in zip(angles_deg, trace_paths(metric, r_obs, np.radians(angles_deg))):
        b = metric.viewing_angle_to_impact_parameter(np.radians(alpha_deg), r_obs)
        rows.append((alpha_deg, float(b), "CAPTURED" if outcome == 'captured' else "ESCAPED"))
    return rows


if __name__ == '__main__':
    metric = Schwarzschild(M=1.0)
    r_obs = 50.0 * metric.M
    angles = [0, 2, 4, 5, 5.5, 5.97, 6.5, 8, 10, 15]
    print("=" * 60)
    print("Geodesic Tracer")
    print("=" * 60)
    print(f"Metric: {type(metric).__name__}")
    print(f"Observer radius: r_obs = {r_obs} M")
    print(f"Critical viewing angle: {np.degrees(metric.alpha_crit(r_obs)):.4f}°")
    print("=" * 60)
    print("\nTracing rays:")
    print("-" * 40)
    for alpha_deg, b, status in outcome_table(metric, r_obs, angles):
        print(f"  α = {alpha_deg:6.2f}°  →  b = {b:6.3f} M  →  {status}")
