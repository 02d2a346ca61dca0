"""Single-ray example — drop-in for the reference's ``main`` module (reference: main.py:12-76).

One ray (observer at r = 50 M, viewing angle 8 deg) goes through the generic integrator
(``geodesic_tracer.trace_ray`` -> CUDA kernel lp_rk45_kernel); the summary the reference prints is
printed, and the trajectory figure ``example_geodesic.png`` is drawn when matplotlib is installed
(drawing is host-side and not part of the GPU path).
"""
import numpy as np

from . import geodesic_tracer as gt
from .metrics import Schwarzschild

R_OBS_IN_M = 50.0       # main.py:16
ALPHA_DEG = 8.0         # main.py:18


def _summary(metric, r_obs, b, outcome):
    rows = (("Metric:", type(metric).__name__),
            ("Observer radius:", f"r_obs = {r_obs} M"),
            ("Viewing angle:", f"α = {ALPHA_DEG}°"),
            ("Impact parameter:", f"b = {b:.4f} M"),
            ("Outcome:", outcome.upper()))
    return "\n".join(f"{label:<20}{value}" for label, value in rows)


def _figure(metric, r_obs, b, solution, outcome, path="example_geodesic.png"):
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        print(f"\n(matplotlib not installed: {path} not drawn; {solution.t.size} trajectory points computed)")
        return
    _, ax = plt.subplots(figsize=(10, 10))
    gt.draw_black_hole(ax, metric, photon_sphere_label='Photon sphere')
    gt.draw_path(ax, solution, outcome, linewidth=2, label=f'Photon path ({outcome})', dashed_if_captured=False)
    ax.plot(r_obs, 0, 'go', markersize=12, label='Observer')
    ax.set(xlabel='x / M', ylabel='y / M', xlim=(-1.1 * r_obs, 1.1 * r_obs), ylim=(-1.1 * r_obs, 1.1 * r_obs))
    ax.xaxis.label.set_size(12)
    ax.yaxis.label.set_size(12)
    ax.set_title(f'{type(metric).__name__} geodesic (α = {ALPHA_DEG}°, b = {b:.2f} M)', fontsize=14)
    ax.set_aspect('equal')
    ax.legend(loc='upper left', fontsize=10)
    ax.grid(True, alpha=0.3)
    plt.tight_layout()
    plt.savefig(path, dpi=150)
    print(f"\nSaved: {path}")


def main(metric=None):
    metric = Schwarzschild(M=1.0) if metric is None else metric
    r_obs = R_OBS_IN_M * metric.M
    alpha = np.radians(ALPHA_DEG)
    solution, outcome = gt.trace_ray(metric, r_obs, alpha)
    b = metric.viewing_angle_to_impact_parameter(alpha, r_obs)
    print(_summary(metric, r_obs, b, outcome))
    _figure(metric, r_obs, b, solution, outcome)


if __name__ == '__main__':
    main()
