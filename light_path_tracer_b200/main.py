"""Single-ray example — drop-in for the reference's ``main`` module (reference: main.py:12-76):
one ray (r_obs = 50 M, alpha = 8 deg) through the generic integrator, printed summary, and
the trajectory plot when matplotlib is available (drawing is not part of the GPU path)."""
import numpy as np

from .geodesic_tracer import trace_ray
from .metrics import Schwarzschild


def main(metric=None):
    if metric is None:
        metric = Schwarzschild(M=1.0)
    r_obs = 50.0 * metric.M
    alpha_deg = 8.0
    alpha = np.radians(alpha_deg)

    solution, outcome = trace_ray(metric, r_obs, alpha)

    b = metric.viewing_angle_to_impact_parameter(alpha, r_obs)
    print(f"Metric:             {type(metric).__name__}")
    print(f"Observer radius:    r_obs = {r_obs} M")
    print(f"Viewing angle:      α = {alpha_deg}°")
    print(f"Impact parameter:   b = {b:.4f} M")
    print(f"Outcome:            {outcome.upper()}")

    r = solution.y[1]
    phi = solution.y[3]
    x, y = r * np.cos(phi), r * np.sin(phi)
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        print("\n(matplotlib not installed: example_geodesic.png not drawn; "
              f"{x.size} trajectory points computed)")
        return
    fig, ax = plt.subplots(figsize=(10, 10))
    theta = np.linspace(0, 2 * np.pi, 200)
    r_horizon = metric.capture_radius()
    ax.fill(r_horizon * np.cos(theta), r_horizon * np.sin(theta), 'k', label='Event horizon')
    if hasattr(metric, 'R_PHOTON'):
        ax.plot(metric.R_PHOTON * np.cos(theta), metric.R_PHOTON * np.sin(theta), 'r--', linewidth=1.5,
                label='Photon sphere')
    ax.plot(x, y, color='steelblue' if outcome == 'escaped' else 'crimson', linewidth=2,
            label=f'Photon path ({outcome})')
    ax.plot(r_obs, 0, 'go', markersize=12, label='Observer')
    ax.set_xlabel('x / M', fontsize=12)
    ax.set_ylabel('y / M', fontsize=12)
    ax.set_title(f'{type(metric).__name__} geodesic (α = {alpha_deg}°, b = {b:.2f} M)', fontsize=14)
    limit = r_obs * 1.1
    ax.set_xlim(-limit, limit)
    ax.set_ylim(-limit, limit)
    ax.set_aspect('equal')
    ax.legend(loc='upper left', fontsize=10)
    ax.grid(True, alpha=0.3)
    plt.tight_layout()
    plt.savefig('example_geodesic.png', dpi=150)
    print("\nSaved: example_geodesic.png")


if __name__ == '__main__':
    main()
