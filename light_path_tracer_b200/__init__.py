"""light_path_tracer_b200 — B200-native (sm_100a) implementation of the per-pixel
null-geodesic ray-tracing path of dhg14n9/Light-path-tracer (Schwarzschild: Binet RK4 fast path
and the generic RK45 integrator; Kerr: the reference's Dormand-Prince tracer).

The sub-modules mirror the reference's flat modules and keep their call signatures:

    light_path_tracer_b200.metrics            (reference metrics.py)
    light_path_tracer_b200.image_lens         (reference image_lens.py)
    light_path_tracer_b200.geodesic_tracer    (reference geodesic_tracer.py)
    light_path_tracer_b200.black_hole_shadow  (reference black_hole_shadow.py)
    light_path_tracer_b200.main               (reference main.py)

All compute goes through hand-written CUDA kernels behind a C ABI
(include/lightpath.h, light_path_tracer_b200/_C/liblightpath.so); importing this package
does not need a GPU, calling any compute function does (no CPU fallback).
"""
__version__ = "0.1.0"
