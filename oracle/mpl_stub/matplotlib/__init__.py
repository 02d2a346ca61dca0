"""Import-only stand-in for matplotlib (absent from this image, no network).

TEST INFRASTRUCTURE.  The unmodified reference modules under /root/reference
import `matplotlib.pyplot` / `matplotlib.image` at module scope
(geodesic_tracer.py:13, image_lens.py:5, main.py:6, black_hole_shadow.py:2) but
none of the hot-path functions we use as the parity oracle touch them.  This
stub only lets those imports succeed; every attribute access raises.
"""
