"""Out-of-bounds guard: every kernel entry point writes inside its output buffers only.
(compute-sanitizer is not available on the GPU pool, so each output is carved out of a larger
tensor pre-filled with a sentinel and the margins are checked after the call; shapes are chosen
so that grids have partial CTAs / warps and the vectorised paths see both aligned and misaligned
pointers.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PAD = 257          # elements of margin on each side (odd: exercises misaligned starts too)


class Guarded:
    def __init__(self, torch, shape, dtype, shift=0):
        self.torch = torch
        n = int(np.prod(shape))
        self.shape = tuple(shape)
        self.sentinel = {torch.float32: 777.0, torch.float64: 777.0, torch.int64: -777, torch.int32: -777,
                         torch.int8: -77, torch.uint8: 177, torch.uint16: 7777}[dtype]
        self.big = torch.full((n + 2 * PAD + 16,), self.sentinel, dtype=torch.float64 if dtype == torch.uint16 else dtype,
                              device="cuda")
        if dtype == torch.uint16:       # torch.full has no uint16 fill: build from int32
            self.big = torch.full((n + 2 * PAD + 16,), self.sentinel, dtype=torch.int32, device="cuda").to(torch.uint16)
        self.lo = PAD + shift
        self.view = self.big[self.lo:self.lo + n].view(self.shape)

    def intact(self):
        n = int(np.prod(self.shape))
        a = self.big[:self.lo].to(self.torch.float64)
        b = self.big[self.lo + n:].to(self.torch.float64)
        return bool((a == float(self.sentinel)).all() and (b == float(self.sentinel)).all())


def test_outputs_stay_in_bounds(native):
    import torch
    from light_path_tracer_b200 import image_lens as il, _lib, _device as dev, black_hole_shadow as bs
    from light_path_tracer_b200.metrics import Schwarzschild, Kerr
    from light_path_tracer_b200 import geodesic_tracer as gt
    e = _lib.ext()
    m = Schwarzschild(1.0)
    checked = []
    for (H, W) in ((37, 100), (41, 67), (33, 64)):
        vfov = np.radians(25.0)
        fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
        cam = dev.camera_vector((H, W), fov, (0.03, -0.02), il._psi_frame)
        src = torch.rand(H, W, 3, device="cuda")
        a32 = il.build_alpha_lookup((H, W), fov, psi=(0.03, -0.02), device=True)
        for shift in (0, 3):                        # element shift: 16-byte aligned, and not
            rows = (5, H - 11)
            n_rows = rows[1]
            # alpha lookup
            g_a = Guarded(torch, (n_rows, W), torch.float32, shift)
            e.build_alpha_lookup(cam, rows[0], n_rows, -1, g_a.view)
            # tracer on the table
            g_fa = Guarded(torch, (H, W), torch.float32, shift)
            g_w = Guarded(torch, (H, W), torch.uint16, shift)
            g_st = Guarded(torch, (H, W), torch.int8, shift)
            g_steps = Guarded(torch, (H, W), torch.int32, shift)
            e.trace_alpha32(a32, 1.0, 2.0, 100.0, dev.PHI_MAX, dev.H_MAX, g_fa.view, g_w.view, g_st.view, g_steps.view,
                            None, dev.TRACE_HYBRID)
            # fused frame tile with lookups, plain and staged stores
            outs = [g_a, g_fa, g_w, g_st, g_steps]
            for flags in (dev.TRACE_HYBRID, dev.TRACE_HYBRID | dev.RENDER_STAGED_STORES, dev.TRACE_STRICT):
                g_out = Guarded(torch, (n_rows, W, 3), torch.float32, shift)
                g_fa2 = Guarded(torch, (n_rows, W), torch.float32, shift)
                g_w2 = Guarded(torch, (n_rows, W), torch.uint16, shift)
                il.render_frame(src, fov, 100.0, m, psi=(0.03, -0.02), rows=rows, flags=flags, out=g_out.view)
                e.render_frame(src, 3, cam, rows[0], n_rows, 1.0, 2.0, 100.0, dev.PHI_MAX, dev.H_MAX, False, 0,
                               g_out.view, g_fa2.view, g_w2.view, None, int(flags), False)
                outs += [g_out, g_fa2, g_w2]
            # remap tile: vector path (float32 RGB) and generic paths (uint8, 1 channel, bilinear)
            fa, w = m.trace_alpha_table(a32, 100.0)
            for srcv, C, sampling in ((src, 3, 0), (src, 3, 1), ((src * 255).to(torch.uint8), 3, 0),
                                      (src[..., 0].contiguous(), 1, 0)):
                shape = (n_rows, W, 3) if C == 3 else (n_rows, W)
                g_out = Guarded(torch, shape, srcv.dtype, shift)
                e.remap(srcv, C, cam, fa[rows[0]:rows[0] + n_rows].contiguous(), w[rows[0]:rows[0] + n_rows].contiguous(),
                        False, sampling, rows[0], n_rows, g_out.view, False)
                outs.append(g_out)
            # Kerr tile
            k = Kerr(1.0, 0.7)
            g_fk = Guarded(torch, (n_rows, W), torch.float32, shift)
            g_wk = Guarded(torch, (n_rows, W), torch.uint16, shift)
            e.kerr_trace_alpha32(a32[rows[0]:rows[0] + n_rows].contiguous(), cam, rows[0], n_rows, None, 1.0, 0.7,
                                 float(k.r_plus), 100.0, 1.2, 5000.0, g_fk.view, g_wk.view, None, None)
            outs += [g_fk, g_wk]
            torch.cuda.synchronize()
            for g in outs:
                assert g.intact(), (H, W, shift, g.shape)
            checked.append(len(outs))
    # 1-D batches: Binet f64, RK45, Kerr
    n = 1003
    al = torch.rand(n, dtype=torch.float64, device="cuda") * 0.4
    th = torch.rand(n, dtype=torch.float64, device="cuda") * 6 - 3
    g_fa = Guarded(torch, (n,), torch.float64)
    g_w = Guarded(torch, (n,), torch.int64)
    m.trace_rays_batch(100.0, al, g_fa.view, g_w.view)
    g_state = Guarded(torch, (n, 8), torch.float64)
    g_lam = Guarded(torch, (n,), torch.float64)
    g_oc = Guarded(torch, (n,), torch.int8)
    g_ns = Guarded(torch, (n, 2), torch.int32)
    g_st = Guarded(torch, (n,), torch.int8)
    e.rk45_trace_batch(al, 1.0, 2.0, 100.0, 1000.0, 1e-8, 1e-10, 1.0, 0.0, 0.0, g_state.view, g_lam.view, g_oc.view,
                       g_ns.view, g_st.view)
    g_fk = Guarded(torch, (n,), torch.float64)
    g_wk = Guarded(torch, (n,), torch.int64)
    Kerr(1.0, -0.4).trace_rays_batch(60.0, al, th, 1.0, None, g_fk.view, g_wk.view)
    g_sh = Guarded(torch, (70, 45), torch.float64)
    e.shadow_classify(70, 45, float(np.radians(40)), 0.1, g_sh.view, None)
    torch.cuda.synchronize()
    for g in (g_fa, g_w, g_state, g_lam, g_oc, g_ns, g_st, g_fk, g_wk, g_sh):
        assert g.intact(), g.shape
    print("guarded outputs checked:", sum(checked) + 10)
