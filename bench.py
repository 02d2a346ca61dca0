#!/usr/bin/env python
"""bench.py — geodesic rays/s and ms/frame of the 4K Schwarzschild lensed render.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric "geodesic rays/sec and ms/frame at 4K"):
the image_lens pipeline — per-pixel viewing angle (float32 table semantics) -> Binet RK4
null-geodesic trace (fp64, the reference's own RK4 stepper; hybrid arithmetic, see `arithmetic` in config) -> deflection
remap of a synthetic float32 RGB checkerboard — at 3840x2160, M=1, r_obs=100 M, vertical
FOV 40 deg, psi=(0,0).  One step = one frame = ONE launch of the fused kernel
(lp_render_frame).  N > 1: weak scaling by row tiles — the frame grows to 3840 x (2160 N) at the
same pixel scale, rank g renders its 2160-row tile in bands, and each finished band is gathered
to rank 0 over NCCL while the next band is being rendered (dist.BandGather).

Prints ONE JSON line on rank 0 (contract in the task statement): value = whole-job rays/s
with the source image resident in HBM; e2e = the same through the host-buffer API with the
H2D copy of the source and the D2H copy of the frame inside the timed region; roofline =
the fused kernel's algorithmic fp64 flops (43/RK4 step + 40/ray) over its
CUDA-event time against the FP64 peak MEASURED in this run by a DFMA micro-benchmark
(MEASURED_PEAKS.json has no fp64 figure); cpu_baseline = the oracle (C/numpy port of the
reference's CPU path) timed on this box's host cores on one full frame.

--impl reference times that CPU port alone (all host threads), one bounded sample per step.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

H0, W0 = 2160, 3840
M, R_OBS, VFOV_DEG = 1.0, 100.0, 40.0
FLOP_PER_STEP, FLOP_PER_RAY = 43, 40     # SURVEY.md §8d work model of the integrator


def fov_for(H, W):
    """40 deg vertical FOV for the 2160-row frame; a taller (weak-scaled) frame keeps the same
    pixel scale, i.e. tan(vfov/2) grows with H, so every GPU's 2160-row tile holds the same
    kind of rays as the single-GPU frame.  hfov from vfov as image_lens.py:461-463."""
    vfov = 2 * np.arctan(np.tan(np.radians(VFOV_DEG) / 2) * H / H0)
    return (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)


# ---------------------------------------------------------------------------------------
# clocks / throttle reasons during the timed region (NVML, sampled from a thread)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": int(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference's numpy/numba path) — checker code, timed only
# ---------------------------------------------------------------------------------------
def cpu_frame(O, H, W, rows=None):
    """build_alpha_lookup -> precompute_final_alpha_lookup -> render_lensed_image on the host
    (image_lens.py:480-505) for the frame rows `rows` (None = all).  Returns (seconds, rays)."""
    fov = fov_for(H, W)
    src = O.checkerboard(H, W)
    t0 = time.perf_counter()
    alpha = O.build_alpha_lookup((H, W), fov, rows=rows)
    fa, w, n, _ = O.precompute_final_alpha_lookup(alpha, M, R_OBS)
    O.render_lensed_image(src, fa, w, fov, rows=rows)
    return time.perf_counter() - t0, int(alpha.size)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is the only work on the box and
    # must use all the host threads it can, so size the OpenMP pool before libgomp initialises
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    import lp_oracle as O
    O.build()
    H, W = H0 * args.gpus, W0
    rows = np.arange(0, H, 8)              # bounded sample: every 8th row of the frame
    for _ in range(args.warmup):
        cpu_frame(O, H, W, rows)
    times, rays = [], 0
    for _ in range(args.steps):
        t, rays = cpu_frame(O, H, W, rows)
        times.append(t)
    total = float(np.sum(times))
    value = rays * args.steps / total
    cores = O.num_threads()
    sample = "every 8th row of the %dx%d frame (%d rays) per step; trace on %d OpenMP threads, " \
             "alpha lookup and remap single-threaded numpy like the reference" % (W, H, rays, cores)
    line = {
        "impl": "reference", "metric": "geodesic rays/sec (4K Schwarzschild lensed render)",
        "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, H, W, args.gather if args.gpus > 1 else "none"),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n, H, W, gather_mode="nccl"):
    return {"workload": "image_lens Schwarzschild lensed render %dx%d (alpha lookup -> Binet RK4 trace "
                        "-> remap), M=1, r_obs=100M, vfov=40deg, psi=(0,0), float32 RGB checkerboard source"
                        % (W, H),
            "rays_per_frame": H * W, "rows_per_gpu": H // n,
            "parallelism": ("row tiles x%d (2160 rows per GPU, same pixel scale as the 1-GPU frame); " % n +
                            ("every rank's render kernel stores its tile straight into rank 0's frame through "
                             "NVLink peer memory (symmetric memory), one 4-byte NCCL all-reduce orders completion"
                             if gather_mode == "peer" else
                             "each tile rendered in bands whose NCCL gather to rank 0 overlaps the next band's render"))
            if n > 1 else "single GPU",
            "arithmetic": "hybrid (LP_TRACE_HYBRID, the image pipeline's default): FMA-contracted RK4 loop, strict "
                          "re-trace of rays longer than 240 steps; same classification / winding / float32 "
                          "final_alpha as the strict kernel on this frame (tests/test_gpu_frame.py)",
            "l2": "256 MiB buffer written between timed steps (L2 flush)"}


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def measure_fp64_peak(torch, ext):
    sink = torch.empty(148 * 32 * 256, dtype=torch.float64, device="cuda")
    iters, best = 20000, None
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ext.bench_dfma(148 * 32, 256, iters, sink)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if i and (best is None or ms < best):
            best = ms
    return 148 * 32 * 256 * iters * 16 / (best * 1e-3) / 1e12


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from light_path_tracer_b200 import _device as dev, _lib
    from light_path_tracer_b200 import image_lens as il
    from light_path_tracer_b200 import dist as lpdist
    from light_path_tracer_b200.metrics import Schwarzschild

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ext = _lib.ext()
    N = world
    H, W = H0 * N, W0
    fov = fov_for(H, W)
    metric = Schwarzschild(M)
    tiles = lpdist.row_tiles(H, N)
    row0, rows = tiles[rank]

    import lp_oracle as O                          # only for the synthetic source + CPU leg
    src_host = torch.from_numpy(O.checkerboard(H, W)).pin_memory()
    src = src_host.to("cuda", non_blocking=True)
    tile_host = torch.empty((rows, W, 3), dtype=torch.float32).pin_memory()
    # N > 1: where the tile goes.  "peer": straight into rank 0's frame through NVLink peer memory
    # (the render kernel's own stores; dist.PeerFrame).  "nccl": local tile, then NCCL gather of
    # row bands overlapped with the next band's render (dist.BandGather).
    bg = pf = None
    gather_mode = "none"
    if N > 1:
        gather_mode = args.gather
        if gather_mode == "peer":
            try:
                pf = lpdist.PeerFrame(H, (W, 3), torch.float32, torch.device("cuda", local), dst=0)
            except Exception as exc:      # symmetric memory unavailable: say so and use NCCL
                sys.stderr.write("PeerFrame unavailable (%r): falling back to the NCCL band gather\n" % (exc,))
                gather_mode = "nccl"
        # every rank must take the same path
        agree = torch.tensor([1.0 if gather_mode == "peer" else 0.0], device="cuda")
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if gather_mode == "peer" and float(agree[0]) == 0.0:
            gather_mode, pf = "nccl", None
        if gather_mode == "nccl":
            bg = lpdist.BandGather(rows, (W, 3), torch.float32, "cuda", dst=0, bands=args.bands)
    tile = pf.tile if pf is not None else (bg.tile if bg is not None else
                                           torch.empty((rows, W, 3), dtype=torch.float32, device="cuda"))
    local_tile = torch.empty((rows, W, 3), dtype=torch.float32, device="cuda") if pf is not None else tile
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if pf is not None:
            il.render_frame(src, fov, R_OBS, metric, rows=(row0, rows), out=tile,
                            flags=dev.TRACE_HYBRID | dev.RENDER_STAGED_STORES)
            pf.complete()
            return
        if bg is None:
            il.render_frame(src, fov, R_OBS, metric, rows=(row0, rows), out=tile)
            return
        # row tile rendered band by band; each finished band is gathered to rank 0 over NCCL
        # while the next band is being rendered (dist.BandGather)
        for first, n in bg.bands:
            il.render_frame(src, fov, R_OBS, metric, rows=(row0 + first, n), out=tile[first:first + n])
            bg.push(first, n)
        bg.finish()

    # end to end through the host-buffer API (image_lens.HostFramePipeline): per frame, the
    # source image goes pinned host -> device, the fused kernel renders this rank's tile, the
    # tile goes device -> pinned host; three slots on three streams, so one frame's H2D overlaps the
    # previous frame's D2H and the render in between leaves no bubble on either copy engine.  Timed as K frames between two events on the current stream.
    # N > 1 (dist.ShardedHostFrames): every rank uploads only ITS rows of the source over its own
    # PCIe link and an NCCL all-gather over NVLink replicates the source on every GPU.
    tile_hosts = [tile_host] + [torch.empty((rows, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    if N == 1:
        host_pipe = il.HostFramePipeline((H, W, 3), torch.float32, VFOV_DEG, metric, depth=3)

        def submit(j):
            host_pipe.submit(src_host, R_OBS, out=tile_hosts[j % 3], rows=(row0, rows), fov=fov)
    else:
        host_pipe = lpdist.ShardedHostFrames((H, W, 3), torch.float32, metric=metric, depth=3)

        def submit(j):
            host_pipe.submit(src_host, fov, R_OBS, out=tile_hosts[j % 3])

    def run_e2e(k):
        for j in range(k):
            submit(j)
        host_pipe.synchronize()

    def timed_e2e(k):
        run_e2e(2)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for j in range(k):
            submit(j)
        for slot in host_pipe._slots:
            torch.cuda.current_stream().wait_stream(slot["stream"])
        b.record()
        barrier()
        return a.elapsed_time(b)

    def timed(step, k):
        """k steps, each bracketed by CUDA events on the launching stream; L2 flushed in between.
        One untimed call first: the first launch of a kernel pays CUDA's lazy module load."""
        step()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        for a, b in ev:
            flush.zero_()
            a.record()
            step()
            b.record()
        barrier()
        return [a.elapsed_time(b) for a, b in ev]

    # --- untimed: work model numerator (sum of RK4 steps) and kernel-only timings --------
    stats = dev.new_stats()
    il.render_frame(src, fov, R_OBS, metric, rows=(row0, rows), out=tile, stats=stats)
    st = dev.read_stats(stats)
    flops_tile = FLOP_PER_STEP * st["sum_steps"] + FLOP_PER_RAY * st["n_rays"]
    peak_tf = measure_fp64_peak(torch, ext)

    for _ in range(args.warmup):
        step_resident()
    with ClockSampler(local) as clk:
        ms = timed(step_resident, args.steps)
    total_ms = float(np.sum(ms))
    run_e2e(args.warmup)
    total_e2e = float(timed_e2e(args.steps))
    # the frame that came back through the host path is the frame the resident path renders
    il.render_frame(src, fov, R_OBS, metric, rows=(row0, rows), out=local_tile)
    if not torch.equal(tile_hosts[(args.steps - 1) % 3], local_tile.cpu()):
        raise SystemExit("e2e frame differs from the device-resident frame")

    # dominant kernel alone (no gather), same events: roofline numerator / denominator
    def step_kernel():
        il.render_frame(src, fov, R_OBS, metric, rows=(row0, rows), out=local_tile)
    ms_k = timed(step_kernel, args.steps)
    kern_ms = float(np.mean(ms_k))

    extra = {}
    if rank == 0 and N == 1:
        a32 = il.build_alpha_lookup((H, W), fov, device=True)
        fa32, w16 = metric.trace_alpha_table(a32, R_OBS)
        t_strict = float(np.mean(timed(lambda: metric.trace_alpha_table(a32, R_OBS, flags=0), args.steps)))
        t_fused = float(np.mean(timed(lambda: metric.trace_alpha_table(a32, R_OBS, flags=1), args.steps)))
        t_hybrid = float(np.mean(timed(lambda: metric.trace_alpha_table(a32, R_OBS, flags=4), args.steps)))
        t_render_strict = float(np.mean(timed(
            lambda: il.render_frame(src, fov, R_OBS, metric, rows=(row0, rows), out=tile, flags=0), args.steps)))
        t_remap = float(np.mean(timed(lambda: il.render_lensed_image(src, a32, fa32, w16, 0.0, fov), args.steps)))
        # the 8-bit boundary of image_lens.main (uint8 image in, uint8 frame out; LP_DTYPE_U8_UNIT)
        src8_host = (src_host * 255).to(torch.uint8).pin_memory()
        pipe8 = il.HostFramePipeline((H, W, 3), torch.uint8, VFOV_DEG, metric, depth=2, unit_u8=True)
        out8 = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        for j in range(4):
            pipe8.submit(src8_host, R_OBS, out=out8[j % 2], fov=fov)
        pipe8.synchronize()
        e8a, e8b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e8a.record()
        for j in range(args.steps):
            pipe8.submit(src8_host, R_OBS, out=out8[j % 2], fov=fov)
        for slot in pipe8._slots:
            torch.cuda.current_stream().wait_stream(slot["stream"])
        e8b.record()
        torch.cuda.synchronize()
        t_u8 = e8a.elapsed_time(e8b) / args.steps
        # the other BASELINE configs that fit one GPU, one timed launch each after a warm-up:
        # config 3 (generic 8-D RK45 integrator over the 4K alpha table) and a Kerr 4K lookup
        from light_path_tracer_b200 import geodesic_tracer as gt
        from light_path_tracer_b200.metrics import Kerr
        a64 = a32.double()

        def once(fn):
            fn()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            res = fn()
            eb.record()
            torch.cuda.synchronize()
            return ea.elapsed_time(eb), res
        t_rk45, rk = once(lambda: gt.trace_rays(metric, R_OBS, a64))
        rk_attempts = float(((rk[3][..., 1].double() - 2) / 6).sum())
        kerr = Kerr(M, 0.9)
        cam_k = dev.camera_vector((H, W), fov, (0.0, 0.0), il._psi_frame)
        t_kerr, _ = once(lambda: kerr.trace_alpha_table_2d(a32, cam_k, R_OBS, np.pi / 2))
        # the same Kerr frame through the reference-facing call, which traces the top half only and
        # mirrors it for an equatorial observer with psi_y = 0 (image_lens.py:236-240, :264-270)
        try:
            t_kerr_api, _ = once(lambda: il.precompute_final_alpha_lookup_2d(
                a32, fov, kerr.alpha_crit(R_OBS), R_OBS, kerr, theta_obs=np.pi / 2, psi=(0.0, 0.0)))
        except Exception as exc:                      # auxiliary figure only
            t_kerr_api = None
            sys.stderr.write("kerr API lookup skipped: %r\n" % (exc,))
        extra = {
            "config3_rk45_frame_4k": {"ms": t_rk45, "rays_per_s": H * W / t_rk45 * 1e3,
                                      "step_attempts_per_s": rk_attempts / t_rk45 * 1e3,
                                      "what": "geodesic_tracer.trace_ray semantics (scipy RK45, rtol 1e-8) for every "
                                              "pixel of the 3840x2160 alpha table, lp_rk45_kernel"},
            "kerr_lookup_4k": {"ms": t_kerr, "rays_per_s": H * W / t_kerr * 1e3,
                               "ms_api_with_mirror": t_kerr_api,
                               "what": "Kerr a=0.9 M, equatorial observer: (alpha, theta) lookup of the full "
                                       "3840x2160 frame without the top/bottom mirror, lp_kerr_queued_kernel; "
                                       "ms_api_with_mirror = image_lens.precompute_final_alpha_lookup_2d on the "
                                       "same device-resident alpha table (top half traced, bottom mirrored)"},
            "e2e_u8_io": {"value": H * W / t_u8 * 1e3, "unit": "rays/s", "ms_per_frame": t_u8,
                          "h2d_bytes_per_step": int(src8_host.numel()), "d2h_bytes_per_step": int(out8[0].numel()),
                          "path": "as e2e, with the uint8 image boundary of image_lens.main (imread uint8 ... imsave "
                                  "8 bit): bytes in, bytes out, /255 and trunc(255 v) on the device"},
            "trace_kernel_strict": {"ms": t_strict, "rays_per_s": H * W / t_strict * 1e3,
                                    "tflops": flops_tile / t_strict / 1e9,
                                    "frac_of_measured_fp64_peak": flops_tile / t_strict / 1e9 / peak_tf},
            "trace_kernel_hybrid": {"ms": t_hybrid, "rays_per_s": H * W / t_hybrid * 1e3,
                                    "tflops": flops_tile / t_hybrid / 1e9,
                                    "frac_of_measured_fp64_peak": flops_tile / t_hybrid / 1e9 / peak_tf},
            "render_kernel_strict": {"ms": t_render_strict, "rays_per_s": H * W / t_render_strict * 1e3,
                                     "tflops": flops_tile / t_render_strict / 1e9,
                                     "frac_of_measured_fp64_peak": flops_tile / t_render_strict / 1e9 / peak_tf},
            "trace_kernel_fma_contracted": {"ms": t_fused, "rays_per_s": H * W / t_fused * 1e3,
                                            "tflops": flops_tile / t_fused / 1e9,
                                            "frac_of_measured_fp64_peak": flops_tile / t_fused / 1e9 / peak_tf},
            "remap_kernel": {"ms": t_remap, "gb_per_s": H * W * 30 / t_remap / 1e6, "bytes_per_px": 30,
                             "frac_of_measured_hbm": H * W * 30 / t_remap / 1e6 / hbm_peak()},
            # kernel (2) in the shape of the `roofline` object (its bound is HBM)
            "roofline_remap": {"bound": "hbm", "kernel": "lp_remap_f32rgb_x4_kernel (stand-alone remap)",
                               "achieved": H * W * 30 / t_remap / 1e6, "peak": hbm_peak(), "unit": "GB/s",
                               "frac": H * W * 30 / t_remap / 1e6 / hbm_peak(),
                               "traffic": ncu_traffic("lp_remap_f32rgb_x4_kernel"),
                               "peak_source": "MEASURED_PEAKS.json hbm_gbs (6650 fallback)"},
        }

    # max over ranks, on the device
    if N > 1:
        t = torch.tensor([total_ms, total_e2e, kern_ms, float(flops_tile), float(st["sum_steps"]),
                          float(st["sum_warp_steps"])], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, total_e2e, kern_ms = float(tmax[0]), float(tmax[1]), float(tmax[2])
        flops_max_tile = float(tmax[3])
        sum_steps, sum_warp = float(tsum[4]), float(tsum[5])
    else:
        flops_max_tile, sum_steps, sum_warp = float(flops_tile), float(st["sum_steps"]), float(st["sum_warp_steps"])

    if rank == 0:
        rays = H * W
        value = rays * args.steps / (total_ms * 1e-3)
        e2e_value = rays * args.steps / (total_e2e * 1e-3)
        achieved = flops_max_tile / (kern_ms * 1e-3) / 1e12
        traffic = ncu_traffic()
        line = {
            "metric": "geodesic rays/sec (4K Schwarzschild lensed render)",
            "value": value, "unit": "rays/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(N, H, W, gather_mode),
            "ms_per_frame": total_ms / args.steps,
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_frame": total_e2e / args.steps,
                    "h2d_bytes_per_step": int(src_host.numel() * 4),
                    "d2h_bytes_per_step": int(rays * 12),
                    "path": ("image_lens.HostFramePipeline: pinned float32 source -> H2D -> lp_render_frame -> D2H "
                             "pinned float32 frame, every frame; 3 slots on 3 streams (frame k+1's H2D overlaps frame k's D2H; the third slot covers the render between them)")
                    if N == 1 else
                            ("dist.ShardedHostFrames: every rank uploads its 1/N of the pinned float32 source, NCCL "
                             "all-gather replicates it over NVLink, lp_render_frame renders the rank's tile, D2H of "
                             "the tile to pinned memory, every frame; 3 slots on 3 streams")},
            "gpu_launches": args.steps * (len(bg.bands) if bg is not None else 1),
            "gather": gather_mode,
            "roofline": {"bound": "fp64", "kernel": "lp_render_kernel (alpha + Binet RK4 + remap, fused)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "peak_source": "measured in this run: DFMA micro-benchmark lp_bench_dfma "
                                        "(MEASURED_PEAKS.json has no fp64 entry; nominal 148 SM x 64 FMA/clk x 2 "
                                        "x 1.965 GHz = 37.2)",
                         "flop_model": "43 flop per RK4 step + 40 per ray (SURVEY.md 8d), the reference's own "
                                       "operation count; the hybrid loop issues 22 FP64-pipe slots per step (strict: "
                                       "34, ceiling 0.5 by construction)",
                         "kernel_ms": kern_ms, "traffic": traffic},
            "rk4_steps_per_frame": sum_steps,
            "lane_efficiency": sum_steps / sum_warp if sum_warp else None,
            "clocks": clk.summary(),
        }
        line.update(extra)
        if N == 1 and not args.no_cpu:
            O.build()
            t_cpu, n_cpu = cpu_frame(O, H, W)
            line["cpu_baseline"] = {
                "value": n_cpu / t_cpu, "unit": "rays/s", "cores": O.num_threads(), "kind": "port",
                "sample": "one full %dx%d frame (%.1f s): oracle port of the reference's CPU path, trace on %d "
                          "OpenMP threads, alpha lookup / remap single-threaded numpy like the reference"
                          % (W, H, t_cpu, O.num_threads())}
            # CPU figures beside the two secondary kernels, on a stratified sample of the same frame
            # (SURVEY.md 8d); the C port has none of scipy's per-call Python cost (15-32 ms/ray there)
            stride = (H * W) // 32768
            a_s = O.build_alpha_lookup((H, W), fov).reshape(-1)[::stride].astype(np.float64)
            t0 = time.perf_counter()
            O.rk45_trace_batch(M, R_OBS, a_s)
            t_rk_cpu = time.perf_counter() - t0
            th_s = il._theta_pixel((H, W), fov, (0.0, 0.0), H).reshape(-1)[::stride]
            if "config3_rk45_frame_4k" in line:
                line["config3_rk45_frame_4k"]["cpu_port"] = {
                    "rays_per_s": a_s.size / t_rk_cpu, "cores": O.num_threads(),
                    "sample": "every %dth pixel (%d rays), C port of scipy RK45 + events" % (stride, a_s.size)}
            if "kerr_lookup_4k" in line:
                t0 = time.perf_counter()
                O.kerr_trace_rays_batch(M, 0.9, R_OBS, a_s, np.asarray(th_s, np.float64), np.pi / 2)
                t_k_cpu = time.perf_counter() - t0
                line["kerr_lookup_4k"]["cpu_port"] = {
                    "rays_per_s": a_s.size / t_k_cpu, "cores": O.num_threads(),
                    "sample": "every %dth pixel (%d rays), C restatement of the numba Kerr DP45 tracer"
                              % (stride, a_s.size)}
        print(json.dumps(line), flush=True)
    if N > 1:
        dist.destroy_process_group()


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0     # B200_PROFILING.md fallback


def ncu_traffic(kernel="lp_render_kernel"):
    """dram bytes per launch of a kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return json.load(f).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--bands", type=int, default=4, help="N > 1, --gather nccl: bands per tile")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = tiles stored straight into rank 0's frame over NVLink peer memory; "
                         "nccl = NCCL gather of row bands overlapped with the render")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
