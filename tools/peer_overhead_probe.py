"""Where the fixed per-frame cost of the PeerFrame path goes (run under torchrun, N >= 2):
8K frame, this rank's interleaved rows rendered (a) into local memory, (b) into rank 0's frame over
NVLink without the completion flags, (c) the full begin / render / complete step."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from light_path_tracer_b200 import _device as dev, dist as lpdist, image_lens as il  # noqa: E402
from light_path_tracer_b200.metrics import Schwarzschild  # noqa: E402
from light_path_tracer_b200.synthetic import checkerboard  # noqa: E402

rank, N = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
H, W = 4320, 7680
vfov = np.radians(40.0)
fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
m = Schwarzschild(1.0)
src = torch.from_numpy(checkerboard(H, W, np.uint8)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
b4 = lpdist.band_layout(H, N)
pf = lpdist.PeerFrame(H, (W, 3), torch.uint8, torch.device("cuda"), dst=0, band_rows=b4)
local = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
HYB = dev.TRACE_HYBRID


def timed(fn, k=20):
    fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
    dist.barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in ev]))], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


row0, rows = pf.rows
fl = dev.RENDER_STAGED_STORES | dev.RENDER_OUT_FRAME_ROWS


def render_local():
    il.render_frame(src, fov, 100.0, m, rows=pf.rows, bands=pf.bands, out=local[row0:], flags=HYB | fl, unit_u8=True)


def render_peer():
    il.render_frame(src, fov, 100.0, m, rows=pf.rows, bands=pf.bands, out=pf.root[0][row0:], flags=HYB | fl, unit_u8=True)


def step():
    tile, r, b, extra = pf.begin()
    il.render_frame(src, fov, 100.0, m, rows=r, bands=b, out=tile, flags=HYB | extra, unit_u8=True)
    return pf.complete()


def contiguous_local():
    il.render_frame(src, fov, 100.0, m, rows=(rank * (H // N), H // N), out=local[rank * (H // N):(rank + 1) * (H // N)],
                    flags=HYB | dev.RENDER_STAGED_STORES, unit_u8=True)


res = {"ranks": N, "band_rows": b4,
       "contiguous_rows_local_ms": timed(contiguous_local),
       "interleaved_rows_local_ms": timed(render_local),
       "interleaved_rows_peer_ms": timed(render_peer),
       "full_step_ms": timed(step)}
pf.drain()
if rank == 0:
    def whole():
        il.render_frame(src, fov, 100.0, m, out=local, flags=HYB | dev.RENDER_STAGED_STORES, unit_u8=True)
    whole()
    ev = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_(); a.record(); whole(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    res["whole_frame_1_rank_ms"] = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    res["ideal_ms"] = res["whole_frame_1_rank_ms"] / N
    print(res)
dist.barrier()
dist.destroy_process_group()
