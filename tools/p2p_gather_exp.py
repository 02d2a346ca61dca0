"""Experiment: render row tiles straight into rank 0's frame through NVLink peer memory
(torch symmetric memory) instead of rendering locally and gathering with NCCL."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed._symmetric_memory as symm_mem
from light_path_tracer_b200 import dist as lpdist, image_lens as il
from light_path_tracer_b200.metrics import Schwarzschild

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H0, W = 2160, 3840
H = H0 * world
vfov = 2 * np.arctan(np.tan(np.radians(20.0)) * H / H0)
fov = (2 * np.arctan(np.tan(vfov / 2) * W / H), vfov)
m = Schwarzschild(1.0)
yy, xx = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
r = (((yy // 32) + (xx // 32)) & 1).float()
src = torch.stack([r, 1 - r, xx.float() / W], dim=-1).contiguous()
del yy, xx, r
row0, rows = lpdist.row_tiles(H, world)[rank]

buf = symm_mem.empty((H, W, 3), dtype=torch.float32, device="cuda")
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
root = hdl.get_buffer(0, (H, W, 3), torch.float32)
print(rank, "peer view ok", root.device, root.data_ptr() != buf.data_ptr() or rank == 0, flush=True)
tile_view = root[row0:row0 + rows]
flag = torch.zeros(1, device="cuda")

def step_p2p():
    il.render_frame(src, fov, 100.0, m, rows=(row0, rows), out=tile_view)
    dist.all_reduce(flag)          # completion: every rank's stores are done before rank 0 proceeds

def barrier():
    dist.barrier(); torch.cuda.synchronize()

def timed(fn, k=30):
    for _ in range(5): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
    barrier()
    for a, b in ev:
        a.record(); fn(); b.record()
    barrier()
    t = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in ev]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])

ms_p2p = timed(step_p2p)
g = lpdist.BandGather(rows, (W, 3), torch.float32, "cuda", dst=0, bands=4)
def step_nccl():
    for first, n in g.bands:
        il.render_frame(src, fov, 100.0, m, rows=(row0 + first, n), out=g.tile[first:first + n])
        g.push(first, n)
    g.finish()
ms_nccl = timed(step_nccl)
local_tile = torch.empty((rows, W, 3), device="cuda")
ms_local = timed(lambda: il.render_frame(src, fov, 100.0, m, rows=(row0, rows), out=local_tile))
# correctness: rank 0's buffer after a p2p step == NCCL-gathered frame
buf.zero_(); barrier()
step_p2p(); barrier()
frame = None
for first, n in g.bands:
    il.render_frame(src, fov, 100.0, m, rows=(row0 + first, n), out=g.tile[first:first + n]); g.push(first, n)
frame = g.finish()
if rank == 0:
    print(json.dumps({"world": world, "ms_p2p_direct": ms_p2p, "ms_nccl_band_gather": ms_nccl, "ms_render_local_only": ms_local,
                      "identical": bool(torch.equal(buf, frame)), "rays_per_s_p2p": H * W / ms_p2p * 1e3}), flush=True)
dist.destroy_process_group()
