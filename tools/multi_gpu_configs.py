"""BASELINE configs 4 and 5 under torchrun (one process per GPU):

  config 4: 7680x4320 lensed render, rows sharded over the ranks, NCCL gather to rank 0
            (strong scaling of ONE frame) — pipelined band gather vs plain gather
  config 5: 512-frame observer sweep at 1024x1024, frames sharded round-robin, no collective

    python -m torch.distributed.run --nproc-per-node N tools/multi_gpu_configs.py [--frames 512]
Prints one JSON line per config on rank 0 (device time, max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from light_path_tracer_b200 import dist as lpdist, image_lens as il  # noqa: E402
from light_path_tracer_b200.metrics import Schwarzschild  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def timed(fn, reps):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    barrier()
    for a, b in ev:
        a.record(); fn(); b.record()
    barrier()
    return max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in ev])))


metric = Schwarzschild(1.0)
# ---------------- config 4 ----------------
H, W = 4320, 7680
yy, xx = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
r = (((yy // 32) + (xx // 32)) & 1).float()
src = torch.stack([r, 1 - r, xx.float() / W], dim=-1).contiguous()
del yy, xx, r
rs = lpdist.RowShardedRenderer(src, 40.0, metric)
row0, rows = rs.tiles[rank]
out = {}
for bands in (1, 2, 4, 8):
    if world > 1:
        g = lpdist.BandGather(rows, (W, 3), torch.float32, "cuda", dst=0, bands=bands)
        out["bands%d" % bands] = timed(lambda: rs.render_pipelined(100.0, gather=g), args.reps)
    else:
        tile = torch.empty((rows, W, 3), device="cuda")
        out["bands%d" % bands] = timed(lambda: rs.render_tile(100.0, out=tile), args.reps)
        break
tile = torch.empty((rows, W, 3), device="cuda")
out["render_only"] = timed(lambda: rs.render_tile(100.0, out=tile), args.reps)
if world > 1:
    frame = rs.render_pipelined(100.0, dst=0, bands=4)
    full = il.render_frame(src, rs.pipe.fov, 100.0, metric) if rank == 0 else None
    out["bit_identical_to_single_gpu_frame"] = bool(torch.equal(frame, full)) if rank == 0 else None
    # every rank's kernel stores its tile straight into rank 0's frame over NVLink (PeerFrame)
    try:
        pf = lpdist.PeerFrame(H, (W, 3), torch.float32, src.device, dst=0)
        out["bands_peer"] = timed(lambda: rs.render_peer(100.0, frame=pf), args.reps)
        got = rs.render_peer(100.0, frame=pf)
        out["peer_bit_identical"] = bool(torch.equal(got, full)) if rank == 0 else None
    except Exception as exc:                      # symmetric memory unavailable on this box
        out["peer_error"] = repr(exc)[:200]
if rank == 0:
    best = min(v for k, v in out.items() if k.startswith("bands"))
    print(json.dumps({"config": "4: 7680x4320 row-sharded x%d, gather to rank 0" % world, "ms_per_frame": out,
                      "rays_per_s_best": H * W / best * 1e3}), flush=True)
del src, rs, tile
# ---------------- config 5 ----------------
Hs = Ws = 1024
yy, xx = torch.meshgrid(torch.arange(Hs, device="cuda"), torch.arange(Ws, device="cuda"), indexing="ij")
r = (((yy // 32) + (xx // 32)) & 1).float()
src = torch.stack([r, 1 - r, xx.float() / Ws], dim=-1).contiguous()
pipe = il.LensPipeline(src, 40.0, metric)
grid = lpdist.sweep_grid()[: args.frames]
mine = lpdist.frame_shard(len(grid), rank, world)
frames = torch.empty((len(mine), Hs, Ws, 3), device="cuda")


def sweep():
    for j, k in enumerate(mine):
        r_obs, psi = grid[k]
        pipe.render(r_obs, psi=psi, out=frames[j])


ms = timed(sweep, 3)
check = frames[-1].clone()
graph = pipe.capture_sweep([grid[k] for k in mine], frames)
ms_graph = timed(graph.replay, 3)
same = bool(torch.equal(check, frames[-1]))
if rank == 0:
    print(json.dumps({"config": "5: %d-frame sweep 1024x1024, frame-sharded x%d" % (len(grid), world),
                      "ms_total": ms, "ms_per_frame": ms / len(grid), "frames_per_s": len(grid) / ms * 1e3,
                      "rays_per_s": len(grid) * Hs * Ws / ms * 1e3,
                      "cuda_graph": {"ms_total": ms_graph, "ms_per_frame": ms_graph / len(grid),
                                     "rays_per_s": len(grid) * Hs * Ws / ms_graph * 1e3,
                                     "same_frames": same}}), flush=True)
if world > 1:
    dist.destroy_process_group()
