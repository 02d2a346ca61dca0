"""Multi-GPU sharding of the lensing path (one process per GPU, torch.distributed).

Rays are independent (reference metrics.py:664-668: disjoint ``out_fa[i]``), so there is NO
exchange during compute.  Two partitionings:

* **row tiles** of one frame: rank g renders rows ``[H*g//G, H*(g+1)//G)`` from the camera
  parameters alone (nothing is scattered; the read-only source image is replicated), then the
  finished tiles are gathered with one NCCL collective (``gather_rows``);
* **frames** of a parameter sweep: rank g renders frames ``g, g+G, ...`` (``frame_shard``); no
  data-path collective at all.

The partition / assembly logic is backend-agnostic and is covered on CPU with gloo,
world_size 2 (tests/test_dist_cpu.py); the render itself needs CUDA.
"""
import numpy as np


def row_tiles(height, world_size):
    """Contiguous, balanced row tiles: [(row0, rows)] * world_size, sizes differ by <= 1."""
    if world_size < 1 or height < 0:
        raise ValueError("bad partition request")
    bounds = [height * k // world_size for k in range(world_size + 1)]
    return [(bounds[k], bounds[k + 1] - bounds[k]) for k in range(world_size)]


def frame_shard(n_frames, rank, world_size):
    """Frame indices of `rank` in a sweep of n_frames (round-robin: neighbouring sweep points
    have similar cost, so interleaving balances the ranks)."""
    return list(range(rank, n_frames, world_size))


def gather_rows(tile, height, group=None, dst=None):
    """Assemble the full frame from per-rank row tiles.

    tile: this rank's [rows_g, W, ...] tensor (CPU with gloo, CUDA with NCCL).
    dst=None -> every rank gets the frame (all_gather); dst=r -> only rank r does (gather),
    the others return None.  Tiles may differ by one row; they are padded to the largest
    tile for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    tiles = row_tiles(height, world)
    if tuple(tile.shape[:1]) != (tiles[rank][1],):
        raise ValueError("tile has %d rows, rank %d owns %d" % (tile.shape[0], rank, tiles[rank][1]))
    if world == 1:
        return tile
    max_rows = max(r for _, r in tiles)
    row_shape = tuple(tile.shape[1:])
    if tile.shape[0] == max_rows:
        send = tile.contiguous()
    else:
        send = torch.zeros((max_rows,) + row_shape, dtype=tile.dtype, device=tile.device)
        send[:tile.shape[0]] = tile
    # collectives move bytes: view as uint8 so that every dtype (incl. uint16) is accepted
    send_b = send.view(torch.uint8).reshape(-1)
    if dst is None:
        buf = torch.empty((world, send_b.numel()), dtype=torch.uint8, device=tile.device)
        dist.all_gather_into_tensor(buf.reshape(-1), send_b, group=group)
    else:
        if rank == dst:
            parts = [torch.empty_like(send_b) for _ in range(world)]
            dist.gather(send_b, parts, dst=dst, group=group)
            buf = torch.stack(parts)
        else:
            dist.gather(send_b, None, dst=dst, group=group)
            return None
    buf = buf.view(tile.dtype).reshape((world, max_rows) + row_shape)
    if all(r == max_rows for _, r in tiles):
        return buf.reshape((height,) + row_shape)
    return torch.cat([buf[g, :tiles[g][1]] for g in range(world)], dim=0)


class RowShardedRenderer:
    """Row-tile sharded lensed render (BASELINE config 4): each rank renders its tile with the
    fused kernel, then the frame is gathered over NCCL / NVLink."""

    def __init__(self, source_image, vertical_fov_deg=40.0, metric=None, group=None):
        import torch.distributed as dist
        from .image_lens import LensPipeline
        self.pipe = LensPipeline(source_image, vertical_fov_deg, metric)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.tiles = row_tiles(self.pipe.height, self.world)

    def render_tile(self, r_obs, psi=(0.0, 0.0), stats=None, flags=0, out=None):
        return self.pipe.render(r_obs, psi=psi, rows=self.tiles[self.rank], stats=stats, flags=flags, out=out)

    def render(self, r_obs, psi=(0.0, 0.0), dst=None, stats=None, flags=0):
        tile = self.render_tile(r_obs, psi, stats, flags)
        if self.world == 1:
            return tile
        return gather_rows(tile, self.pipe.height, self.group, dst)


def sweep_grid(n_r=32, n_psi=16, r_lo=15.0, r_hi=1000.0, psi_deg=15.0):
    """The 512-frame sweep of BASELINE config 5: r_obs in geomspace(15, 1000) x
    camera pitch psi_y in linspace(-15 deg, +15 deg); returns [(r_obs, (psi_y, psi_x))]."""
    rs = np.geomspace(r_lo, r_hi, n_r)
    ps = np.radians(np.linspace(-psi_deg, psi_deg, n_psi))
    return [(float(r), (float(p), 0.0)) for r in rs for p in ps]
