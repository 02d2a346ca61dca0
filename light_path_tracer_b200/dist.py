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


def band_splits(rows, bands):
    """Split a tile of `rows` rows into at most `bands` contiguous bands: [(first_row, n_rows)]."""
    bands = max(1, min(int(bands), max(rows, 1)))
    edges = [rows * k // bands for k in range(bands + 1)]
    return [(edges[k], edges[k + 1] - edges[k]) for k in range(bands) if edges[k + 1] > edges[k]]


class BandGather:
    """Gather-to-root of equal-sized row tiles, pipelined with the render: the tile is produced
    band by band and every finished band is handed to NCCL (async) while the next band is
    still being rendered, so the NVLink transfer of the frame overlaps the FP64 work instead of
    following it (at 8 GPUs the gather is the same order as the compute).

        g = BandGather(rows, row_shape, dtype, device, dst=0, bands=4)
        for first, n in g.bands:
            render(rows=(row0 + first, n), out=g.tile[first:first + n])
            g.push(first, n)
        frame = g.finish()        # [world * rows, ...] on dst, None elsewhere

    Works with any backend (gloo on CPU for the tests).  All ranks must own the same number of
    rows (weak scaling, or H divisible by the world size)."""

    def __init__(self, rows, row_shape, dtype, device, dst=0, bands=4, group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.group, self.dst = dist, group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows = rows
        self.tile = torch.empty((rows,) + tuple(row_shape), dtype=dtype, device=device)
        self.frame = (torch.empty((self.world, rows) + tuple(row_shape), dtype=dtype, device=device)
                      if self.rank == dst else None)
        self.bands = band_splits(rows, bands)
        self._works = []

    def push(self, first, n):
        send = self.tile[first:first + n]
        if self.rank == self.dst:
            parts = [self.frame[g, first:first + n] for g in range(self.world)]
            w = self.dist.gather(send, parts, dst=self.dst, group=self.group, async_op=True)
        else:
            w = self.dist.gather(send, None, dst=self.dst, group=self.group, async_op=True)
        self._works.append(w)

    def finish(self):
        for w in self._works:
            w.wait()
        self._works = []
        if self.frame is None:
            return None
        return self.frame.reshape((self.world * self.rows,) + tuple(self.frame.shape[2:]))


class PeerFrame:
    """The frame lives in rank `dst`'s HBM and every rank's render kernel stores its row tile
    STRAIGHT INTO IT through NVLink peer memory (torch symmetric memory: CUDA VMM allocations
    mapped into every process of the group).  The "gather" is the kernel's own pixel stores —
    16-byte vectors, see lp_render_kernel — so the transfer overlaps the FP64 work store by
    store and there is no copy kernel and no data-path collective; one tiny all-reduce orders
    "every rank's kernel has finished" before rank `dst` reads the frame.

        pf = PeerFrame(H, (W, 3), torch.float32, device)       # collective (rendezvous)
        render(rows=pf.rows, out=pf.tile)                      # every rank, its own tile
        frame = pf.complete()                                  # [H, W, 3] on dst, None elsewhere

    Raises if symmetric memory is unavailable (callers fall back to BandGather / gather_rows)."""

    def __init__(self, height, row_shape, dtype, device, dst=0, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.dist, self.dst = dist, dst
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        shape = (height,) + tuple(row_shape)
        self.buf = symm_mem.empty(shape, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.root = self.handle.get_buffer(dst, shape, dtype)
        self.rows = row_tiles(height, self.world)[self.rank]
        self.tile = self.root[self.rows[0]:self.rows[0] + self.rows[1]]
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)

    def complete(self):
        # stream-ordered after this rank's render kernel; finishes on dst only after every
        # rank's kernel (and therefore its peer stores) has finished
        self.dist.all_reduce(self._flag, group=self.group)
        return self.buf if self.rank == self.dst else None


class ShardedHostFrames:
    """Host image in -> this rank's row tile of the lensed frame out, for streams of frames, with
    the upload SHARDED: every rank copies only its own 1/N of the rows of the source image over
    its own PCIe link, one NCCL all-gather over NVLink assembles the replicated source on every
    GPU (the remap may sample any source pixel), the fused kernel renders the rank's tile, the
    tile goes back to (pinned) host memory.  PCIe carries 1/N of the source per GPU instead of
    all of it; ``depth`` slots on their own streams overlap one frame's upload with the previous
    frame's download, as image_lens.HostFramePipeline does on one GPU.  Needs equal row tiles."""

    def __init__(self, shape, dtype, vertical_fov=None, metric=None, depth=3, group=None):
        import torch
        import torch.distributed as dist
        from .metrics import Schwarzschild
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.shape = tuple(shape)
        self.height, self.width = self.shape[0], self.shape[1]
        tiles = row_tiles(self.height, self.world)
        if any(r != tiles[0][1] for _, r in tiles):
            raise ValueError("ShardedHostFrames needs equal row tiles")
        self.rows = tiles[self.rank]
        self.metric = metric if metric is not None else Schwarzschild(M=1.0)
        self.fov = vertical_fov
        device = torch.device("cuda", torch.cuda.current_device())
        tile_shape = (self.rows[1],) + self.shape[1:]
        self._slots = [dict(stream=torch.cuda.Stream(device=device),
                            part=torch.empty(tile_shape, dtype=dtype, device=device),
                            src=torch.empty(self.shape, dtype=dtype, device=device),
                            frame=torch.empty(tile_shape, dtype=dtype, device=device))
                       for _ in range(max(1, int(depth)))]
        self._k = 0
        self._torch = torch

    def submit(self, host_src, fov, r_obs, psi=(0.0, 0.0), out=None, flags=None):
        """host_src: the full [H, W, ...] pinned host image (only this rank's rows are read).
        Returns the pinned host tensor that holds this rank's tile after synchronize()."""
        from . import image_lens as il
        from . import _device as dev
        t = self._torch
        slot = self._slots[self._k % len(self._slots)]
        self._k += 1
        row0, n = self.rows
        if out is None:
            out = t.empty((n,) + self.shape[1:], dtype=slot["src"].dtype).pin_memory()
        st = slot["stream"]
        st.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(st):
            slot["part"].copy_(host_src[row0:row0 + n], non_blocking=True)
            self.dist.all_gather_into_tensor(slot["src"].view(self.world, -1), slot["part"].view(-1),
                                             group=self.group)
            il.render_frame(slot["src"], fov, r_obs, self.metric, psi=psi, rows=(row0, n),
                            flags=dev.TRACE_HYBRID if flags is None else flags, out=slot["frame"])
            out.copy_(slot["frame"], non_blocking=True)
        return out

    def synchronize(self):
        cur = self._torch.cuda.current_stream()
        for slot in self._slots:
            cur.wait_stream(slot["stream"])
        cur.synchronize()


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

    def render_tile(self, r_obs, psi=(0.0, 0.0), stats=None, flags=None, out=None):
        flags = self._default_flags() if flags is None else flags
        return self.pipe.render(r_obs, psi=psi, rows=self.tiles[self.rank], stats=stats, flags=flags, out=out)

    def render(self, r_obs, psi=(0.0, 0.0), dst=None, stats=None, flags=None):
        flags = self._default_flags() if flags is None else flags
        tile = self.render_tile(r_obs, psi, stats, flags)
        if self.world == 1:
            return tile
        return gather_rows(tile, self.pipe.height, self.group, dst)

    def render_pipelined(self, r_obs, psi=(0.0, 0.0), dst=0, bands=4, stats=None, flags=None, gather=None):
        """Render this rank's tile band by band and gather each band while the next one is
        being rendered (BandGather).  Needs equal tiles (H divisible by the world size).
        Pass a BandGather to reuse its buffers across frames."""
        flags = self._default_flags() if flags is None else flags
        row0, rows = self.tiles[self.rank]
        if any(r != rows for _, r in self.tiles):
            raise ValueError("render_pipelined needs equal row tiles")
        if gather is None:
            gather = BandGather(rows, (self.pipe.width,) + tuple(self.pipe.src.shape[2:]), self.pipe.src.dtype,
                                self.pipe.src.device, dst=dst, bands=bands, group=self.group)
        for first, n in gather.bands:
            self.pipe.render(r_obs, psi=psi, rows=(row0 + first, n), stats=stats, flags=flags,
                             out=gather.tile[first:first + n])
            gather.push(first, n)
        return gather.finish()

    def render_to_root(self, r_obs, psi=(0.0, 0.0), dst=0, stats=None, flags=None):
        """The frame on rank ``dst`` (None elsewhere) by the fastest path this job supports, chosen
        once and cached: peer-memory stores (render_peer; 8K frame on 8 GPUs 0.53 ms) when every
        rank could map the symmetric frame, else the pipelined NCCL band gather (0.88 ms), else
        (ragged tiles) one gather of whole tiles.  Collective: every rank calls it."""
        if self.world == 1:
            return self.render_tile(r_obs, psi, stats, flags)
        import torch
        import torch.distributed as dist
        mode = getattr(self, "_root_mode", None)
        if mode is None or self._root_dst != dst:
            equal = all(r == self.tiles[0][1] for _, r in self.tiles)
            self._root_dst, self._root_frame, self._root_gather = dst, None, None
            ok = 0.0
            if equal and self.pipe.src.is_cuda:
                try:
                    self._root_frame = self._new_peer_frame(dst)
                    ok = 1.0
                except Exception:                       # symmetric memory unavailable
                    self._root_frame = None
            agree = torch.tensor([ok], device=self.pipe.src.device)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
            if float(agree[0]) == 1.0:
                mode = "peer"
            else:
                self._root_frame = None
                mode = "bands" if equal else "tiles"
            self._root_mode = mode
        if mode == "peer":
            return self.render_peer(r_obs, psi, stats, flags, frame=self._root_frame, dst=dst)
        if mode == "bands":
            if self._root_gather is None:
                rows = self.tiles[self.rank][1]
                self._root_gather = BandGather(rows, (self.pipe.width,) + tuple(self.pipe.src.shape[2:]),
                                               self.pipe.src.dtype, self.pipe.src.device, dst=dst, bands=4,
                                               group=self.group)
            return self.render_pipelined(r_obs, psi, dst=dst, stats=stats, flags=flags, gather=self._root_gather)
        return self.render(r_obs, psi, dst=dst, stats=stats, flags=flags)

    def _new_peer_frame(self, dst):
        return PeerFrame(self.pipe.height, (self.pipe.width,) + tuple(self.pipe.src.shape[2:]),
                         self.pipe.src.dtype, self.pipe.src.device, dst=dst, group=self.group)

    def render_peer(self, r_obs, psi=(0.0, 0.0), stats=None, flags=None, frame=None, dst=0):
        """Render this rank's tile directly into rank dst's frame over NVLink (PeerFrame).  Pass a
        PeerFrame to reuse its mapping across frames (creating one is a rendezvous)."""
        flags = self._default_flags() if flags is None else flags
        if frame is None:
            frame = self._new_peer_frame(dst)
        from . import _device as dev
        self.pipe.render(r_obs, psi=psi, rows=frame.rows, stats=stats, flags=flags | dev.RENDER_STAGED_STORES,
                         out=frame.tile)
        return frame.complete()

    @staticmethod
    def _default_flags():
        from . import _device as dev
        return dev.TRACE_HYBRID


def sweep_grid(n_r=32, n_psi=16, r_lo=15.0, r_hi=1000.0, psi_deg=15.0):
    """The 512-frame sweep of BASELINE config 5: r_obs in geomspace(15, 1000) x
    camera pitch psi_y in linspace(-15 deg, +15 deg); returns [(r_obs, (psi_y, psi_x))]."""
    rs = np.geomspace(r_lo, r_hi, n_r)
    ps = np.radians(np.linspace(-psi_deg, psi_deg, n_psi))
    return [(float(r), (float(p), 0.0)) for r in rs for p in ps]
