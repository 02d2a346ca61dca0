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


def band_layout(height, world_size, target_rows=27):
    """Interleaved row bands for `world_size` ranks: the largest band height b <= target_rows with
    height % (b * world_size) == 0, so that rank g owns bands g, g + G, g + 2G, ... (b rows each,
    height / G rows in total).  Returns b, or None when the rows do not divide evenly (callers
    then use contiguous row_tiles).  The black hole sits in the centre rows of a frame, so
    contiguous tiles are unevenly expensive (config 4 at 8 GPUs: the slowest tile took 9 % longer
    than the mean); a few dozen bands per rank average that out."""
    if world_size < 1 or height <= 0 or height % world_size:
        return None
    per = height // world_size
    for b in range(min(int(target_rows), per), 0, -1):
        if per % b == 0:
            return b
    return None


def band_rows_of(height, rank, world_size, band_rows):
    """(row0, n_rows, (band_rows, band_stride)) of `rank` in the interleaved layout, and the frame
    rows it owns as a numpy index array (for host-side assembly and tests)."""
    per = height // world_size
    stride = band_rows * world_size
    r = np.arange(per)
    frame_rows = rank * band_rows + (r // band_rows) * stride + (r % band_rows)
    return rank * band_rows, per, (band_rows, stride), frame_rows


class PeerFrame:
    """The frame lives in rank `dst`'s HBM and every rank's render kernel stores its rows STRAIGHT
    INTO IT through NVLink peer memory (torch symmetric memory: CUDA VMM allocations mapped into
    every process of the group).  The "gather" is the kernel's own pixel stores — 16-byte vectors,
    see lp_render_kernel / lp_render_repack_kernel — so the transfer overlaps the FP64 work store
    by store; there is no copy kernel and NO collective on the data path or for completion: a
    one-thread kernel per rank raises an 8-byte epoch flag in rank dst's memory (lp_peer_signal)
    and rank dst's stream waits for all of them (lp_peer_wait).

        pf = PeerFrame(H, (W, 3), torch.uint8, device)          # collective (rendezvous), once
        for every frame:
            tile, rows, bands, fl = pf.begin()                  # every rank
            render(rows=rows, bands=bands, flags=flags | fl, out=tile)
            frame = pf.complete()                               # [H, W, 3] on dst, None elsewhere

    Ownership: the tensor `complete()` returns on dst is one of ``buffers`` (2) symmetric frame
    buffers and stays valid UNTIL THE NEXT complete() on dst; work that consumes it must be
    enqueued (same stream) before that call.  Peers never overwrite a buffer the root may still be
    reading: frame e goes to buffer e % buffers, and begin() makes the render wait (on the device)
    until the root has released frame e - buffers, which the root does — stream-ordered after its
    consumers — at the start of complete().  A peer can therefore run at most one frame ahead.
    Device-side waits give up after ``timeout_ms`` (a dead peer cannot hang the GPU); `timed_out()`
    reports it.

    ``band_rows``: interleave the ranks' rows in bands of that many rows (band_layout) instead of
    contiguous tiles.  Raises if symmetric memory is unavailable (callers fall back to BandGather /
    gather_rows); the availability check is agreed on with a collective BEFORE the rendezvous, so a
    one-sided failure cannot leave the other ranks waiting in it."""

    _DONE, _CONSUMED, _TIMEOUT = 0, 32, 40        # int64 slots of the flag pad

    def __init__(self, height, row_shape, dtype, device, dst=0, group=None, buffers=2, band_rows=None,
                 timeout_ms=10000):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.dist, self.dst = dist, dst
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 16:
            raise ValueError("PeerFrame serves one NVLink domain (<= 16 ranks)")
        self._ext = _lib.ext()
        self.height, self.row_shape = int(height), tuple(row_shape)
        self.buffers = max(1, int(buffers))
        self.timeout_ms = int(timeout_ms)
        shape = (self.buffers, self.height) + self.row_shape
        ok, err = 1.0, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            self.buf = symm_mem.empty(shape, dtype=dtype, device=device)
            self.pad = symm_mem.empty(64, dtype=torch.int64, device=device)
        except Exception as exc:                       # local allocation failed: tell everybody
            ok, err = 0.0, exc
        agree = torch.tensor([ok], device=device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
        if float(agree[0]) != 1.0:
            raise RuntimeError("symmetric memory unavailable on at least one rank (%r)" % (err,))
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.pad_handle = symm_mem.rendezvous(self.pad, self.group)
        self.pad.zero_()
        # load the two flag kernels NOW (CUDA loads kernels lazily and a load may wait for the device to drain:
        # it must never happen while a wait kernel is spinning on this GPU)
        self._ext.peer_signal([int(self.pad.data_ptr()) + 8 * 48], 1, self.pad)
        self._ext.peer_wait(self.pad[48:49], 1, 1, 1000, None)
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)                 # every pad is zero before anybody signals
        frame_numel = self.height * int(np.prod(self.row_shape))
        self.root = [self.handle.get_buffer(dst, (self.height,) + self.row_shape, dtype, b * frame_numel)
                     for b in range(self.buffers)]
        self.local = [self.buf[b] for b in range(self.buffers)]
        self._pad_ptrs = [int(p) for p in self.pad_handle.buffer_ptrs]
        if band_rows:
            if self.height % (int(band_rows) * self.world):
                raise ValueError("band_rows * world_size must divide the frame height")
            row0, rows, bands, _ = band_rows_of(self.height, self.rank, self.world, int(band_rows))
            self.rows, self.bands = (row0, rows), bands
        else:
            self.rows, self.bands = row_tiles(self.height, self.world)[self.rank], None
        self.epoch = 0
        # more than four writers saturate the root's NVLink ingress with the 8 x 4 tile's 8-byte stores (lightpath.h)
        self._run_flag = 0 if self.world <= 4 else 32       # _device.RENDER_ROW_RUNS

    def begin(self):
        """Start frame epoch+1: returns (tile tensor in rank dst's buffer, (row0, n_rows), bands,
        extra render flags).  Makes the current stream wait until the buffer is free."""
        from . import _device as dev
        self.epoch += 1
        e = self.epoch
        if e > self.buffers:
            self._ext.peer_wait(self.pad[self._CONSUMED:self._CONSUMED + 1], 1, e - self.buffers, self.timeout_ms,
                                self.pad[self._TIMEOUT:self._TIMEOUT + 1].view(self._i32()))
        root = self.root[e % self.buffers]
        row0, rows = self.rows
        if self.bands is None:
            return root[row0:row0 + rows], self.rows, None, dev.RENDER_STAGED_STORES | self._run_flag
        return root[row0:], self.rows, self.bands, dev.RENDER_STAGED_STORES | dev.RENDER_OUT_FRAME_ROWS | self._run_flag

    def complete(self):
        """Stream-ordered after this rank's render kernel: raise this rank's flag on dst; on dst,
        release the previous frame to the peers and wait for every rank's flag."""
        e = self.epoch
        self._ext.peer_signal([self._pad_ptrs[self.dst] + 8 * (self._DONE + self.rank)], e, self.pad)
        if self.rank != self.dst:
            return None
        if e >= 2:
            self._ext.peer_signal([p + 8 * self._CONSUMED for p in self._pad_ptrs], e - 1, self.pad)
        self._ext.peer_wait(self.pad[self._DONE:self._DONE + self.world], self.world, e, self.timeout_ms,
                            self.pad[self._TIMEOUT:self._TIMEOUT + 1].view(self._i32()))
        return self.local[e % self.buffers]

    def drain(self):
        """dst: release every frame handed out so far (call when done, before the buffers are
        reused by a new sequence or freed); then a barrier so that nobody tears the pads down early."""
        if self.rank == self.dst and self.epoch >= 1:
            self._ext.peer_signal([p + 8 * self._CONSUMED for p in self._pad_ptrs], self.epoch, self.pad)
        self.dist.barrier(group=self.group)

    def timed_out(self):
        """True if a device-side wait of this rank gave up (synchronises)."""
        return bool(int(self.pad[self._TIMEOUT].item()) != 0)

    def _i32(self):
        import torch
        return torch.int32


class ShardedHostFrames:
    """Host image in -> this rank's rows of the lensed frame out, for streams of frames, with the
    upload SHARDED and the source RESIDENT: every rank copies only its own 1/N of the rows of a
    NEW source image over its own PCIe link and one NCCL all-gather over NVLink assembles the
    replicated source on every GPU (the remap may sample any source pixel); a source that has not
    changed (same ``version``) is not uploaded or gathered again — sweeps over one background move
    only camera parameters in and tiles out.  The fused kernel renders the rank's rows; the tile
    goes back to (pinned) host memory.  ``depth`` slots on their own streams overlap one frame's
    upload with the previous frame's download, as image_lens.HostFramePipeline does on one GPU; two
    resident source buffers let the next image arrive while the current one is still being sampled.
    Needs equal row tiles.  ``unit_u8``: uint8 images standing for float32/255 — the 8-bit
    boundary of image_lens.main (image_lens.py:448-450, :510), 4x fewer bytes on PCIe than
    float32 frames."""

    def __init__(self, shape, dtype, vertical_fov=None, metric=None, depth=3, group=None, unit_u8=False,
                 band_rows=None):
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
        self.part_rows = tiles[self.rank]              # the rows of the SOURCE this rank uploads
        if band_rows:
            row0, rows, bands, self.frame_rows = band_rows_of(self.height, self.rank, self.world, int(band_rows))
            self.rows, self.bands = (row0, rows), bands
        else:
            self.rows, self.bands = tiles[self.rank], None
            self.frame_rows = np.arange(self.rows[0], self.rows[0] + self.rows[1])
        self.metric = metric if metric is not None else Schwarzschild(M=1.0)
        self.fov = vertical_fov
        self.unit_u8 = bool(unit_u8)
        device = torch.device("cuda", torch.cuda.current_device())
        tile_shape = (self.rows[1],) + self.shape[1:]
        self._slots = []
        for _ in range(max(1, int(depth))):
            st = torch.cuda.Stream(device=device)
            with torch.cuda.stream(st):
                self._slots.append(dict(stream=st, part=torch.empty(tile_shape, dtype=dtype, device=device),
                                        frame=torch.empty(tile_shape, dtype=dtype, device=device), done=None))
        # two resident, replicated sources: {buf, version, ready event, last-use events}
        self._res = [dict(buf=torch.empty(self.shape, dtype=dtype, device=device), version=None, ready=None, uses=[])
                     for _ in range(2)]
        self._cur = 0
        self._k = 0
        self.uploads = 0
        self._torch = torch

    def submit(self, host_src, fov, r_obs, psi=(0.0, 0.0), out=None, flags=None, version=None):
        """host_src: the full [H, W, ...] pinned host image (only this rank's rows are read, and
        only when ``version`` differs from the resident one; ``version=None`` = a new image every
        call).  Returns the pinned host tensor that holds this rank's tile after synchronize()."""
        from . import image_lens as il
        from . import _device as dev
        t = self._torch
        slot = self._slots[self._k % len(self._slots)]
        self._k += 1
        row0, n = self.rows
        if out is None:
            out = t.empty((n,) + self.shape[1:], dtype=slot["frame"].dtype).pin_memory()
        st = slot["stream"]
        st.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(st):
            res = self._res[self._cur]
            if version is None or res["version"] != version:
                res = self._res[1 - self._cur]
                for ev in res["uses"]:                 # frames still sampling the buffer we are about to overwrite
                    st.wait_event(ev)
                res["uses"] = []
                p0, pn = self.part_rows
                slot["part"].copy_(host_src[p0:p0 + pn], non_blocking=True)
                self.dist.all_gather_into_tensor(res["buf"].view(self.world, -1), slot["part"].view(-1),
                                                 group=self.group)
                res["version"] = version
                res["ready"] = t.cuda.Event()
                res["ready"].record(st)
                self._cur = 1 - self._cur
                self.uploads += 1
            else:
                st.wait_event(res["ready"])
            il.render_frame(res["buf"], fov, r_obs, self.metric, psi=psi, rows=(row0, n), bands=self.bands,
                            flags=dev.TRACE_HYBRID if flags is None else flags, out=slot["frame"],
                            unit_u8=self.unit_u8)
            used = t.cuda.Event()
            used.record(st)
            res["uses"] = res["uses"][-(len(self._slots) - 1):] + [used]
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
            # every rank takes the same branch here (`equal` is a function of the frame and the world
            # size), and PeerFrame agrees on availability with a collective before its rendezvous
            if equal and self.pipe.src.is_cuda:
                try:
                    self._root_frame = self._new_peer_frame(dst)
                    ok = 1.0
                except RuntimeError:                    # symmetric memory unavailable (agreed by all ranks)
                    self._root_frame = None
            mode = "peer" if ok == 1.0 else ("bands" if equal else "tiles")
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
                         self.pipe.src.dtype, self.pipe.src.device, dst=dst, group=self.group,
                         band_rows=band_layout(self.pipe.height, self.world))

    def render_peer(self, r_obs, psi=(0.0, 0.0), stats=None, flags=None, frame=None, dst=0, unit_u8=False):
        """Render this rank's rows directly into rank dst's frame over NVLink (PeerFrame).  Pass a
        PeerFrame to reuse its mapping across frames (creating one is a rendezvous).  The returned
        frame (dst only) is valid until the next call (PeerFrame's ownership rule)."""
        flags = self._default_flags() if flags is None else flags
        if frame is None:
            frame = self._new_peer_frame(dst)
        tile, rows, bands, extra = frame.begin()
        self.pipe.render(r_obs, psi=psi, rows=rows, bands=bands, stats=stats, flags=flags | extra, out=tile,
                         unit_u8=unit_u8)
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
