"""CPU, gloo, world_size 2: the host-side sharding logic of the multi-GPU path
(light_path_tracer_b200/dist.py): row-tile partition, frame round-robin, tile gather."""
import os
import socket

import numpy as np
import pytest


def test_row_tiles_partition():
    from light_path_tracer_b200.dist import row_tiles, frame_shard, sweep_grid
    for H in (0, 1, 7, 2160, 4320, 4321):
        for G in (1, 2, 3, 4, 8):
            t = row_tiles(H, G)
            assert len(t) == G and t[0][0] == 0 and sum(r for _, r in t) == H
            assert all(t[k][0] + t[k][1] == t[k + 1][0] for k in range(G - 1))
            assert max(r for _, r in t) - min(r for _, r in t) <= 1
    with pytest.raises(ValueError):
        row_tiles(10, 0)
    all_frames = sorted(sum((frame_shard(512, r, 8) for r in range(8)), []))
    assert all_frames == list(range(512))
    assert len(frame_shard(512, 3, 8)) == 64
    from light_path_tracer_b200.dist import band_splits
    for rows in (0, 1, 5, 2160):
        for bands in (1, 4, 7, 5000):
            b = band_splits(rows, bands)
            assert sum(n for _, n in b) == rows and all(n > 0 for _, n in b)
            assert all(b[k][0] + b[k][1] == b[k + 1][0] for k in range(len(b) - 1))
    from light_path_tracer_b200.dist import band_layout, band_rows_of
    for H, G in ((4320, 8), (4320, 2), (2160 * 8, 8), (540, 2), (1080, 4), (96, 3)):
        b = band_layout(H, G)
        assert b is not None and 1 <= b <= 27 and H % (b * G) == 0
        owned = []
        for g in range(G):
            row0, rows, (br, stride), frame_rows = band_rows_of(H, g, G, b)
            assert row0 == g * b and rows == H // G and br == b and stride == b * G
            # the kernel's mapping (lp_trace.cuh tile_pixel): r -> row0 + (r // b) * stride + r % b
            r = np.arange(rows)
            assert np.array_equal(frame_rows, row0 + (r // b) * stride + (r % b))
            owned.append(frame_rows)
        assert np.array_equal(np.sort(np.concatenate(owned)), np.arange(H))
    assert band_layout(4321, 8) is None and band_layout(7, 2) is None and band_layout(8, 2, 3) == 2
    grid = sweep_grid()
    assert len(grid) == 512 and grid[0][0] == 15.0 and abs(grid[-1][0] - 1000.0) < 1e-9


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, height, q):
    import torch
    import torch.distributed as dist
    from light_path_tracer_b200.dist import row_tiles, gather_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        row0, rows = row_tiles(height, world)[rank]
        ok = True
        for dtype, shape in ((torch.float32, (5, 3)), (torch.uint8, (4,)), (torch.float64, ())):
            # the "render": every element encodes its frame row, so assembly errors show
            tile = (torch.arange(row0, row0 + rows, dtype=torch.float64).reshape((rows,) + (1,) * len(shape))
                    .expand((rows,) + shape) % 251).to(dtype).contiguous()
            full = gather_rows(tile, height)
            expect = (torch.arange(height, dtype=torch.float64).reshape((height,) + (1,) * len(shape))
                      .expand((height,) + shape) % 251).to(dtype)
            ok = ok and full.shape == expect.shape and torch.equal(full, expect)
            root = gather_rows(tile, height, dst=1 % world)
            if rank == 1 % world:
                ok = ok and torch.equal(root, expect)
            else:
                ok = ok and root is None
        # pipelined band gather (equal tiles only): same frame, root only
        if height % world == 0:
            from light_path_tracer_b200.dist import BandGather
            for bands in (1, 3, 64):
                g = BandGather(rows, (6, 3), torch.float32, "cpu", dst=0, bands=bands)
                for first, n in g.bands:
                    g.tile[first:first + n] = (torch.arange(row0 + first, row0 + first + n, dtype=torch.float32)
                                               .reshape(n, 1, 1).expand(n, 6, 3))
                    g.push(first, n)
                frame = g.finish()
                if rank == 0:
                    expect = torch.arange(height, dtype=torch.float32).reshape(height, 1, 1).expand(height, 6, 3)
                    ok = ok and torch.equal(frame, expect)
                else:
                    ok = ok and frame is None
        # RowShardedRenderer.render_to_root: path selection and caching with a stub renderer (the
        # real one needs a GPU): no CUDA -> no peer mapping -> pipelined bands when the tiles are
        # equal, whole-tile gather otherwise; the second call reuses the cached choice and buffers
        import types
        from light_path_tracer_b200.dist import RowShardedRenderer
        src = torch.zeros(height, 6, 3)
        rs = RowShardedRenderer.__new__(RowShardedRenderer)     # LensPipeline itself refuses to exist without CUDA
        rs.pipe = types.SimpleNamespace(src=src, height=height, width=6)
        rs.group, rs.world, rs.rank, rs.tiles = None, world, rank, row_tiles(height, world)

        def stub_render(r_obs, psi=(0.0, 0.0), rows=None, stats=None, flags=None, out=None):
            r0, n = rows
            val = (torch.arange(r0, r0 + n, dtype=torch.float32) + float(r_obs)).reshape(n, 1, 1).expand(n, 6, 3)
            if out is None:
                return val.contiguous()
            out.copy_(val)
            return out
        rs.pipe.render = stub_render
        for r_obs in (1.0, 2.0):
            frame = rs.render_to_root(r_obs, dst=0)
            expect = (torch.arange(height, dtype=torch.float32) + r_obs).reshape(height, 1, 1).expand(height, 6, 3)
            ok = ok and ((torch.equal(frame, expect)) if rank == 0 else (frame is None))
        ok = ok and rs._root_mode == ("bands" if height % world == 0 else "tiles")
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("height", [10, 11])
def test_gather_rows_gloo_world2(height):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, height, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
