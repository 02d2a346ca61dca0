"""Multi-GPU data path on real devices (needs >= 2 GPUs; skipped otherwise): a frame rendered
as row tiles by two ranks — NCCL band gather and NVLink peer-memory stores — must be
bit-identical to the single-GPU frame (SURVEY.md §8e correctness test)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from light_path_tracer_b200 import dist as lpdist, image_lens as il
    from light_path_tracer_b200.metrics import Schwarzschild
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        H, W = 540, 960
        g = torch.Generator().manual_seed(3)
        src = torch.rand(H, W, 3, generator=g).cuda()
        metric = Schwarzschild(1.0)
        rs = lpdist.RowShardedRenderer(src, 40.0, metric)
        full = il.render_frame(src, rs.pipe.fov, 30.0, metric, psi=(0.05, -0.02))
        res = {}
        a = rs.render(30.0, psi=(0.05, -0.02), dst=0)
        b = rs.render_pipelined(30.0, psi=(0.05, -0.02), dst=0, bands=3)
        try:
            c = rs.render_peer(30.0, psi=(0.05, -0.02), dst=0)
            peer_ok = True
        except Exception as exc:                      # symmetric memory not available on this box
            c, peer_ok = None, repr(exc)
        if rank == 0:
            res = {"gather": bool(torch.equal(a, full)), "bands": bool(torch.equal(b, full)),
                   "peer": (bool(torch.equal(c, full)) if peer_ok is True else peer_ok)}
        # the auto path (peer stores, else band gather), twice: the second call reuses the mapping
        d1 = rs.render_to_root(30.0, psi=(0.05, -0.02), dst=0)
        d2 = rs.render_to_root(30.0, psi=(0.05, -0.02), dst=0)
        if rank == 0:
            res["to_root"] = bool(torch.equal(d1, full)) and bool(torch.equal(d2, full))
            res["to_root_mode"] = rs._root_mode
        else:
            res["to_root"] = (d1 is None and d2 is None)
        every = rs.render(30.0, psi=(0.05, -0.02), dst=None)      # all_gather: every rank gets the frame
        res["all"] = bool(torch.equal(every, full))
        # STREAMING through one double-buffered PeerFrame: a different camera every frame and no
        # host synchronisation in between, so a peer that overwrote a buffer the root had not
        # consumed yet (write-after-read) would tear a frame.  uint8 image boundary, interleaved bands.
        src8 = (src * 255).to(torch.uint8)
        pipe8 = il.LensPipeline(src8, 40.0, metric)
        cams = [(15.0 + 7.0 * j, (0.01 * j, -0.02 * j)) for j in range(7)]
        try:
            pf = lpdist.PeerFrame(H, (W, 3), torch.uint8, src.device, dst=0, band_rows=lpdist.band_layout(H, world, 9))
            kept = []
            for r_obs, psi in cams:
                tile, rows, bands, extra = pf.begin()
                pipe8.render(r_obs, psi=psi, rows=rows, bands=bands, out=tile, unit_u8=True,
                             flags=4 | extra)
                fr = pf.complete()
                if rank == 0:
                    kept.append(fr.clone())          # consumer, stream-ordered before the next complete()
            pf.drain()
            if rank == 0:
                ok = True
                for (r_obs, psi), got in zip(cams, kept):
                    ok = ok and bool(torch.equal(got, pipe8.render(r_obs, psi=psi, unit_u8=True)))
                res["peer_stream_u8"] = ok and not pf.timed_out()
                res["peer_bands"] = pf.bands
            else:
                res["peer_stream_u8"] = not pf.timed_out()
        except RuntimeError as exc:
            res["peer_stream_u8"] = repr(exc)
        # host image in / host tile out with the upload sharded over the ranks; a source with an
        # unchanged version is not uploaded again
        host_src = src.cpu().pin_memory()
        for band_rows in (None, lpdist.band_layout(H, world, 9)):
            pipe = lpdist.ShardedHostFrames((H, W, 3), torch.float32, metric=metric, depth=2, band_rows=band_rows)
            outs = [pipe.submit(host_src, rs.pipe.fov, 30.0, psi=(0.05, -0.02)) for _ in range(3)]
            outs += [pipe.submit(host_src, rs.pipe.fov, 30.0, psi=(0.05, -0.02), version=7) for _ in range(3)]
            pipe.synchronize()
            idx = torch.from_numpy(pipe.frame_rows)
            key = "sharded_host" if band_rows is None else "sharded_host_bands"
            res[key] = all(bool(torch.equal(o, full.cpu()[idx])) for o in outs) and pipe.uploads == 4
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_two_rank_frame_bit_identical(native):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0]["gather"] and results[0]["bands"] and results[0]["all"] and results[1]["all"]
    assert results[0]["sharded_host"] and results[1]["sharded_host"]
    assert results[0]["sharded_host_bands"] and results[1]["sharded_host_bands"]
    assert results[0]["peer_stream_u8"] is True and results[1]["peer_stream_u8"] is True, results
    assert results[0]["peer_bands"] is not None
    assert results[0]["to_root"] and results[1]["to_root"]
    assert results[0]["to_root_mode"] in ("peer", "bands")
    assert results[0]["peer"] is True, results[0]["peer"]
