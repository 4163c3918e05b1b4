"""N > 1 host path on CPU: two gloo ranks each render their interleaved row tiles with the
oracle, exchange them with one all_gather and must reproduce the single-rank frame exactly."""
import importlib
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from _util import oracle_lib, params, small_config, vp

sharding = importlib.import_module("raytracer-voxpopuli_b200.sharding")


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = oracle_lib()
        sc, cam, fpd = small_config(4, 64, 48, spp=2)
        with vp.Session(lib) as s:
            s.set_scene(sc)
            acc, rgb8, st = s.render(cam, params(fpd, **sharding.shard_params(world, rank)))
        frame = torch.from_numpy(rgb8.astype(np.int32))
        accum = torch.from_numpy(acc)
        sharding.FrameGather(frame, world, rank)(frame)
        sharding.FrameGather(accum, world, rank)(accum)
        rays = torch.tensor([st.rays], dtype=torch.int64)
        dist.all_reduce(rays)
        if rank == 0:
            q.put((frame.numpy().copy(), accum.numpy().copy(), int(rays[0])))
    finally:
        dist.destroy_process_group()


def test_two_rank_tile_sharding_matches_single_rank():
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    frame, accum, rays = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc, cam, fpd = small_config(4, 64, 48, spp=2)
    with vp.Session(oracle_lib()) as s:
        s.set_scene(sc)
        want_acc, want_rgb8, st = s.render(cam, params(fpd))
    assert np.array_equal(frame.astype(np.uint32), want_rgb8)
    assert np.array_equal(accum.view(np.uint32), want_acc.view(np.uint32))
    assert rays == st.rays


def _reproject_cams(W, H, cam):
    from _util import host
    pos = np.array(cam.cam_pos, np.float32)
    tl, tr, bl = (np.array(v, np.float32) for v in (cam.top_left, cam.top_right, cam.bottom_left))
    target = tl + (tr - tl) * np.float32(0.5) + (bl - tl) * np.float32(0.5)
    cams = [host.look_at(pos + np.array([0.02 * f, 0, 0], np.float32), target, W, H) for f in range(2)]
    return cams, host.prev_camera(pos, target, W, H)


def _worker_reproject(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = oracle_lib()
        sc, cam, fpd = small_config(4, 64, 48)
        cams, prev = _reproject_cams(64, 48, cam)
        hist = None
        with vp.Session(lib) as s:
            s.set_scene(sc)
            for f in range(2):
                h, rgb8, _, _, st = s.render_reproject(cams[f], prev, params(fpd, frame_index=f, **sharding.shard_params(world, rank)), hist)
                ht = torch.from_numpy(h)
                sharding.FrameGather(ht, world, rank)(ht)  # every rank needs the WHOLE history for the next frame
                hist = ht.numpy().copy()
        frame = torch.from_numpy(rgb8.astype(np.int32))
        sharding.FrameGather(frame, world, rank)(frame)
        if rank == 0:
            q.put((frame.numpy().copy(), hist))
    finally:
        dist.destroy_process_group()


def test_two_rank_reprojection_matches_single_rank():
    """Reprojection frame mode sharded over two gloo ranks: halo rows in pass 1, history gathered between frames."""
    world, port = 2, 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_reproject, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    frame, hist = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc, cam, fpd = small_config(4, 64, 48)
    cams, prev = _reproject_cams(64, 48, cam)
    with vp.Session(oracle_lib()) as s:
        s.set_scene(sc)
        want = None
        for f in range(2):
            want, want_rgb8, _, _, _ = s.render_reproject(cams[f], prev, params(fpd, frame_index=f), want)
    assert np.array_equal(frame.astype(np.uint32), want_rgb8)
    assert np.array_equal(hist.view(np.uint32), want.view(np.uint32))


def test_tile_ownership_partitions_the_frame():
    for H in (48, 1080, 2160):
        for world in (1, 2, 4, 8):
            owned = sorted(t for r in range(world) for t in sharding.tiles_of_rank(H, world, r))
            assert owned == list(range((H + sharding.ROW_TILE - 1) // sharding.ROW_TILE))
            sizes = [len(sharding.tiles_of_rank(H, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_frame_gather_deinterleave_ragged(monkeypatch):
    """FrameGather's de-interleave (one permuted copy + the ranks that own one tile more) for tile counts that do not
    divide by the number of ranks; the collective is replaced by a function that fills the gathered buffer."""
    import torch
    import torch.distributed as dist
    sharding = importlib.import_module("raytracer-voxpopuli_b200.sharding")
    for world, ntile in ((3, 7), (8, 135), (2, 4), (4, 5), (8, 3)):
        H, W = ntile * sharding.ROW_TILE, 5
        truth = torch.arange(H * W, dtype=torch.int32).view(H, W)
        tv = truth.view(ntile, sharding.ROW_TILE, W)
        for rank in range(world):
            frame = torch.zeros((H, W), dtype=torch.int32)
            frame.view(ntile, sharding.ROW_TILE, W)[rank::world] = tv[rank::world]
            g = sharding.FrameGather(frame, world, rank)

            def fake(out, inp, g=g, world=world):
                o = out.view(world, g.max_tiles, sharding.ROW_TILE, W)
                for r in range(world):
                    o[r].zero_()
                    o[r, : g.counts[r]] = tv[r::world]
            monkeypatch.setattr(dist, "all_gather_into_tensor", fake)
            assert torch.equal(g(frame), truth), (world, ntile, rank)
