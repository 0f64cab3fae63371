"""The sharded primary + AO frame (hiprt_b200/frame.py, bench.py's step) on the GPU: image against the oracle's hits and
occlusion, the compact record formats against the full hiprtHit, and a world of 2 PROCESSES sharing ONE GPU -- gloo for the
rendezvous, the framebuffer exchange itself over a CUDA-IPC mapping with release / acquire flags exactly as on an 8-GPU box
(peer stores from hiprtB200ResolveAoFrame, hiprtB200WaitFlags on rank 0)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, SPP, TILE = 160, 96, 8, 16
EYE, LOOK = (0.5, 0.6, -0.6), (0.5, 0.0, 0.5)


def _scene():
    from hiprt_b200 import scenes

    return scenes.terrain_mesh(48)


def _oracle_image():
    from hiprt_b200 import frame
    from oracle import binding as orc

    verts, idx = _scene()
    g = orc.Geometry(verts, idx)
    px = frame.shard_pixels(W, H, 0, 1, TILE)
    prim = frame.frame_rays(W, H, EYE, LOOK, px)
    hits = g.trace(prim, tie_mode=1)
    ao, _ = frame.ao_rays_of(prim, hits, px, SPP)
    occluded = g.trace(ao, tie_mode=1)["primID"] != orc.INVALID
    return frame.reference_ao_image(W, H, px, hits, occluded, SPP)


@pytest.fixture(scope="module")
def ctx():
    from hiprt_b200 import api

    return api.Context(0)


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_single_gpu_frame_matches_the_oracle(ctx, exchange):
    from hiprt_b200 import frame

    verts, idx = _scene()
    t = ctx.torch
    geom = ctx.build(t.from_numpy(verts).to(ctx.device), t.from_numpy(idx.view(np.int32)).to(ctx.device))
    fr = frame.AoFrame(ctx, geom, W, H, EYE, LOOK, spp=SPP, tile=TILE, exchange=exchange)
    for _ in range(3):  # the flag protocol counts frames
        fr.step()
    img = fr.image()
    ref = _oracle_image()
    assert (ref >= 0).mean() > 0.3 and (ref[ref >= 0] < 1).mean() > 0.02  # terrain in view, some occlusion
    assert np.array_equal(img, ref)
    # the frame as one CUDA graph launch (peer-store exchange only): replays keep counting frames on the device
    assert fr.capture() == (exchange == "p2p")
    fr.framebuffer.fill_(-7.0)
    for _ in range(3):
        fr.step()
    assert np.array_equal(fr.image(), ref)
    fr.close()
    geom.destroy()


def test_record_formats_agree_with_the_full_hit(ctx):
    from hiprt_b200 import api, frame

    verts, idx = _scene()
    t = ctx.torch
    geom = ctx.build(t.from_numpy(verts).to(ctx.device), t.from_numpy(idx.view(np.int32)).to(ctx.device))
    px = frame.shard_pixels(W, H, 0, 1, TILE)
    rays = frame.frame_rays(W, H, EYE, LOOK, px)
    d_rays = ctx.rays_to_device(rays)
    full = ctx.hits_to_host(ctx.trace(geom, d_rays))
    assert (full["instanceIDs"] == api.INVALID).all()  # a geometry trace reports no instance at any level (all four words are written)
    for kind in (api.CLOSEST_HIT, api.ANY_HIT):
        ref = ctx.hits_to_host(ctx.trace(geom, d_rays, kind=kind))
        compact = t.empty((len(rays), 4), dtype=t.int32, device=ctx.device)
        ctx.trace_format(geom, d_rays, api.FORMAT_COMPACT, compact, kind)
        prim_only = t.empty((len(rays), 1), dtype=t.int32, device=ctx.device)
        ctx.trace_format(geom, d_rays, api.FORMAT_PRIM_ID, prim_only, kind)
        rec = compact.cpu().numpy().view(api.COMPACT_DTYPE).reshape(-1)
        hit = ref["primID"] != api.INVALID
        assert np.array_equal(hit, full["primID"] != api.INVALID)
        assert np.array_equal(prim_only.cpu().numpy().view(np.uint32).reshape(-1) != api.INVALID, hit)
        assert np.array_equal(rec["primID"] != api.INVALID, hit)
        if kind == api.CLOSEST_HIT:
            assert np.array_equal(rec["primID"], ref["primID"]) and np.array_equal(rec["t"][hit], ref["t"][hit])
            assert np.array_equal(rec["uv"][hit], ref["uv"][hit])
            assert np.array_equal(prim_only.cpu().numpy().view(np.uint32).reshape(-1), ref["primID"])
    geom.destroy()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _two_rank_worker(rank, world, port, exchange, out_path):
    import torch
    import torch.distributed as dist

    from hiprt_b200 import api, frame

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    ctx = api.Context(0)
    verts, idx = _scene()
    built = ctx.build(torch.from_numpy(verts).to(ctx.device), torch.from_numpy(idx.view(np.int32)).to(ctx.device)) if rank == 0 else None
    # the BVH reaches rank 1 as a broadcast blob + header rebase (bench.py's --replicate broadcast)
    handle, _ = frame.replicate_geometry(ctx, built, dist, src=0)
    geom = built if rank == 0 else api.Geometry(ctx, handle, None, None, api.FAST_BUILD)
    fr = frame.AoFrame(ctx, geom, W, H, EYE, LOOK, spp=SPP, rank=rank, world=world, tile=TILE, dist=dist, exchange=exchange)
    for _ in range(3):
        dist.barrier()
        fr.step()
        torch.cuda.synchronize()
    dist.barrier()
    if fr.exchange == "p2p":  # graph replays on both ranks: the flags keep counting
        assert fr.capture()
        if rank == 0:
            fr.framebuffer.fill_(-7.0)
        torch.cuda.synchronize()
        for _ in range(3):
            dist.barrier()
            fr.step()
            torch.cuda.synchronize()
    if rank == 0:
        np.save(out_path, fr.image())
        with open(out_path + ".exchange", "w") as f:
            f.write(fr.exchange)
    dist.barrier()
    fr.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["p2p"])  # (the NCCL gather variant needs one GPU per rank: bench.py --exchange nccl)
def test_two_ranks_on_one_gpu_deliver_the_same_frame(exchange, tmp_path):
    import torch.multiprocessing as mp

    out = str(tmp_path / "img.npy")
    mp.spawn(_two_rank_worker, args=(2, _free_port(), exchange, out), nprocs=2, join=True)
    img = np.load(out)
    assert np.array_equal(img, _oracle_image())
    if exchange == "p2p":
        assert open(out + ".exchange").read() == "p2p", "CUDA IPC was not available: the peer-store path did not run"
