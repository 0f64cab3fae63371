"""CPU tests of the multi-GPU host logic (tile sharding, batch split, framebuffer gather) with torch.distributed
on the gloo backend, world_size 2 and 3 (the GPU box runs the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hiprt_b200 import multi_gpu as mg


def test_tiles_partition_the_image():
    for (w, h, tile, world) in [(1920, 1080, 64, 8), (100, 70, 32, 3), (64, 64, 64, 2), (4096, 4096, 64, 8)]:
        seen = np.concatenate([mg.pixels_of_rank(w, h, r, world, tile) for r in range(world)])
        assert len(seen) == w * h and np.array_equal(np.sort(seen), np.arange(w * h))
        counts = [len(mg.tiles_of_rank(w, h, r, world, tile)) for r in range(world)]
        assert max(counts) - min(counts) <= 1


def test_split_batch_is_contiguous_and_complete():
    for count, world in [(100000, 8), (7, 3), (2, 4), (0, 2)]:
        slices = [mg.split_batch(count, r, world) for r in range(world)]
        assert slices[0][0] == 0 and slices[-1][1] == count
        assert all(a[1] == b[0] for a, b in zip(slices, slices[1:]))
        sizes = [e - b for b, e in slices]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, tile, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    px = mg.pixels_of_rank(w, h, rank, world, tile)
    # a stand-in for "trace my tiles": a deterministic function of the pixel index (hit distance + primitive id)
    local = torch.stack([torch.from_numpy(px).float() * 0.5, torch.from_numpy(px % 97).float()], -1)
    img = mg.gather_framebuffer(local, w, h, rank, world, tile, dist=dist)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the timing reduction bench.py does
    assert t.item() == world
    if rank == 0:
        np.save(out_path, img.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world, w, h, tile", [(2, 200, 136, 32), (3, 200, 136, 32), (2, 256, 128, 32), (4, 256, 128, 32)])
def test_framebuffer_gather_gloo(world, w, h, tile, tmp_path):
    # (ragged borders / an uneven deal go through index tensors, full tiles dealt evenly through reshapes on the device)
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(world, _free_port(), w, h, tile, out), nprocs=world, join=True)
    img = np.load(out)
    px = np.arange(w * h)
    assert np.array_equal(img[:, 0], (px * 0.5).astype(np.float32)) and np.array_equal(img[:, 1], (px % 97).astype(np.float32))


# ---- the sharded primary + AO frame (hiprt_b200/frame.py): host-side logic ----------------------------------------------
from hiprt_b200 import frame as fr  # noqa: E402
from hiprt_b200 import scenes  # noqa: E402


@pytest.mark.parametrize("shape", [(1920, 1080, 32, 8), (100, 70, 32, 3), (64, 40, 16, 2), (37, 19, 8, 4)])
def test_frame_shards_partition_the_image_and_keep_8x4_blocks(shape):
    w, h, tile, world = shape
    shards = [fr.shard_pixels(w, h, r, world, tile) for r in range(world)]
    seen = np.concatenate(shards)
    assert len(seen) == w * h and np.array_equal(np.sort(seen), np.arange(w * h))
    # a rank's first 32 rays cover one 8x4 pixel block (when the image has full blocks there)
    first = shards[0][:32]
    ys, xs = first // w, first % w
    if w >= 8 and h >= 4:
        assert ys.max() - ys.min() == 3 and xs.max() - xs.min() == 7
    # world = 1 reproduces the single-GPU trace order at the same tile size
    assert np.array_equal(np.sort(fr.shard_pixels(w, h, 0, 1, tile)), np.arange(w * h))


def test_ao_rays_do_not_depend_on_the_sharding():
    """AO rays are seeded by the GLOBAL pixel: the union of the shards' rays is the single-GPU frame's ray set."""
    from oracle import binding as orc

    w, h, spp = 96, 64, 4
    verts, idx = scenes.terrain_mesh(24)
    g = orc.Geometry(verts, idx)
    eye, look = (0.5, 0.6, -0.6), (0.5, 0.0, 0.5)

    def rays_of(rank, world):
        px = fr.shard_pixels(w, h, rank, world, 16)
        prim = fr.frame_rays(w, h, eye, look, px)
        hits = g.trace(prim, tie_mode=1)
        ao, first = fr.ao_rays_of(prim, hits, px, spp)
        hit = hits["primID"] != 0xFFFFFFFF
        assert np.array_equal(first[hit], np.arange(hit.sum()) * spp)
        return px[hit], ao.reshape(-1, spp)

    px1, ao1 = rays_of(0, 1)
    parts = [rays_of(r, 3) for r in range(3)]
    px3, ao3 = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    o1, o3 = np.argsort(px1), np.argsort(px3)
    assert np.array_equal(px1[o1], px3[o3]) and len(px1) > 500
    assert ao1[o1].tobytes() == ao3[o3].tobytes()


def _frame_worker(rank, world, port, w, h, tile, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = [len(fr.shard_pixels(w, h, r, world, tile)) for r in range(world)]
    px = fr.shard_pixels(w, h, rank, world, tile)
    local = torch.zeros(max(counts))
    local[: len(px)] = torch.from_numpy((px % 251).astype(np.float32)) / 251.0  # stand-in for the resolved AO of my pixels
    pieces = [torch.empty(max(counts)) for _ in range(world)] if rank == 0 else None
    piece_px = [torch.from_numpy(fr.shard_pixels(w, h, r, world, tile)) for r in range(world)] if rank == 0 else None
    fb = torch.full((w * h,), -2.0) if rank == 0 else None
    fr.gather_shards(dist, rank, local, pieces, counts, piece_px, fb)
    if rank == 0:
        np.save(out_path, fb.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_frame_gather_variant_gloo(world, tmp_path):
    w, h, tile = 200, 136, 32
    out = str(tmp_path / "fb.npy")
    mp.spawn(_frame_worker, args=(world, _free_port(), w, h, tile, out), nprocs=world, join=True)
    assert np.array_equal(np.load(out), ((np.arange(w * h) % 251).astype(np.float32) / np.float32(251.0)))


def test_reference_ao_image():
    hits = np.zeros(4, fr.api.HIT_DTYPE)
    hits["primID"] = [5, 0xFFFFFFFF, 7, 9]
    occluded = np.array([1, 1, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1], bool)  # 3 hit pixels x 4 spp
    img = fr.reference_ao_image(2, 2, np.array([3, 0, 1, 2]), hits, occluded, 4)
    assert img.reshape(-1).tolist() == [-1.0, 0.75, 0.0, 0.5]
