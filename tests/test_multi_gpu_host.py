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


@pytest.mark.parametrize("world", [2, 3])
def test_framebuffer_gather_gloo(world, tmp_path):
    w, h, tile = 200, 136, 32
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(world, _free_port(), w, h, tile, out), nprocs=world, join=True)
    img = np.load(out)
    px = np.arange(w * h)
    assert np.array_equal(img[:, 0], (px * 0.5).astype(np.float32)) and np.array_equal(img[:, 1], (px % 97).astype(np.float32))
