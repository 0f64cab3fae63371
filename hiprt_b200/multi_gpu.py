"""Multi-GPU plumbing: one process per GPU, the BVH is replicated (every rank builds it
from the same input), work is sharded with NO data-path exchange:

  * image tiles: tile i of a row-major tiling goes to rank i mod world (interleaving balances sky / geometry tiles);
  * batch builds: the geometry list is split into contiguous slices;
  * the only collectives are the framebuffer gather to rank 0 and the timing reductions.

The reference is single-GPU (no collective call sites: SURVEY.md section 2); this module is the new part.
Works on any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def tile_grid(width, height, tile=64):
    return (width + tile - 1) // tile, (height + tile - 1) // tile


def tiles_of_rank(width, height, rank, world, tile=64):
    """Indices (row-major over the tile grid) of the tiles rank `rank` renders."""
    tx, ty = tile_grid(width, height, tile)
    return np.arange(rank, tx * ty, world, dtype=np.int64)


def tile_pixels(width, height, tile_index, tile=64):
    """Pixel indices (row-major over the image) of one tile, clipped to the image; tile-local row-major order."""
    tx, _ = tile_grid(width, height, tile)
    x0, y0 = (tile_index % tx) * tile, (tile_index // tx) * tile
    xs = np.arange(x0, min(x0 + tile, width), dtype=np.int64)
    ys = np.arange(y0, min(y0 + tile, height), dtype=np.int64)
    return (ys[:, None] * width + xs[None, :]).reshape(-1)


def pixels_of_rank(width, height, rank, world, tile=64):
    """Pixel indices of all tiles of a rank, tile after tile (each tile row-major): what the rank renders, in render order."""
    t = tiles_of_rank(width, height, rank, world, tile)
    if len(t) == 0:
        return np.zeros(0, np.int64)
    tx, _ = tile_grid(width, height, tile)
    if width % tile == 0 and height % tile == 0:  # all tiles full: one vectorised expression instead of a loop over tiles
        x0, y0 = (t % tx) * tile, (t // tx) * tile
        dy, dx = np.divmod(np.arange(tile * tile, dtype=np.int64), tile)
        return ((y0[:, None] + dy[None, :]) * width + x0[:, None] + dx[None, :]).reshape(-1)
    return np.concatenate([tile_pixels(width, height, int(i), tile) for i in t])


def split_batch(count, rank, world):
    """Contiguous slice [begin, end) of a list of `count` independent geometries for this rank."""
    base, extra = divmod(count, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _untile(tiles, width, height, tile):
    """[tiles, tile * tile, ...] in row-major tile order -> the row-major image [width * height, ...]: two views and one copy on
    the device, no index tensor (full tiles only)."""
    tx, ty = width // tile, height // tile
    rest = tuple(tiles.shape[2:])
    t = tiles.reshape((ty, tx, tile, tile) + rest)
    dims = (0, 2, 1, 3) + tuple(range(4, 4 + len(rest)))
    return t.permute(*dims).reshape((width * height,) + rest)


def gather_framebuffer(local_values, width, height, rank, world, tile=64, dist=None, dst=0):
    """Gather per-pixel values rendered by each rank (in pixels_of_rank order; any trailing shape/dtype torch supports)
    into the full row-major image on rank `dst`. Returns the image there, None elsewhere.
    When the tiles are full and deal out evenly (the usual case: 4096 x 4096 in 64 x 64 tiles on 1 / 2 / 4 / 8 ranks) the
    image is assembled from the gathered buffers by reshapes on the device: rank r's k-th tile is tile k * world + r.
    Otherwise (ragged borders, uneven deal) through index tensors."""
    import torch

    tx, ty = tile_grid(width, height, tile)
    even = width % tile == 0 and height % tile == 0 and (tx * ty) % world == 0
    rest = tuple(local_values.shape[1:])
    if world == 1:
        if even:
            return _untile(local_values.reshape((tx * ty, tile * tile) + rest), width, height, tile)
        out = torch.empty((width * height,) + rest, dtype=local_values.dtype, device=local_values.device)
        out[torch.from_numpy(pixels_of_rank(width, height, 0, 1, tile)).to(local_values.device)] = local_values
        return out
    if even:
        local_values = local_values.contiguous()
        bufs = [torch.empty_like(local_values) for _ in range(world)] if rank == dst else None
        dist.gather(local_values, bufs, dst=dst)
        if rank != dst:
            return None
        per_rank = (tx * ty) // world
        stacked = torch.stack(bufs).reshape((world, per_rank, tile * tile) + rest)  # [rank, k, pixel-in-tile, ...]
        tiles = stacked.transpose(0, 1).reshape((tx * ty, tile * tile) + rest)     # tile index = k * world + rank
        return _untile(tiles, width, height, tile)
    counts = [len(pixels_of_rank(width, height, r, world, tile)) for r in range(world)]
    most = max(counts)
    padded = torch.zeros((most,) + rest, dtype=local_values.dtype, device=local_values.device)
    padded[: len(local_values)] = local_values
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((width * height,) + rest, dtype=local_values.dtype, device=local_values.device)
    for r in range(world):
        idx = torch.from_numpy(pixels_of_rank(width, height, r, world, tile)).to(local_values.device)
        out[idx] = bufs[r][: counts[r]]
    return out
