"""One primary + ambient-occlusion frame of ONE job sharded over the GPUs of a box (north_star
"ray batches and image tiles are sharded across GPUs ... gather the framebuffer").

  * the BVH is replicated: built once and broadcast (`replicate_geometry`: ncclBroadcast of the blob + hiprtB200RelocateGeometry)
    or rebuilt on every rank (deterministic);
  * pixels are sharded in tiles, tile i -> rank i mod world (interleaving balances sky and terrain), inside a tile in the
    8x4 sub-tile order the single-GPU frame uses, so a warp's 32 rays cover the same pixel block at every world size;
  * a step on a rank = closest-hit launch over its primary rays, any-hit launch over its AO rays, then the resolve kernel
    (`hiprtB200ResolveAoFrame`) that turns the occlusion words into one float per pixel and stores it straight into the
    framebuffer that lives on rank 0 -- peer stores over NVLink from inside the kernel, completion by a release flag per rank
    that rank 0 waits on (`hiprtB200WaitFlags`). `exchange="nccl"` keeps the resolve local and gathers with dist.gather instead.

The sharding and ray-generation logic is plain numpy (tested on the CPU with a gloo world of 2: tests/test_multi_gpu_host.py);
device work goes through the C-ABI (hiprt_b200/api.py).
"""
import numpy as np

from . import api, scenes


def shard_pixels(width, height, rank, world, tile=32, sub=(8, 4)):
    """Row-major pixel indices rank `rank` traces, in trace order: tiles of `tile` x `tile` pixels row-major over the image,
    tile i owned by rank i mod world; inside a tile sub-blocks of sub[0] x sub[1] pixels row-major, row-major inside a block."""
    sw, sh = sub
    ys, xs = np.meshgrid(np.arange(height, dtype=np.int64), np.arange(width, dtype=np.int64), indexing="ij")
    tiles_x = (width + tile - 1) // tile
    tile_id = (ys // tile) * tiles_x + xs // tile
    mine = (tile_id % world) == rank
    ys, xs, tile_id = ys[mine], xs[mine], tile_id[mine]
    ly, lx = ys % tile, xs % tile
    blocks_x = (tile + sw - 1) // sw
    key = ((tile_id * (blocks_x * ((tile + sh - 1) // sh)) + (ly // sh) * blocks_x + lx // sw) * (sw * sh) + (ly % sh) * sw + lx % sw)
    order = np.argsort(key, kind="stable")
    return (ys * width + xs)[order]


def frame_rays(width, height, eye, look, pixels):
    """Primary rays (hiprtRay records) of the given row-major pixel indices."""
    return scenes.pinhole_rays(width, height, eye, look)[pixels]


def ao_rays_of(primary, hits, pixels, spp):
    """AO rays of the hit pixels, seeded by the GLOBAL pixel index (the image does not depend on the sharding), and for every
    local pixel the index of its first AO ray (undefined for misses)."""
    ao = scenes.ao_rays(primary, hits, spp=spp, pixel_ids=pixels)
    hit = hits["primID"] != api.INVALID
    first = (np.cumsum(hit) - 1).astype(np.int64) * spp
    first[~hit] = 0
    return ao, first.astype(np.uint32)


def replicate_geometry(ctx, geom, dist, src=0):
    """Broadcast a built geometry blob from rank `src` to every rank (NCCL over NVLink) and rebase its header there.
    `geom` is the api.Geometry on `src`, None elsewhere. Returns (handle, seconds) -- the handle is `geom.h` on src."""
    import time

    t = ctx.torch
    rank = dist.get_rank()
    size = t.zeros(1, dtype=t.int64, device=ctx.device)
    if rank == src:
        size[0] = ctx.geometry_size(geom)
    dist.broadcast(size, src)
    nbytes = int(size.item())
    handle = geom.h if rank == src else ctx.allocate_geometry_blob(nbytes)
    blob = ctx.blob_tensor(handle, nbytes)
    t.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    dist.broadcast(blob, src)
    if rank != src:
        ctx.relocate_geometry(handle)
    t.cuda.synchronize()
    return handle, time.perf_counter() - t0


def gather_shards(dist, rank, local, pieces, counts, piece_pixels, framebuffer, dst=0):
    """The collective variant of the exchange: every rank's (padded, equally sized) piece to rank `dst`, which scatters
    piece r's first counts[r] values to the pixels rank r owns. Works on any backend (NCCL on the box, gloo in the CPU tests)."""
    dist.gather(local, pieces if rank == dst else None, dst=dst)
    if rank == dst:
        for r, piece in enumerate(pieces):
            framebuffer[piece_pixels[r]] = piece[: counts[r]]


class AoFrame:
    """The sharded frame on one rank. dist=None: single process."""

    FLAG_STRIDE = 16  # uint32 words between the completion flags of two ranks (one 64-byte line each)

    def __init__(self, ctx, geom, width, height, eye, look, spp=16, rank=0, world=1, tile=32, dist=None, exchange="p2p",
                 primary_format=api.FORMAT_FULL, ao_format=api.FORMAT_PRIM_ID, light=None, overlap=True):
        """light=None: `spp` AO rays per hit pixel; light=(x, y, z): ONE shadow ray per hit pixel toward that point light
        (the resolve then yields visibility 0 / 1).
        overlap: the two trace launches of a frame are independent (the secondary rays of the frame are resident, as the primary
        ones are), so they are queued on two streams: the closest-hit launch ends in a long tail of a few warps with grazing
        rays (tools/probe_tail.py: one warp of horizon rays runs 0.10 ms on its own) under which the any-hit launch fills the
        machine. The resolve waits for both."""
        t = ctx.torch
        self.overlap = overlap
        self.side = t.cuda.Stream(device=ctx.device) if overlap else None
        self.fork, self.join = (t.cuda.Event(), t.cuda.Event()) if overlap else (None, None)
        if light is not None:
            spp = 1
        self.light = light
        self.ctx, self.geom, self.width, self.height, self.spp = ctx, geom, width, height, spp
        self.rank, self.world, self.dist, self.exchange, self.tile = rank, world, dist, exchange, tile
        self.primary_format, self.ao_format = primary_format, ao_format
        self.pixels = shard_pixels(width, height, rank, world, tile)
        self.primary = frame_rays(width, height, eye, look, self.pixels)
        self.d_primary = ctx.rays_to_device(self.primary)
        self.hits0 = ctx.hits_to_host(ctx.trace(geom, self.d_primary))
        if light is None:
            self.ao, first = ao_rays_of(self.primary, self.hits0, self.pixels, spp)
        else:
            self.ao, ok = scenes.shadow_rays(self.primary, self.hits0, light)
            first = np.where(ok, np.cumsum(ok) - 1, 0).astype(np.uint32)
        self.d_ao = ctx.rays_to_device(self.ao) if len(self.ao) else t.zeros((1, 32), dtype=t.uint8, device=ctx.device)
        self.n_primary, self.n_ao = len(self.primary), len(self.ao)
        self.d_first = t.from_numpy(first.view(np.int32)).to(ctx.device)
        self.d_pixels = t.from_numpy(self.pixels.astype(np.uint32).view(np.int32)).to(ctx.device)
        self.d_primary_rec = t.empty((self.n_primary, api.FORMAT_WORDS[primary_format]), dtype=t.int32, device=ctx.device)
        self.d_ao_rec = t.empty((max(self.n_ao, 1), api.FORMAT_WORDS[ao_format]), dtype=t.int32, device=ctx.device)
        self.block_counter = t.zeros(16, dtype=t.int32, device=ctx.device)
        self.timeout = t.zeros(2, dtype=t.int32, device=ctx.device)
        self.frame_no = 0
        self.graph = None
        self._setup_exchange()

    # ---- framebuffer placement -------------------------------------------------------------------
    def _setup_exchange(self):
        ctx, t = self.ctx, self.ctx.torch
        n_px = self.width * self.height
        self.fb_bytes = n_px * 4
        self.shared_ptr = self.peer_ptr = None
        if self.exchange == "p2p":
            total = self.fb_bytes + self.world * self.FLAG_STRIDE * 4
            handle = None
            if self.rank == 0:
                self.shared_ptr, handle = ctx.shared_alloc(total)
            if self.world > 1:
                box = [handle]
                self.dist.broadcast_object_list(box, src=0)
                ok = 1
                if self.rank != 0:
                    try:
                        self.peer_ptr = ctx.shared_open(box[0])
                    except api.HiprtError:
                        ok = 0
                flag = t.tensor([ok], dtype=t.int32, device=ctx.device)
                self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
                if int(flag.item()) == 0:  # the box does not allow CUDA IPC between these processes: gather with NCCL instead
                    if self.peer_ptr:
                        ctx.shared_close(self.peer_ptr)
                    if self.shared_ptr:
                        ctx.shared_free(self.shared_ptr)
                    self.shared_ptr = self.peer_ptr = None
                    self.exchange = "nccl (CUDA IPC unavailable)"
        if self.exchange == "p2p":
            base = self.shared_ptr if self.rank == 0 else self.peer_ptr
            self.fb_ptr = base
            self.flags_ptr = base + self.fb_bytes
            self.my_flag_ptr = self.flags_ptr + self.rank * self.FLAG_STRIDE * 4
            if self.rank == 0:
                holder = type("Shared", (), {"__cuda_array_interface__": {"shape": (n_px,), "typestr": "<f4", "data": (self.shared_ptr, False), "version": 2}})()
                self.framebuffer = t.as_tensor(holder, device=ctx.device)
        else:
            # local resolve target, then dist.gather of equally sized (padded) pieces to rank 0
            counts = [len(shard_pixels(self.width, self.height, r, self.world, tile=self.tile)) for r in range(self.world)] if self.world > 1 else [len(self.pixels)]
            self.counts, self.most = counts, max(counts)
            self.local_fb = t.zeros(self.most, dtype=t.float32, device=ctx.device)
            self.local_index = t.arange(len(self.pixels), dtype=t.int32, device=ctx.device)
            if self.rank == 0:
                self.framebuffer = t.zeros(n_px, dtype=t.float32, device=ctx.device)
                self.pieces = [t.empty(self.most, dtype=t.float32, device=ctx.device) for _ in range(self.world)]
                self.piece_pixels = [t.from_numpy(shard_pixels(self.width, self.height, r, self.world, tile=self.tile)).to(ctx.device) for r in range(self.world)]

    # ---- one frame ---------------------------------------------------------------------------------
    def capture(self):
        """Records one frame (two trace launches on two streams, resolve + exchange, the wait on rank 0) into a CUDA graph;
        step() then replays it with one launch. The completion flags count on the device (doneValue = 0 mode), so every replay
        releases / waits for the next frame number. Only the peer-store exchange can be recorded (NCCL calls are not)."""
        t = self.ctx.torch
        if self.exchange != "p2p":
            return False
        t.cuda.synchronize()
        graph = t.cuda.CUDAGraph()
        cap = t.cuda.Stream(device=self.ctx.device)
        cap.wait_stream(t.cuda.current_stream())
        with t.cuda.graph(graph, stream=cap):
            self._enqueue(None, None)
        t.cuda.synchronize()
        self.graph = graph
        return True

    def step(self, stream=None, marks=None):
        """Queues one frame on the current stream (asynchronous); on rank 0 the stream ends with the complete framebuffer.
        marks: optional (event, event) recorded after the primary and after the secondary launch (not with a captured graph)."""
        if self.graph is not None and marks is None:
            self.frame_no += 1
            self.graph.replay()
            return
        self._enqueue(stream, marks)

    def _enqueue(self, stream, marks):
        ctx = self.ctx
        self.frame_no += 1
        frame_value = 0  # counting mode: every launch releases / waits for the next frame number, kept on the device
        if self.overlap and self.n_ao:
            main = stream if stream is not None else ctx.torch.cuda.current_stream()
            self.fork.record(main)
            self.side.wait_event(self.fork)
            ctx.trace_format(self.geom, self.d_primary, self.primary_format, self.d_primary_rec, api.CLOSEST_HIT, main)
            if marks:
                marks[0].record(main)
            ctx.trace_format(self.geom, self.d_ao, self.ao_format, self.d_ao_rec, api.ANY_HIT, self.side)
            self.join.record(self.side)
            main.wait_event(self.join)
            if marks:
                marks[1].record(main)
        else:
            ctx.trace_format(self.geom, self.d_primary, self.primary_format, self.d_primary_rec, api.CLOSEST_HIT, stream)
            if marks:
                marks[0].record()
            if self.n_ao:
                ctx.trace_format(self.geom, self.d_ao, self.ao_format, self.d_ao_rec, api.ANY_HIT, stream)
            if marks:
                marks[1].record()
        if self.exchange == "p2p":
            ctx.resolve_ao_frame(self.d_primary_rec, self.primary_format, self.d_ao_rec, self.d_first, self.d_pixels, self.spp, self.fb_ptr,
                                 self.my_flag_ptr, frame_value, self.block_counter, stream)
            if self.rank == 0:
                ctx.wait_flags(self.flags_ptr, self.world, self.FLAG_STRIDE, frame_value, self.timeout, stream)
        else:
            ctx.resolve_ao_frame(self.d_primary_rec, self.primary_format, self.d_ao_rec, self.d_first, self.local_index, self.spp,
                                 self.local_fb.data_ptr(), 0, 0, self.block_counter, stream)
            if self.world > 1:
                gather_shards(self.dist, self.rank, self.local_fb, self.pieces if self.rank == 0 else None, self.counts,
                              self.piece_pixels if self.rank == 0 else None, self.framebuffer if self.rank == 0 else None)
            else:
                self.framebuffer[self.d_pixels.long()] = self.local_fb[: len(self.pixels)]

    launches_per_step = property(lambda s: 2 + 1 + (1 if (s.exchange == "p2p" and s.rank == 0) else 0))

    def image(self):
        """Rank 0: the framebuffer as float32 [height, width] on the host (ao in [0, 1], -1 = primary miss)."""
        self.ctx.torch.cuda.synchronize()
        assert int(self.timeout[0].item()) == 0, "a rank did not deliver its tiles (hiprtB200WaitFlags ran into its bound)"
        return self.framebuffer.cpu().numpy().reshape(self.height, self.width)

    def close(self):
        if self.peer_ptr:
            self.ctx.shared_close(self.peer_ptr)
            self.peer_ptr = None
        if self.shared_ptr:
            self.ctx.torch.cuda.synchronize()
            if self.world > 1:
                self.dist.barrier()  # peers unmap first
            self.ctx.shared_free(self.shared_ptr)
            self.shared_ptr = None
        elif self.world > 1 and self.exchange == "p2p":
            self.dist.barrier()


def reference_ao_image(width, height, pixels, hits, occluded, spp):
    """The image a frame must produce, from host-side hits / occlusion flags in `pixels` order (tests)."""
    img = np.full(width * height, -1.0, np.float32)
    hit = hits["primID"] != api.INVALID
    open_count = (~occluded.reshape(-1, spp)).sum(1).astype(np.float32)
    img[pixels[hit]] = open_count / np.float32(spp)
    return img.reshape(height, width)
