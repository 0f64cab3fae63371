#!/usr/bin/env python
"""Config C5, render half: tile-sharded diffuse path tracing of the C3 terrain on N GPUs (one process per GPU, BVH replicated,
64x64 tiles round-robin over ranks, NCCL only for the framebuffer gather and the timing reductions).

    python tools/bench_pathtrace.py [--size 1024] [--spp 8] [--depth 5]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_pathtrace.py ...

Prints one JSON line on rank 0: whole-job Mrays/s (rays of all ranks / slowest rank's device time) and the gather time."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from hiprt_b200 import api, multi_gpu, pathtrace, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--depth", type=int, default=5)
ap.add_argument("--terrain", type=int, default=2236)
ap.add_argument("--torch-bounce", action="store_true", help="bounce math in torch ops instead of the fused NVRTC kernel")
ap.add_argument("--batch-paths", type=int, default=1 << 25, help="paths per wavefront: samples are traced in batches of batch_paths / pixels")
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
ctx = api.Context(local)
verts, idx = scenes.terrain_mesh(args.terrain)
geom = ctx.build(torch.from_numpy(verts).to(ctx.device), torch.from_numpy(idx.view(np.int32)).to(ctx.device), flags=api.BALANCED_BUILD)
W = H = args.size
pix = torch.from_numpy(multi_gpu.pixels_of_rank(W, H, rank, world, tile=64)).to(ctx.device)
eye, look = (0.5, 0.6, -0.6), (0.5, 0.0, 0.5)
import functools
render = pathtrace.render if args.torch_bounce else functools.partial(pathtrace.render_fused, batch_paths=args.batch_paths)
warm = render(ctx, geom, pix, W, H, eye, look, spp=args.spp, max_depth=args.depth)  # warm-up = the whole render once: NVRTC compile, allocator growth
multi_gpu.gather_framebuffer(warm, W, H, rank, world, tile=64, dist=dist)  # ... and the gather once (NCCL sets up its peer channels lazily)
del warm
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
    torch.cuda.synchronize()  # (the barrier's own kernel waits for the slowest rank: it must be over before the clock starts)
if os.environ.get("PT_TWICE"):  # (diagnostic: the first full render after the process group came up may carry one-time costs)
    render(ctx, geom, pix, W, H, eye, look, spp=args.spp, max_depth=args.depth)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
stats = {"time_trace": True} if os.environ.get("PT_TIME_TRACE") else {}
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
img = render(ctx, geom, pix, W, H, eye, look, spp=args.spp, max_depth=args.depth, stats=stats)
e1.record()
full = multi_gpu.gather_framebuffer(img, W, H, rank, world, tile=64, dist=dist)
e2.record()
e2.synchronize()
ms_render, ms_gather = e0.elapsed_time(e1), e1.elapsed_time(e2)
t = torch.tensor([ms_render, ms_gather, float(stats["rays"])], dtype=torch.float64, device=ctx.device)
per_rank = None
if dist is not None:
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    per_rank = [round(float(g[0]), 3) for g in gathered]
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
else:
    tmax = tsum = t
if rank == 0:
    print(json.dumps({"workload": "c5_pathtrace_terrain", "n_gpus": world, "image": [W, H], "spp": args.spp, "max_depth": args.depth, "samples_per_batch": stats.get("samples_per_batch"),
                      "rays": int(tsum[2].item()), "render_ms_max_over_ranks": round(float(tmax[0]), 3), "gather_ms": round(float(tmax[1]), 3),
                      "mrays_per_s": round(float(tsum[2]) / float(tmax[0]) / 1e3, 1), "mean_radiance": round(float(full.mean()), 5),
                      "render_ms_per_rank": per_rank, "trace_ms_rank0": round(stats.get("trace_ms", 0.0), 3), "host_enqueue_ms_rank0": round(stats.get("host_enqueue_ms", 0.0), 3), "per_depth_ms_rank0": [round(x, 3) for x in stats.get("per_depth_ms", [])],
                      "per_depth_rays_rank0": stats.get("per_depth_rays", []),
                      "note": "wavefront over batches of samples: per batch one camera launch, per bounce one indirect trace launch + one bounce launch; the host queues the whole render and never reads a counter back"}))
if dist is not None:
    dist.destroy_process_group()
