#!/usr/bin/env python
"""Config C5 (build half): hiprtBuildGeometries over many small geometries (2 .. 512 triangles, batchBuildMaxPrimCount = 512).
Prints geometries/s and Mtris/s of one batched call (CUDA events, mean of `reps`).

    python tools/bench_batch.py [geometries] [reps]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_batch.py ...

Under torchrun the batch is split across the ranks (multi_gpu.split_batch: independent geometries, no exchange); the time is the
slowest rank's, the rates are whole-job."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, multi_gpu
n_geoms = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
ctx = api.Context(local)
g = np.arange(n_geoms, dtype=np.uint64)
sizes = (2 + ((g * 2654435761) % (2 ** 32) >> 7) % 511).astype(np.int64)
begin, end = multi_gpu.split_batch(n_geoms, rank, world)
sizes_mine = sizes[begin:end]
first = np.concatenate([[0], np.cumsum(sizes_mine)[:-1]])
total_mine, total = int(sizes_mine.sum()), int(sizes.sum())
rng = np.random.default_rng(5 + rank)
pool = (rng.random((total_mine, 1, 3), dtype=np.float32) + (rng.random((total_mine, 3, 3), dtype=np.float32) - 0.5) * 0.2).reshape(-1, 3)
d_pool = torch.from_numpy(pool).to(ctx.device)
inputs = [ctx.mesh_input(d_pool[3 * int(first[k]): 3 * int(first[k] + sizes_mine[k])]) for k in range(end - begin)]
geoms = ctx.build_geometries(inputs, batch_max=512)
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
    torch.cuda.synchronize()
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.rebuild_geometries(geoms, batch_max=512); e1.record(); e1.synchronize()
    ms.append(e0.elapsed_time(e1))
t = torch.tensor([float(np.mean(ms))], dtype=torch.float64, device=ctx.device)
if dist is not None:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
m = float(t.item())
if rank == 0:
    print(json.dumps({"workload": "c5_batch_build", "n_gpus": world, "geometries": n_geoms, "triangles": total, "ms_max_over_ranks": round(m, 3),
                      "mgeoms_per_s": round(n_geoms / m / 1e3, 2), "mtris_per_s": round(total / m / 1e3, 1),
                      "note": "one hiprtBuildGeometries call per rank over its slice; includes the host->device copy of the 128-byte build inputs"}))
if dist is not None:
    dist.destroy_process_group()
