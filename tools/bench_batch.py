#!/usr/bin/env python
"""Config C5 (build half): hiprtBuildGeometries over many small geometries (2 .. 512 triangles, batchBuildMaxPrimCount = 512).
Prints geometries/s and Mtris/s of one batched call (CUDA events, mean of `reps`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api
n_geoms = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = api.Context(0)
g = np.arange(n_geoms, dtype=np.uint64)
sizes = (2 + ((g * 2654435761) % (2 ** 32) >> 7) % 511).astype(np.int64)
first = np.concatenate([[0], np.cumsum(sizes)[:-1]])
total = int(sizes.sum())
rng = np.random.default_rng(5)
pool = (rng.random((total, 1, 3), dtype=np.float32) + (rng.random((total, 3, 3), dtype=np.float32) - 0.5) * 0.2).reshape(-1, 3)
d_pool = torch.from_numpy(pool).cuda()
inputs = [ctx.mesh_input(d_pool[3 * int(first[k]): 3 * int(first[k] + sizes[k])]) for k in range(n_geoms)]
geoms = ctx.build_geometries(inputs, batch_max=512)
torch.cuda.synchronize()
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.rebuild_geometries(geoms, batch_max=512); e1.record(); e1.synchronize()
    ms.append(e0.elapsed_time(e1))
m = float(np.mean(ms))
print(f"batch build: {n_geoms} geometries, {total} triangles: {m:.3f} ms -> {n_geoms / m / 1e3:.2f} Mgeoms/s, {total / m / 1e3:.1f} Mtris/s "
      f"(includes the host->device copy of {n_geoms} x 128 B build inputs)")
