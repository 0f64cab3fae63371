#!/usr/bin/env python
"""End-to-end (pinned host buffers -> hiprtB200TraceGeometryHost -> pinned host hits) throughput of the C2 frame only;
HIPRT_B200_HOST_CHUNK selects the pipeline chunk size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, scenes
import bench
wl = bench.WORKLOADS["c2_sphere_1M"]
verts, idx = bench.make_mesh(wl["mesh"])
ctx = api.Context(0)
geom = ctx.build(torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda())
primary = scenes.pinhole_rays(bench.WIDTH, bench.HEIGHT, wl["eye"], wl["look"])[scenes.tile_order(bench.WIDTH, bench.HEIGHT)]
hits0 = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(primary)))
shadow, _ = scenes.shadow_rays(primary, hits0, wl["light"])
pin = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory().numpy().view(a.dtype)
hp, hs = pin(primary), pin(shadow)
op, os_ = pin(np.zeros(len(primary), api.HIT_DTYPE)), pin(np.zeros(len(shadow), api.HIT_DTYPE))
ts = []
for it in range(13):
    t0 = time.perf_counter()
    ctx.trace_host(geom, hp, op, api.CLOSEST_HIT)
    ctx.trace_host(geom, hs, os_, api.ANY_HIT)
    if it >= 3:
        ts.append(time.perf_counter() - t0)
n = len(primary) + len(shadow)
print("chunk", os.environ.get("HIPRT_B200_HOST_CHUNK", "default"), "e2e Mrays/s", round(n / np.mean(ts) / 1e6, 1), "best", round(n / min(ts) / 1e6, 1),
      "GB/s down", round(48 * n / np.mean(ts) / 1e9, 1))
