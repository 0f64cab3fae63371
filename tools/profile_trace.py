"""Trace-only driver for ncu: builds one workload, traces primary (closest) + shadow (any-hit) rays `reps` times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, scenes
import bench
wl_name = sys.argv[1] if len(sys.argv) > 1 else "c2_sphere_1M"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wl = bench.WORKLOADS[wl_name]
verts, idx = bench.make_mesh(wl["mesh"])
ctx = api.Context(0)
geom = ctx.build(torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda(), flags=flags)
primary = scenes.pinhole_rays(bench.WIDTH, bench.HEIGHT, wl["eye"], wl["look"])[scenes.tile_order(bench.WIDTH, bench.HEIGHT)]
dp = ctx.rays_to_device(primary)
hits = ctx.hits_to_host(ctx.trace(geom, dp))
shadow, _ = scenes.shadow_rays(primary, hits, wl["light"])
ds = ctx.rays_to_device(shadow)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    flush.fill_(1)
    ctx.trace(geom, dp)
    ctx.trace(geom, ds, kind=api.ANY_HIT)
torch.cuda.synchronize()
print("traced", wl_name, len(primary), len(shadow))
