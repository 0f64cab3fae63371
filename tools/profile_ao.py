"""AO-ray driver for ncu: C3 terrain, balanced build, primary rays + `spp` AO rays per hit pixel traced any-hit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, scenes
import bench
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
flags = int(sys.argv[2]) if len(sys.argv) > 2 else api.BALANCED_BUILD
wl = bench.WORKLOADS["c3_terrain_10M_ao"]
verts, idx = bench.make_mesh(wl["mesh"])
ctx = api.Context(0)
geom = ctx.build(torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda(), flags=flags)
primary = scenes.pinhole_rays(bench.WIDTH, bench.HEIGHT, wl["eye"], wl["look"])[scenes.tile_order(bench.WIDTH, bench.HEIGHT)]
hits = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(primary)))
ao = scenes.ao_rays(primary, hits, spp=spp)
d_ao = ctx.rays_to_device(ao)
out = torch.empty((len(ao), 12), dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for kind, name in ((api.ANY_HIT, "any-hit"), (api.CLOSEST_HIT, "closest")):
    for _ in range(3):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.trace(geom, d_ao, out, kind); e1.record(); e1.synchronize()
    print("ao rays", name, len(ao), "ms", round(e0.elapsed_time(e1), 4), "Mrays/s", round(len(ao) / e0.elapsed_time(e1) / 1e3, 1))
