"""Driver for ncu: the two trace launches of the C3 frame exactly as bench.py issues them (closest-hit over the primary rays into
48-byte hits, any-hit over the 16-spp AO rays into 4-byte occlusion words), cold L2 before each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, frame
import bench
wl = bench.WORKLOADS["c3_terrain_10M_ao"]
verts, idx = bench.make_mesh(wl["mesh"])
ctx = api.Context(0)
geom = ctx.build(torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda(), flags=api.BALANCED_BUILD)
fr = frame.AoFrame(ctx, geom, bench.WIDTH, bench.HEIGHT, wl["eye"], wl["look"], spp=16, overlap=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    flush.fill_(1)
    ctx.trace_format(geom, fr.d_primary, fr.primary_format, fr.d_primary_rec, api.CLOSEST_HIT)
    flush.fill_(1)
    ctx.trace_format(geom, fr.d_ao, fr.ao_format, fr.d_ao_rec, api.ANY_HIT)
torch.cuda.synchronize()
print("profiled launches:", fr.n_primary, "primary rays (closest-hit),", fr.n_ao, "AO rays (any-hit)")
