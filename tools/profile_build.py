"""Build-only driver for ncu: builds one workload `reps` times through the C-ABI (no tracing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, scenes
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "c3_terrain_10M"
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
verts, idx = bench.make_mesh(bench.WORKLOADS[wl]["mesh"])
ctx = api.Context(0)
dv, di = torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda()
inp = ctx.mesh_input(dv, di)
temp = ctx.alloc_temp(ctx.geometry_temp_size(inp, flags))
g = ctx.create_geometry(inp, flags)
for _ in range(reps):
    ctx.build_geometry(inp, g, temp, flags)
    torch.cuda.synchronize()
print("built", wl, flags, reps)
