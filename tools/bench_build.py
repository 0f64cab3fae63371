#!/usr/bin/env python
"""Build-only timing of one bench workload (per-phase, CUDA events through the library's build profile)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "c3_terrain_10M"
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
verts, idx = bench.make_mesh(bench.WORKLOADS[wl]["mesh"])
ctx = api.Context(0)
dv, di = torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda()
inp = ctx.mesh_input(dv, di)
temp = ctx.alloc_temp(ctx.geometry_temp_size(inp, flags))
g = ctx.create_geometry(inp, flags)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ctx.enable_build_profile(True)
tot, phases = [], {}
for it in range(reps + 3):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.build_geometry(inp, g, temp, flags); e1.record(); e1.synchronize()
    if it >= 3:
        tot.append(e0.elapsed_time(e1))
        for n, ms in ctx.build_profile():
            phases.setdefault(n, []).append(ms)
print(wl, "flags", flags, "build ms", round(float(np.mean(tot)), 4), "Mtris/s", round(len(idx) / np.mean(tot) / 1e3, 1),
      {k: round(float(np.mean(v)), 4) for k, v in phases.items()})
