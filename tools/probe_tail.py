"""Critical path of the primary launch: time of single 32-ray batches (one warp) taken along the frame, and of 1/2/4-row strips."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, scenes, frame
import bench
wl = bench.WORKLOADS["c3_terrain_10M_ao"]
verts, idx = bench.make_mesh(wl["mesh"])
ctx = api.Context(0)
geom = ctx.build(torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda(), flags=api.BALANCED_BUILD)
W, H = bench.WIDTH, bench.HEIGHT
full = frame.shard_pixels(W, H, 0, 1, 32)
prim = frame.frame_rays(W, H, wl["eye"], wl["look"], full)
d_prim = ctx.rays_to_device(prim)
rec = torch.empty((len(prim), 12), dtype=torch.int32, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)
def t_of(lo, n, reps=5):
    sub = d_prim[lo:lo + n]
    out = rec[lo:lo + n]
    ts = []
    for _ in range(reps):
        e0, e1 = ev(), ev()
        e0.record(); ctx.trace_format(geom, sub, api.FORMAT_FULL, out, api.CLOSEST_HIT); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
n = len(prim)
print("empty-ish launch (32 sky rays):", round(t_of(0, 32), 4), "ms")
# per-warp latency along the frame: every 997th batch
lat = []
for b in range(0, n // 32, 499):
    lat.append((b, t_of(b * 32, 32)))
lat.sort(key=lambda x: -x[1])
print("slowest single warps (batch, ms):", [(b, round(t, 4)) for b, t in lat[:12]])
print("median single warp ms:", round(float(np.median([t for _, t in lat])), 4))
worst = lat[0][0]
# around the worst batch: 1 warp, 4 warps, 32 warps, 1024 warps
for k in (1, 4, 32, 256, 2048):
    lo = max(0, worst - k // 2) * 32
    print(f"{k} warps around the slowest: {t_of(lo, 32 * k):.4f} ms")
# per-ray: the 32 rays of the worst batch one by one
per = [t_of(worst * 32 + i, 1, reps=3) for i in range(32)]
print("single rays of the slowest warp (ms):", [round(x, 4) for x in per])
bp, lp = ctx.trace_stats(geom, d_prim[worst * 32:worst * 32 + 32], api.CLOSEST_HIT)
print("visits of that warp's 32 rays: box", bp, "leaf", lp)
