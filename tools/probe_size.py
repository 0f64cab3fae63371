"""Launch time against launch size: T(n) for the first n rays of the frame's primary / AO ray lists (every 8th tile = rank 0 of 8, and the
whole frame), warm L2, to separate the fixed part of a launch (ramp + tail) from the per-ray part."""
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
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, reps=6):
    ts = []
    for _ in range(reps):
        e0, e1 = ev(), ev()
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))
for world in (8, 1):
    px = frame.shard_pixels(W, H, 0, world, 32)
    prim = frame.frame_rays(W, H, wl["eye"], wl["look"], px)
    d_prim = ctx.rays_to_device(prim)
    hits = ctx.hits_to_host(ctx.trace(geom, d_prim))
    ao, _ = frame.ao_rays_of(prim, hits, px, 16)
    d_ao = ctx.rays_to_device(ao)
    rec = torch.empty((len(ao), 1), dtype=torch.int32, device="cuda")
    # shuffled AO order (pixel groups of 16 kept together) to remove any ordering effect
    perm = np.random.default_rng(1).permutation(len(ao) // 16)
    ao_sh = ao.reshape(-1, 16)[perm].reshape(-1)
    d_ao_sh = ctx.rays_to_device(ao_sh)
    print(f"world {world}: primary {len(prim)} ao {len(ao)}")
    for frac in (1 / 16, 1 / 8, 1 / 4, 1 / 2, 1.0):
        n = int(len(ao) * frac) // 32 * 32
        t_in = timed(lambda: ctx.trace_format(geom, d_ao[:n], api.FORMAT_PRIM_ID, rec[:n], api.ANY_HIT))
        t_sh = timed(lambda: ctx.trace_format(geom, d_ao_sh[:n], api.FORMAT_PRIM_ID, rec[:n], api.ANY_HIT))
        print(f"  ao first {n:9d} rays: {t_in:.4f} ms ({n / t_in / 1e3:8.1f} Mrays/s)   shuffled pixels: {t_sh:.4f} ms ({n / t_sh / 1e3:8.1f} Mrays/s)")
