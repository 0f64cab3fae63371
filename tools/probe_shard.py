"""Why does half a frame cost more than half the time? Primary / AO launch time of rank 0's share of the C3 frame for several
shardings, cold (256 MiB flush) and warm L2, on one GPU."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)

def timed(fn, cold, reps=7):
    ts = []
    for _ in range(reps):
        if cold:
            flush.fill_(1)
        e0, e1 = ev(), ev()
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))

def probe(name, pixels):
    prim = frame.frame_rays(W, H, wl["eye"], wl["look"], pixels)
    d_prim = ctx.rays_to_device(prim)
    hits = ctx.hits_to_host(ctx.trace(geom, d_prim))
    ao, _ = frame.ao_rays_of(prim, hits, pixels, 16)
    d_ao = ctx.rays_to_device(ao)
    rec_p = torch.empty((len(prim), 12), dtype=torch.int32, device="cuda")
    rec_a = torch.empty((max(len(ao), 1), 1), dtype=torch.int32, device="cuda")
    out = {}
    for cold in (True, False):
        p = timed(lambda: ctx.trace_format(geom, d_prim, api.FORMAT_FULL, rec_p, api.CLOSEST_HIT), cold)
        a = timed(lambda: ctx.trace_format(geom, d_ao, api.FORMAT_PRIM_ID, rec_a, api.ANY_HIT), cold)
        out["cold" if cold else "warm"] = (round(p, 4), round(a, 4))
    print(f"{name:34s} primary {len(prim):8d} ao {len(ao):9d}  cold p/a {out['cold']}  warm p/a {out['warm']}", flush=True)

full = frame.shard_pixels(W, H, 0, 1, 32)
probe("full frame", full)
for world in (2, 8):
    for tile in (32, 128):
        probe(f"rank0 of {world}, tile {tile} interleaved", frame.shard_pixels(W, H, 0, world, tile))
    # contiguous: the first 1/world of the frame's tiles in trace order (rows of tiles from the top), and a middle chunk
    n = len(full) // world
    probe(f"first 1/{world} of the trace order", full[:n])
    probe(f"middle 1/{world} of the trace order", full[(world // 2) * n:(world // 2 + 1) * n])
    probe(f"last 1/{world} of the trace order", full[-n:])
