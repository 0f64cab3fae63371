"""CPU model of warp utilisation of one-thread-per-ray traversal, from the oracle's per-ray visit counts."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from hiprt_b200 import scenes
from oracle import binding as orc
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2_sphere_1M"]
verts, idx = bench.make_mesh(wl["mesh"])
t0 = time.time()
g = orc.Geometry(verts, idx)
print("oracle build", round(time.time() - t0, 1), "s")
W, H = 1920, 1080
rays = scenes.pinhole_rays(W, H, wl["eye"], wl["look"])[scenes.tile_order(W, H)]
sub = rays[: 32 * 8192 * 2]  # a band of the frame
sub = rays[len(rays) // 2 - 32 * 8192: len(rays) // 2 + 32 * 8192]
t0 = time.time()
bv, lv = g.trace_visits(sub)
print("trace", round(time.time() - t0, 1), "s;", "mean box visits", bv.mean(), "mean leaf visits", lv.mean())
cb, cl = 160.0, 190.0  # instruction cost per box / leaf visit
work = cb * bv + cl * lv
w = work.reshape(-1, 32)
print("tail efficiency (sum / 32*max) per warp:", (w.sum(1) / (32 * np.maximum(w.max(1), 1))).mean(), " weighted:", w.sum() / (32 * w.max(1)).sum())
b32, l32 = bv.reshape(-1, 32), lv.reshape(-1, 32)
# while-while lower bound: each warp iteration is max over lanes of boxes then leaves
print("box-phase eff", b32.sum() / (32 * b32.max(1)).sum(), "leaf-phase eff", l32.sum() / (32 * l32.max(1)).sum())
print("fraction of rays with zero leaf visits", (lv == 0).mean(), "zero-work warps", (w.max(1) <= cb).mean())
