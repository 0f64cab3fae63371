#!/usr/bin/env python
"""Config C4 (two-level scene: 100 geometries x 10 000 instances, hiprtBuildScene + hiprtSceneTraversalClosest):
scene-build time and Mrays/s of this repo next to the unmodified HIPRT kernels on the same GPU (oracle/_ref), same
protocol for both arms (warm L2, best of 5 launches after one warm-up). Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hiprt_b200 import api, scenes  # noqa: E402

W, H = 1920, 1080
n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
with_ref = "--no-ref" not in sys.argv

ctx = api.Context(0)
meshes = scenes.instanced_geometries(100)
geom_of = np.arange(n_inst) % len(meshes)
frames = scenes.random_matrix_frames(n_inst, seed=42, extent=50.0)
geoms = [ctx.build(torch.from_numpy(v).cuda(), torch.from_numpy(i.view(np.int32)).cuda()) for v, i in meshes]
scene = ctx.build_scene([geoms[g] for g in geom_of], frames)
rays = scenes.pinhole_rays(W, H, (0.0, 15.0, 130.0), (0.0, 0.0, 0.0))[scenes.tile_order(W, H)]
d_rays = ctx.rays_to_device(rays)
d_hits = torch.empty((len(rays), 12), dtype=torch.int32, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)


def best_of(fn, reps=6):
    best = 1e30
    for r in range(reps):
        e0, e1 = ev(), ev()
        e0.record(); fn(); e1.record(); e1.synchronize()
        if r > 0:
            best = min(best, e0.elapsed_time(e1))
    return best


ms_closest = best_of(lambda: ctx.trace(scene, d_rays, d_hits, api.CLOSEST_HIT))
ms_any = best_of(lambda: ctx.trace(scene, d_rays, d_hits, api.ANY_HIT))
ms_build = best_of(lambda: ctx.rebuild_scene(scene))
hits = ctx.hits_to_host(ctx.trace(scene, d_rays))
line = {"workload": "c4_scene", "geometries": 100, "instances": n_inst, "rays": len(rays), "hit_fraction": round(float((hits["primID"] != api.INVALID).mean()), 4),
        "this_repo": {"closest_mrays_per_s": round(len(rays) / ms_closest / 1e3, 1), "anyhit_mrays_per_s": round(len(rays) / ms_any / 1e3, 1),
                      "scene_build_ms": round(ms_build, 4), "scene_build_minstances_per_s": round(n_inst / ms_build / 1e3, 2)},
        "protocol": "warm L2, best of 5 launches"}
if with_ref:
    try:
        from oracle import ref_binding as ref

        rs = ref.RefScene([geoms[g].h for g in geom_of], frames, reps=6)
        rhits, rms = ref.trace_scene(rs.device_handle, rays, kind=2, reps=6)
        _, rms_any = ref.trace_scene(rs.device_handle, rays, kind=1, reps=6)
        same = rhits["primID"] == hits["primID"]
        line["hiprt_reference_gpu"] = {"closest_mrays_per_s": round(len(rays) / rms / 1e3, 1), "anyhit_mrays_per_s": round(len(rays) / rms_any / 1e3, 1),
                                       "scene_build_ms": round(rs.build_ms, 4), "primID_agreement": round(float(same.mean()), 6)}
    except Exception as e:  # optional infrastructure
        line["hiprt_reference_gpu"] = {"unavailable": str(e)[:300]}
print(json.dumps(line))
