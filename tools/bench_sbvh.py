"""High-quality build (SBVH): this repo against the unmodified reference on the same GPU -- build time, tree quality (SAH-style cost,
references), and the time the REFERENCE's own traversal kernel needs through each tree (same kernel, same rays: the trees alone differ).
Usage: python tools/bench_sbvh.py [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiprt_b200 import api, scenes
from oracle import ref_binding as ref
from tests import helpers as H
from tests.test_gpu_parity import to_dev
from tests.test_oracle_sbvh import long_thin_soup

ctx = api.Context(0)
t = ctx.torch
cases = {
    "needles 200k (long thin triangles, mixed sizes)": lambda: long_thin_soup(200000, seed=3),
    "sphere 1M": lambda: scenes.sphere_mesh(1000, 500),
    "terrain 10M": lambda: scenes.terrain_mesh(2236),
}
out = {}
for name, make in cases.items():
    verts, idx = make()
    d_v = to_dev(ctx, verts)
    d_i = None if idx is None else to_dev(ctx, idx)
    n_tris = len(verts) // 3 if idx is None else len(idx)
    rays = scenes.random_rays(1000000, verts.min(0), verts.max(0), seed=5)
    row = {"triangles": n_tris}
    ours = {}
    for tag, flags in (("fast", api.FAST_BUILD), ("balanced", api.BALANCED_BUILD), ("high_quality", api.HIGH_QUALITY_BUILD), ("high_quality_no_splits", api.HIGH_QUALITY_BUILD | 4)):
        g = ctx.build(d_v, d_i, flags=flags)
        best = 1e9
        for _ in range(5):
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            ctx.build_geometry(g.inp, g.h, g.temp, flags)
            e1.record()
            t.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        boxes, leaves = g.nodes()
        cost = H.validate_wide_bvh(boxes, leaves, n_tris, duplicated_references=True)
        hits, ms = ref.trace_geometry(g.h, rays, kind=2, reps=5)
        ours[tag] = {"build_ms": round(best, 3), "mtris_per_s": round(n_tris / best / 1e3, 1), "leaves": int(len(leaves)), "cost": round(cost, 2),
                     "reference_kernel_closest_ms": round(ms, 4), "temp_MB": round(g.temp.numel() / 2**20, 1)}
        if tag == "fast":
            base_hits = hits
        else:
            # (the reference's kernel keeps the hit it finds LAST among equally near ones: near a shared edge the primitive, and the
            # last bits of t with it, depend on the visiting order, i.e. on the tree)
            same = hits["primID"] == base_hits["primID"]
            dt = np.abs(hits["t"][~same] - base_hits["t"][~same]) / np.maximum(np.abs(base_hits["t"][~same]), 1e-20)
            assert (dt.size == 0 or dt.max() < 1e-5) and same.mean() > 0.999, (dt.max() if dt.size else 0, same.mean())
        g.destroy()
    row["this_repo"] = ours
    try:
        rg = ref.RefGeometry(verts, idx, build_flags=api.HIGH_QUALITY_BUILD, reps=2)
        boxes, leaves = rg.nodes()
        cost = H.validate_wide_bvh(boxes.view(api.BOX4_DTYPE), leaves.view(api.TRIPAIR_DTYPE), n_tris, duplicated_references=True)
        hits, ms = rg.trace(rays, kind=2, reps=5)
        same = hits["primID"] == base_hits["primID"]
        row["hiprt_reference"] = {"build_ms": round(rg.build_ms, 3), "mtris_per_s": round(n_tris / rg.build_ms / 1e3, 1), "leaves": int(len(leaves)), "cost": round(cost, 2),
                                  "reference_kernel_closest_ms": round(ms, 4), "primID_agreement": float(same.mean())}
        rg.destroy()
    except Exception as e:  # noqa: BLE001 -- the reference's builder may not run on this device
        row["hiprt_reference"] = {"unavailable": str(e)[:300]}
    print(name, json.dumps(row))
    out[name] = row
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
