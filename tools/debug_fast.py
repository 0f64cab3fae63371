"""Debug aid: hits of the batch-test geometries under the current HIPRT_B200_TRACE_VARIANT against the oracle."""
import sys
import numpy as np
sys.path.insert(0, ".")
from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests.test_gpu_batch import small_mesh
from tests.test_gpu_parity import to_dev

ctx = api.Context(0)
sizes = [1, 2, 3, 4, 5, 7, 8, 16, 17, 31, 32, 33, 64, 100]
meshes = [small_mesh(k, s) for s, k in enumerate(sizes)]
inputs = [ctx.mesh_input(to_dev(ctx, v), None if i is None else to_dev(ctx, i)) for v, i in meshes]
geoms = ctx.build_geometries(inputs, batch_max=512)
for (v, i), g, k in zip(meshes, geoms, sizes):
    og = orc.Geometry(v, i, allow_pairing=False)
    lo, hi = v.min(0), v.max(0)
    rays = scenes.random_rays(2000, lo - 0.05, hi + 0.05, seed=k)
    got = ctx.hits_to_host(ctx.trace(g, ctx.rays_to_device(rays)))
    ref = og.trace(rays, tie_mode=1)
    bad = np.nonzero(got["primID"] != ref["primID"])[0]
    print(f"k={k}: {len(bad)} mismatches; ref hits {(ref['primID'] != orc.INVALID).sum()} got hits {(got['primID'] != orc.INVALID).sum()}")
    for j in bad[:4]:
        print("   ray", j, "ref", ref["primID"][j], ref["t"][j], "got", got["primID"][j], got["t"][j], "ray", rays[j])
    if len(bad):
        boxes, leaves = g.nodes()
        print("   box nodes:", len(boxes), "children of root", boxes[0]["child"], "count", boxes[0]["childCount"])
