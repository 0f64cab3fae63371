"""H-PLOC search radius against tree quality: SAH-style cost of the wide BVH and node visits per ray on several meshes
(HIPRT_B200_PLOC_RADIUS is read once per process: run once per radius)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, scenes
from tests import helpers as H
ctx = api.Context(0)
def soup_clustered(n, seed):
    rng = np.random.default_rng(seed)
    centers = rng.random((40, 3)) * 4
    c = centers[rng.integers(0, 40, n)] + rng.normal(0, 0.15, (n, 3))
    size = rng.lognormal(-3.5, 1.0, (n, 1, 1))
    v = c[:, None, :] + (rng.random((n, 3, 3)) - 0.5) * size
    return np.ascontiguousarray(v.reshape(-1, 3).astype(np.float32)), None
cases = {"terrain 1024": lambda: scenes.terrain_mesh(1024), "sphere 600x300": lambda: scenes.sphere_mesh(600, 300),
         "soup 300k uniform": lambda: (scenes.random_soup(300000, seed=3), None), "soup 300k clustered, mixed sizes": lambda: soup_clustered(300000, 5),
         "icosphere 6": lambda: scenes.icosphere(6, 1.0)}
print("radius", os.environ.get("HIPRT_B200_PLOC_RADIUS", "8 (default)"))
for name, make in cases.items():
    verts, idx = make()
    d_v = torch.from_numpy(verts).cuda()
    d_i = None if idx is None else torch.from_numpy(idx.view(np.int32)).cuda()
    out = []
    for flags, tag in ((api.FAST_BUILD, "lbvh"), (api.BALANCED_BUILD, "hploc")):
        g = ctx.build(d_v, d_i, flags=flags)
        boxes, leaves = g.nodes()
        cost = H.validate_wide_bvh(boxes, leaves)
        lo, hi = verts.min(0), verts.max(0)
        rays = scenes.random_rays(200000, lo, hi, seed=3)
        b, l = ctx.trace_stats(g, ctx.rays_to_device(rays))
        out.append(f"{tag}: cost {cost:9.2f} box visits/ray {b / len(rays):6.2f} leaf {l / len(rays):5.2f}")
        g.destroy()
    print(f"{name:34s}", " | ".join(out), flush=True)
