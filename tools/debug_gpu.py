"""Step-by-step GPU bring-up script (prints after every stage; run under `timeout`)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
t0 = time.time()
def log(*a):
    print(f"[{time.time()-t0:7.2f}s]", *a, flush=True)
import torch
log("torch imported", torch.cuda.is_available())
from hiprt_b200 import api, scenes
from oracle import binding as orc
ctx = api.Context(0)
log("context created")
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/cornell_box.npz"))
verts = d["vertices"]; idx = np.arange(96, dtype=np.int32).reshape(-1, 3)
def run(verts, idx, name):
    v = torch.from_numpy(verts).to(ctx.device)
    i = None if idx is None else torch.from_numpy(idx).to(ctx.device)
    inp = ctx.mesh_input(v, i)
    sz = ctx.geometry_temp_size(inp)
    log(name, "temp size", sz)
    temp = ctx.alloc_temp(sz); temp.zero_()
    g = ctx.create_geometry(inp)
    log(name, "created")
    ctx.build_geometry(inp, g, temp)
    torch.cuda.synchronize()
    log(name, "built")
    geom = api.Geometry(ctx, g, inp, temp, 0)
    h = geom.header()
    log(name, "header", {k: int(h[k]) for k in h.dtype.names})
    ctl = temp[:64].cpu().numpy().view(np.uint32)
    log(name, "control", ctl.tolist())
    return geom
geom = run(verts, idx, "cornell")
rays = scenes.cornell_rays(64, 64)
hits = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
torch.cuda.synchronize()
log("traced", int((hits["primID"] != api.INVALID).sum()))
ref = orc.Geometry(verts, idx.astype(np.uint32)).trace(rays, tie_mode=1)
log("prim match", float((hits["primID"] == ref["primID"]).mean()), "t match", float((hits["t"] == ref["t"]).mean()))
for n in (1, 2, 5, 1000, 100000, 2000000):
    vs = scenes.random_soup(n, seed=n)
    t1 = time.time()
    g2 = run(vs, None, f"soup{n}")
    log("soup", n, "done in", time.time() - t1)
v, i = scenes.sphere_mesh(300, 300)
g3 = run(v, i.astype(np.int32), "sphere300")
log("all done")
