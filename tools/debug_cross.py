"""Debug: dump the reference-built BVH for a case and walk it with our kernel (run under timeout)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiprt_b200 import api, scenes
from oracle import ref_binding as ref, binding as orc
out = sys.argv[1]
os.makedirs(out, exist_ok=True)
verts = np.load("tests/golden/cornell_box.npz")["vertices"]
rays = scenes.cornell_rays(64, 64)
rg = ref.RefGeometry(verts, None, build_flags=0)
hdr = rg.header()
print("header", hdr, hex(int(hdr["boxNodes"])), hex(int(hdr["primNodes"])), "handle", hex(rg.device_handle), flush=True)
ctx0 = api.Context(0)
raw = ctx0.read_device(rg.device_handle, 256)
print("raw header words", raw.view(np.uint32)[:16], flush=True)
boxes, leaves = rg.nodes()
np.savez(os.path.join(out, "ref_cornell_bvh.npz"), boxes=boxes, leaves=leaves, header=np.array([hdr]))
print("header", hdr, flush=True)
theirs, _ = rg.trace(rays, kind=2)
print("their hits", (theirs["primID"] != api.INVALID).mean(), flush=True)
o = orc.trace_nodes(boxes.view(orc.BOX4_DTYPE), leaves.view(orc.TRIPAIR_DTYPE), rays)
print("oracle on their nodes agrees:", (o["primID"] == theirs["primID"]).mean(), flush=True)
ctx = api.Context(0)
print("handle", hex(rg.device_handle), flush=True)
h = ctx.read_device(rg.device_handle, 48).view(api.GEOM_HEADER_DTYPE)[0]
print("header via our lib", h, flush=True)
ours = ctx.hits_to_host(ctx.trace(rg.device_handle, ctx.rays_to_device(rays[:32])))
print("ours on theirs (32 rays)", (ours["primID"] == theirs["primID"][:32]).mean(), flush=True)
ours = ctx.hits_to_host(ctx.trace(rg.device_handle, ctx.rays_to_device(rays)))
print("ours on theirs", (ours["primID"] == theirs["primID"]).mean(), flush=True)
