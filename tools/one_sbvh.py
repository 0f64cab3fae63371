"""One high-quality build of the 10 M-triangle terrain (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiprt_b200 import api, scenes
from tests.test_gpu_parity import to_dev
ctx = api.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
verts, idx = scenes.terrain_mesh(n)
g = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx), flags=api.HIGH_QUALITY_BUILD)
ctx.torch.cuda.synchronize()
print("leaves", int(g.header()["primNodeCount"]))
