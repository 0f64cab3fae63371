"""Where the time of a high-quality build goes: %globaltimer stamps after every grid barrier (HIPRT_B200_SBVH_TIMELINE)."""
import os, sys
os.environ["HIPRT_B200_SBVH_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiprt_b200 import api, scenes
from tests.test_gpu_parity import to_dev
from tests.test_oracle_sbvh import long_thin_soup
ctx = api.Context(0)
cases = {"needles200k": lambda: long_thin_soup(200000, seed=3), "sphere1M": lambda: scenes.sphere_mesh(1000, 500), "terrain10M": lambda: scenes.terrain_mesh(2236)}
names = ["reset", "objbin", "objsplit", "spbin", "demand", "split", "distribute"]
for name in sys.argv[1:] or list(cases):
    verts, idx = cases[name]()
    g = ctx.build(to_dev(ctx, verts), None if idx is None else to_dev(ctx, idx), flags=api.HIGH_QUALITY_BUILD)
    ctx.build_geometry(g.inp, g.h, g.temp, api.HIGH_QUALITY_BUILD)
    ctx.torch.cuda.synchronize()
    st = g.temp[64:128].cpu().numpy().view(np.uint32)
    levels = int(st[5])
    tl = g.temp[128:128 + 128 * 8 * 8].cpu().numpy().view(np.uint64).reshape(128, 8)[:levels].astype(np.int64)
    d = np.diff(tl, axis=1) / 1e3  # us per phase
    print(name, "levels", levels, "total ms", (tl[-1, 7] - tl[0, 0]) / 1e6, "refs", int(st[2]))
    print("  phase totals (ms):", {n: round(float(d[:, i].sum()) / 1e3, 3) for i, n in enumerate(names)})
    for lv in range(levels):
        print("  level", lv, " ".join(f"{n}={d[lv, i]:.0f}" for i, n in enumerate(names)))
    g.destroy()
