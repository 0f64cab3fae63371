"""Hunt for intermittent build failures: fresh geometry + fresh temp each time, mixed inputs, validate everything."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests import helpers as H
from tests.test_gpu_parity import check_build_against_oracle, to_dev
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
ctx = api.Context(0)
tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
cases = {"identical3000": np.tile(tri, (3000, 1)), "identical5000": np.tile(tri, (5000, 1)), "soup5000": scenes.random_soup(5000, seed=4),
         "soup33": scenes.random_soup(33, seed=3), "soup1000": scenes.random_soup(1000, seed=1000)}
oracles = {k: orc.Geometry(v) for k, v in cases.items()}
bad = 0
t0 = time.time()
for r in range(reps):
    for name, verts in cases.items():
        for flags in (1, 0):
            geom = ctx.build(to_dev(ctx, verts), flags=flags)
            n = len(verts) // 3
            try:
                boxes, leaves = geom.nodes()
                H.validate_wide_bvh(boxes, leaves, n)
                assert geom.build_error_flag() == 0, 'watchdog'
                if flags == 0:
                    check_build_against_oracle(ctx, geom, oracles[name])
            except AssertionError as e:
                bad += 1
                ctl = geom.temp[:64].cpu().numpy().view(np.uint32)
                print(f"rep {r} {name} flags {flags} FAILED: {str(e)[:300]}\n  control {ctl.tolist()}\n  header {geom.header()}", flush=True)
                if bad <= 3:
                    np.savez(os.path.join(out, f"flaky_{bad}_{name}_{flags}.npz"), boxes=boxes, leaves=leaves,
                             scratch=geom.temp_array("scratch", api.SCRATCH_DTYPE, n - 1), keys=geom.temp_array("keys0", np.uint32, n),
                             vals=geom.temp_array("vals0", np.uint32, n), flags_arr=geom.temp_array("flags", np.uint32, n), verts=verts)
            geom.destroy()
print("hunt reps", reps, "failures", bad, "seconds", round(time.time() - t0, 1))
