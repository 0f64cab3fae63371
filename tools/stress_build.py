"""Stress: repeat builds of one input and validate each against the oracle (finds intermittent races)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests import helpers as H
from tests.test_gpu_parity import check_build_against_oracle, to_dev
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx = api.Context(0)
verts = scenes.random_soup(n, seed=4)
og = orc.Geometry(verts)
dv = to_dev(ctx, verts)
geom = ctx.build(dv, flags=flags)
bad = 0
garbage = os.environ.get("STRESS_GARBAGE", "1") == "1"
for r in range(reps):
    if garbage:  # a fresh temporary buffer holds arbitrary bytes: the build must not depend on them
        geom.temp.random_(0, 256)
        if r % 2:
            geom.temp.fill_(0 if r % 4 == 1 else 255)
    ctx.build_geometry(geom.inp, geom.h, geom.temp, flags)
    try:
        boxes, leaves = geom.nodes()
        H.validate_wide_bvh(boxes, leaves, n)
        if flags == 0:
            check_build_against_oracle(ctx, geom, og)
    except AssertionError as e:
        bad += 1
        ctl = geom.temp[:64].cpu().numpy().view(np.uint32)
        hdr = geom.header()
        print("rep", r, "FAILED:", str(e)[:200], "control", ctl.tolist(), "header", hdr, flush=True)
        if bad == 1:
            np.savez(f"gpurun_out/stress_fail_{n}.npz", boxes=boxes, leaves=leaves, scratch=geom.temp_array("scratch", api.SCRATCH_DTYPE, n - 1),
                     keys=geom.temp_array("keys0", np.uint32, n), vals=geom.temp_array("vals0", np.uint32, n), verts=verts)
print("stress", n, "reps", reps, "flags", flags, "failures", bad)
