"""GPU: hiprtBuildTraceKernelsFromBitcode end to end -- the test kernels compiled offline by nvcc as relocatable device code,
linked with the generated dispatchers, loaded and launched -- must return exactly the hits of the same kernel built from source
through hiprtBuildTraceKernels (NVRTC). The link step itself is covered without a GPU by tests/test_bitcode_link.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from hiprt_b200 import api, scenes
from tests.test_gpu_parity import to_dev
from tests.test_gpu_user_kernels import Launcher, build_kernels, dev_ptr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def test_kernels_from_bitcode_trace_like_kernels_from_source(tmp_path):
    ctx = api.Context(0)
    t = ctx.torch
    path = str(tmp_path / "user_kernels.cubin")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-dc", "-cubin", "-I", os.path.join(ROOT, "include"),
                           "-include", "hiprt/hiprt_device.h", "-x", "cu", os.path.join(ROOT, "tests", "kernels", "user_kernels.cu"), "-o", path])
    image = open(path, "rb").read()
    sets = (api.FuncNameSet * 1)(api.FuncNameSet(b"intersectSphere", None))
    names = (C.c_char_p * 1)(b"customKernel")
    fns = (C.c_void_p * 1)()
    ctx._check(api.lib().hiprtBuildTraceKernelsFromBitcode(ctx.h, 1, names, b"user_kernels", image, len(image), 1, 1, sets, fns, False),
               "hiprtBuildTraceKernelsFromBitcode")
    (fn_source,) = build_kernels(ctx, ["customKernel"], [("intersectSphere", None)], num_geom_types=1)

    rng = np.random.default_rng(3)
    centers = rng.uniform(-1, 1, (200, 3)).astype(np.float32)
    radii = rng.uniform(0.03, 0.1, 200).astype(np.float32)
    d_spheres = to_dev(ctx, np.concatenate([centers, radii[:, None]], 1).astype(np.float32))
    inp = ctx.aabb_input(to_dev(ctx, np.concatenate([centers - radii[:, None], centers + radii[:, None]], 1).astype(np.float32)), geom_type=0)
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp))
    geom = ctx.create_geometry(inp)
    ctx.build_geometry(inp, geom, temp)
    table = C.c_void_p()
    ctx._check(api.lib().hiprtCreateFuncTable(ctx.h, 1, 1, C.byref(table)), "hiprtCreateFuncTable")
    ctx._check(api.lib().hiprtSetFuncTable(ctx.h, table, 0, 0, api.FuncDataSet(d_spheres.data_ptr(), None)), "hiprtSetFuncTable")
    rays = scenes.random_rays(20000, [-1, -1, -1], [1, 1, 1], seed=4)
    d_rays, n = ctx.rays_to_device(rays), len(rays)
    out = []
    for fn in (C.c_void_p(fns[0]), fn_source):
        d_hits = t.zeros((n, 12), dtype=t.int32, device=ctx.device)
        d_count = t.zeros(1, dtype=t.int32, device=ctx.device)
        Launcher().launch(fn, n, 64, [geom, table, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(n), dev_ptr(d_count)])
        t.cuda.synchronize()
        out.append(ctx.hits_to_host(d_hits).copy())
    hit = out[0]["primID"] != api.INVALID
    assert hit.mean() > 0.1
    assert np.array_equal(out[0]["primID"], out[1]["primID"])
    assert out[0]["t"][hit].tobytes() == out[1]["t"][hit].tobytes()
    api.lib().hiprtDestroyFuncTable(ctx.h, table)
    ctx.destroy_geometry(geom)
    ctx.close()
