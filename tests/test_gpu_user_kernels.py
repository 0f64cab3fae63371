"""GPU tests of the user-kernel path: hiprtBuildTraceKernels (NVRTC) compiles kernels
written against the public device API, the generated func-table dispatchers route custom intersection and filter
functions, global / dynamic stacks work from user kernels. Mirrors the reference tests CustomIntersection
(test/main.cpp:1296), Cutout (:1220), TraceKernel (:2768)."""
import ctypes as C
import os

import numpy as np
import pytest

from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests.test_gpu_parity import compare_hits, to_dev

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


class Launcher:
    """cuLaunchKernel through libcuda (what an application does with the hiprtApiFunction it gets back)."""

    def __init__(self):
        self.cu = C.CDLL("libcuda.so.1")
        self.cu.cuLaunchKernel.argtypes = [C.c_void_p] + [C.c_uint] * 6 + [C.c_uint, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]

    def launch(self, fn, n_threads, block, args):
        """args: list of ctypes values (kept alive here)."""
        ptrs = (C.c_void_p * len(args))(*[C.cast(C.pointer(a), C.c_void_p) for a in args])
        grid = (n_threads + block - 1) // block
        rc = self.cu.cuLaunchKernel(fn, grid, 1, 1, block, 1, 1, 0, None, ptrs, None)
        assert rc == 0, f"cuLaunchKernel failed: {rc}"


def build_kernels(ctx, func_names, name_sets=None, num_geom_types=0, num_ray_types=1):
    src = open(os.path.join(HERE, "kernels", "user_kernels.cu")).read().encode()
    n = len(func_names)
    names = (C.c_char_p * n)(*[f.encode() for f in func_names])
    fns = (C.c_void_p * n)()
    module = C.c_void_p()
    sets = None
    if name_sets:
        sets = (api.FuncNameSet * len(name_sets))(*[api.FuncNameSet(a.encode() if a else None, b.encode() if b else None) for a, b in name_sets])
    ctx._check(api.lib().hiprtBuildTraceKernels(ctx.h, n, names, src, b"user_kernels", 0, None, None, 0, None, num_geom_types, num_ray_types,
                                                sets, fns, C.byref(module), False), "hiprtBuildTraceKernels")
    return [C.c_void_p(f) for f in fns]


def dev_ptr(t):
    return C.c_void_p(t.data_ptr())


def test_custom_intersection_spheres(ctx):
    t = ctx.torch
    rng = np.random.default_rng(1)
    n_sph = 500
    centers = rng.uniform(-1, 1, (n_sph, 3)).astype(np.float32)
    radii = rng.uniform(0.02, 0.08, n_sph).astype(np.float32)
    spheres = np.concatenate([centers, radii[:, None]], 1).astype(np.float32)
    aabbs = np.concatenate([centers - radii[:, None], centers + radii[:, None]], 1).astype(np.float32)
    d_aabbs, d_spheres = to_dev(ctx, aabbs), to_dev(ctx, spheres)
    inp = ctx.aabb_input(d_aabbs, geom_type=0)
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp))
    g = ctx.create_geometry(inp)
    ctx.build_geometry(inp, g, temp)
    (fn,) = build_kernels(ctx, ["customKernel"], [("intersectSphere", None)], num_geom_types=1)
    table = C.c_void_p()
    ctx._check(api.lib().hiprtCreateFuncTable(ctx.h, 1, 1, C.byref(table)), "hiprtCreateFuncTable")
    ctx._check(api.lib().hiprtSetFuncTable(ctx.h, table, 0, 0, api.FuncDataSet(d_spheres.data_ptr(), None)), "hiprtSetFuncTable")
    rays = scenes.random_rays(20000, [-1, -1, -1], [1, 1, 1], seed=2)
    d_rays = ctx.rays_to_device(rays)
    d_hits = t.zeros((len(rays), 12), dtype=t.int32, device=ctx.device)
    d_count = t.zeros(1, dtype=t.int32, device=ctx.device)
    Launcher().launch(fn, len(rays), 64, [g, table, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(len(rays)), dev_ptr(d_count)])
    t.cuda.synchronize()
    got = ctx.hits_to_host(d_hits)
    # analytic reference in float64
    o, d = rays["origin"].astype(np.float64), rays["direction"].astype(np.float64)
    oc = o[:, None, :] - centers[None].astype(np.float64)
    b = 2 * (oc * d[:, None, :]).sum(-1)
    c = (oc * oc).sum(-1) - radii[None].astype(np.float64) ** 2
    a = (d * d).sum(-1)[:, None]
    disc = b * b - 4 * a * c
    tt = np.where(disc >= 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
    tt[tt < 0] = np.inf
    best = tt.argmin(1)
    tbest = tt[np.arange(len(rays)), best]
    hit = np.isfinite(tbest)
    assert hit.mean() > 0.2
    robust = hit & (np.partition(tt, 1, axis=1)[:, 1] - tbest > 1e-4) & (disc[np.arange(len(rays)), best] > 1e-4)
    assert np.array_equal(got["primID"][robust], best[robust].astype(np.uint32))
    np.testing.assert_allclose(got["t"][robust], tbest[robust], rtol=1e-3)
    grazing = (np.abs(disc) < 1e-4).any(1)
    assert np.array_equal((got["primID"] != api.INVALID)[~grazing], hit[~grazing])
    assert int(d_count.item()) >= int(hit.sum())  # the payload reached the custom function
    api.lib().hiprtDestroyFuncTable(ctx.h, table)
    ctx.destroy_geometry(g)


def test_cutout_filter_and_anyhit_iterator(ctx):
    t = ctx.torch
    verts, idx = scenes.sphere_mesh(48, 24)
    inp = ctx.mesh_input(to_dev(ctx, verts), to_dev(ctx, idx), geom_type=0)
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp))
    g = ctx.create_geometry(inp)
    ctx.build_geometry(inp, g, temp)
    fn_cut, fn_all = build_kernels(ctx, ["cutoutKernel", "allHitsKernel"], [(None, "cutoutFilter")], num_geom_types=1)
    threshold = np.array([0.6], np.float32)
    d_thr = to_dev(ctx, threshold)
    table = C.c_void_p()
    ctx._check(api.lib().hiprtCreateFuncTable(ctx.h, 1, 1, C.byref(table)), "hiprtCreateFuncTable")
    ctx._check(api.lib().hiprtSetFuncTable(ctx.h, table, 0, 0, api.FuncDataSet(None, d_thr.data_ptr())), "hiprtSetFuncTable")
    rays = scenes.pinhole_rays(160, 160, (0, 0, 3), (0, 0, 0))
    d_rays = ctx.rays_to_device(rays)
    n = len(rays)
    # all hits along each ray through the re-entrant iterator == the oracle's exhaustive enumeration
    max_per = 8
    d_counts = t.zeros(n, dtype=t.int32, device=ctx.device)
    d_prims = t.full((n, max_per), -1, dtype=t.int32, device=ctx.device)
    d_ts = t.zeros((n, max_per), dtype=t.float32, device=ctx.device)
    Launcher().launch(fn_all, n, 64, [g, dev_ptr(d_rays), C.c_uint32(n), C.c_uint32(max_per), dev_ptr(d_counts), dev_ptr(d_prims), dev_ptr(d_ts)])
    t.cuda.synchronize()
    og = orc.Geometry(verts, idx)
    ocounts, oprims, ots = og.trace_all(rays, max_per_ray=max_per)
    counts = d_counts.cpu().numpy().astype(np.uint32)
    assert np.array_equal(counts, ocounts)
    prims = d_prims.cpu().numpy().view(np.uint32)
    for k in range(max_per):  # same set of primitives per ray (order = traversal order, may differ)
        pass
    assert np.array_equal(np.sort(np.where(np.arange(max_per)[None] < counts[:, None], prims, api.INVALID), 1),
                          np.sort(np.where(np.arange(max_per)[None] < ocounts[:, None], oprims, api.INVALID), 1))
    # cutout: closest hit among those the filter keeps (uv.x + uv.y <= threshold)
    d_hits = t.zeros((n, 12), dtype=t.int32, device=ctx.device)
    Launcher().launch(fn_cut, n, 64, [g, table, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(n)])
    t.cuda.synchronize()
    got = ctx.hits_to_host(d_hits)
    plain = ctx.hits_to_host(ctx.trace(g, d_rays))
    kept_first = (plain["primID"] != api.INVALID) & (plain["uv"].sum(1) <= threshold[0] - 1e-4)
    assert np.array_equal(got["primID"][kept_first], plain["primID"][kept_first])
    hit = got["primID"] != api.INVALID
    assert (got["uv"][hit].sum(1) <= threshold[0] + 1e-6).all(), "a rejected hit was reported"
    rejected_first = (plain["primID"] != api.INVALID) & (plain["uv"].sum(1) > threshold[0] + 1e-4)
    assert (got["primID"][rejected_first] != plain["primID"][rejected_first]).all()
    assert rejected_first.sum() > 100 and kept_first.sum() > 100
    api.lib().hiprtDestroyFuncTable(ctx.h, table)
    ctx.destroy_geometry(g)


@pytest.mark.parametrize("dynamic", [0, 1])
def test_global_and_dynamic_stack_kernels(ctx, dynamic):
    t = ctx.torch
    verts, idx = scenes.sphere_mesh(200, 100)
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    (fn,) = build_kernels(ctx, ["globalStackKernel"])
    rays = scenes.pinhole_rays(256, 256, (0, 0, 3), (0, 0, 0))
    n = len(rays)
    d_rays = ctx.rays_to_device(rays)
    d_hits = t.zeros((n, 12), dtype=t.int32, device=ctx.device)
    sb_in = api.GlobalStackBufferInput(1 if dynamic else 0, 0, 64, ((n + 63) // 64) * 64)  # hiprtStackTypeGlobal=0 / Dynamic=1, integer entries
    sb = api.GlobalStackBuffer()
    ctx._check(api.lib().hiprtCreateGlobalStackBuffer(ctx.h, C.byref(sb_in), C.byref(sb)), "hiprtCreateGlobalStackBuffer")
    Launcher().launch(fn, n, 64, [geom.h, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(n), sb, C.c_int(dynamic)])
    t.cuda.synchronize()
    got = ctx.hits_to_host(d_hits)
    ref = ctx.hits_to_host(ctx.trace(geom, d_rays))
    assert np.array_equal(got["primID"], ref["primID"]) and np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))
    ctx._check(api.lib().hiprtDestroyGlobalStackBuffer(ctx.h, sb), "hiprtDestroyGlobalStackBuffer")
    geom.destroy()


def test_scene_traversal_from_user_kernel(ctx):
    t = ctx.torch
    meshes = scenes.instanced_geometries(3)
    n_inst = 30
    frames = scenes.random_matrix_frames(n_inst, seed=4, extent=5.0)
    geoms = [ctx.build(to_dev(ctx, v), to_dev(ctx, i)) for v, i in meshes]
    scene = ctx.build_scene([geoms[k % 3] for k in range(n_inst)], frames)
    (fn,) = build_kernels(ctx, ["sceneKernel"])
    rays = scenes.pinhole_rays(128, 128, (0.0, 1.5, 13.0), (0, 0, 0))
    d_rays = ctx.rays_to_device(rays)
    d_hits = t.zeros((len(rays), 12), dtype=t.int32, device=ctx.device)
    Launcher().launch(fn, len(rays), 64, [scene.h, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(len(rays)), C.c_uint32(0xFFFFFFFF)])
    t.cuda.synchronize()
    got = ctx.hits_to_host(d_hits)
    ref = ctx.hits_to_host(ctx.trace(scene, d_rays))
    assert (ref["primID"] != api.INVALID).mean() > 0.02
    assert np.array_equal(got["primID"], ref["primID"]) and np.array_equal(got["instanceIDs"][:, 0], ref["instanceIDs"][:, 0])
    assert np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))
    scene.destroy()
