"""GPU parity of the two paths that so far were held to pictures only -- custom primitives through the function table and scenes
instanced in scenes -- against the oracle's general traversal (oracle/hiprt_oracle.cpp walkScene / walkGeometry, itself pinned
by the reference's CustomIntersection / SceneIntersection / SceneIntersectionMlas images in tests/test_oracle_custom.py):
primID and every level of instanceIDs exact, t within 1e-5 relative."""
import ctypes as C

import numpy as np
import pytest

from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests.test_gpu_parity import to_dev
from tests.test_gpu_user_kernels import Launcher, build_kernels, dev_ptr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def launch_scene_kernel(ctx, fn, scene, rays):
    t = ctx.torch
    d_rays = ctx.rays_to_device(rays)
    d_hits = t.zeros((len(rays), 12), dtype=t.int32, device=ctx.device)
    Launcher().launch(fn, len(rays), 64, [scene.h, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(len(rays)), C.c_uint32(0xFFFFFFFF)])
    t.cuda.synchronize()
    return ctx.hits_to_host(d_hits).copy()


def test_three_level_scene_against_the_oracle(ctx):
    """Geometries -> 6 instances in a scene -> that scene 4 times in a middle scene -> the middle scene 3 times at the top."""
    meshes = scenes.instanced_geometries(3)
    leaf_frames = scenes.random_matrix_frames(6, seed=5, extent=2.0, scale_range=(0.6, 1.2))
    mid_frames = scenes.random_matrix_frames(4, seed=6, extent=6.0, scale_range=(0.8, 1.1))
    top_frames = scenes.random_matrix_frames(3, seed=7, extent=14.0, scale_range=(0.9, 1.0))

    geoms = [ctx.build(to_dev(ctx, v), to_dev(ctx, i)) for v, i in meshes]
    leaf = ctx.build_scene([geoms[k % 3] for k in range(6)], leaf_frames)
    mid = ctx.build_scene([leaf] * 4, mid_frames, instance_types=[api.INSTANCE_SCENE] * 4)
    top = ctx.build_scene([mid] * 3, top_frames, instance_types=[api.INSTANCE_SCENE] * 3)

    ogeoms = [orc.Geometry(v, i) for v, i in meshes]
    oleaf = orc.Scene(ogeoms, [k % 3 for k in range(6)], leaf_frames, 1)
    omid = orc.Scene([oleaf], [0] * 4, mid_frames, 1)
    otop = orc.Scene([omid], [0] * 3, top_frames, 1)

    (fn,) = build_kernels(ctx, ["sceneKernel"])
    rays = scenes.pinhole_rays(192, 192, (0.0, 4.0, 45.0), (0, 0, 0))
    got = launch_scene_kernel(ctx, fn, top, rays)
    ref = orc.trace_general(otop, rays, tie_mode=1)
    hit = ref["primID"] != orc.INVALID
    assert hit.mean() > 0.02
    assert np.array_equal(got["primID"] != api.INVALID, hit)
    same = (got["primID"] == ref["primID"]) & (got["instanceIDs"][:, :3] == ref["instanceIDs"][:, :3]).all(1)
    # a different primitive or instance is tolerated only at (nearly) the same distance
    dt = np.abs(got["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-20)
    assert (dt[hit] <= 1e-4).all(), dt[hit].max()
    assert (~same & hit).sum() <= 0.002 * hit.sum(), (~same & hit).sum()
    assert (got["instanceIDs"][hit, 3] == api.INVALID).all() and (got["instanceIDs"][hit, :3] != api.INVALID).all()
    # the library's own rays-in / hits-out entry point on the same scene: its fast kernel walks one instance level and hands every
    # ray that meets an instance of a scene to the multi-level pass that follows it (traceSceneDeferredKernel) -- same hits as the
    # user kernel above, bit for bit, for closest hit; same occlusion for any hit
    lib = ctx.hits_to_host(ctx.trace(top, ctx.rays_to_device(rays)))
    assert np.array_equal(lib["primID"], got["primID"]) and np.array_equal(lib["instanceIDs"][hit], got["instanceIDs"][hit])
    for f in ("t", "uv", "normal"):
        assert lib[f][hit].tobytes() == got[f][hit].tobytes(), f
    occluded = ctx.hits_to_host(ctx.trace(top, ctx.rays_to_device(rays), kind=api.ANY_HIT))["primID"] != api.INVALID
    assert np.array_equal(occluded, hit)
    # a mixed top level: one instance of a scene next to instances of geometries
    mixed_frames = scenes.random_matrix_frames(3, seed=9, extent=5.0, scale_range=(0.9, 1.1))
    mixed = ctx.build_scene([leaf, geoms[0], geoms[1]], mixed_frames, instance_types=[api.INSTANCE_SCENE, api.INSTANCE_GEOMETRY, api.INSTANCE_GEOMETRY])
    rays2 = scenes.pinhole_rays(160, 160, (0.0, 2.0, 20.0), (0, 0, 0))
    user = launch_scene_kernel(ctx, fn, mixed, rays2)
    lib2 = ctx.hits_to_host(ctx.trace(mixed, ctx.rays_to_device(rays2)))
    h2 = user["primID"] != api.INVALID
    assert h2.mean() > 0.02 and np.array_equal(lib2["primID"], user["primID"]) and np.array_equal(lib2["instanceIDs"][h2], user["instanceIDs"][h2])
    assert lib2["t"][h2].tobytes() == user["t"][h2].tobytes()
    mixed.destroy()
    for s in (top, mid, leaf):
        s.destroy()
    for g in geoms:
        g.destroy()


def test_custom_spheres_in_a_scene_against_the_oracle(ctx):
    """Spheres as custom primitives (geomType 0) and a triangle mesh instanced side by side: closest hits through the function
    table from a user kernel against the oracle's walker with the same sphere test as a callback."""
    rng = np.random.default_rng(11)
    centers = rng.uniform(-1, 1, (60, 3)).astype(np.float32)
    radii = rng.uniform(0.05, 0.15, 60).astype(np.float32)
    boxes = np.concatenate([centers - radii[:, None], centers + radii[:, None]], 1).astype(np.float32)
    spheres = np.concatenate([centers, radii[:, None]], 1).astype(np.float32)
    verts, idx = scenes.icosphere(2, 0.7)
    frames = scenes.random_matrix_frames(4, seed=12, extent=2.5, scale_range=(0.7, 1.3))

    inp = ctx.aabb_input(to_dev(ctx, boxes), geom_type=0)
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp))
    g_spheres = ctx.create_geometry(inp)
    ctx.build_geometry(inp, g_spheres, temp)
    g_mesh = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    scene = ctx.build_scene([g_spheres, g_mesh, g_spheres, g_mesh], frames)
    (fn,) = build_kernels(ctx, ["sceneTableKernel"], [("intersectSphere", None)], num_geom_types=1)
    d_spheres = to_dev(ctx, spheres)
    table = C.c_void_p()
    ctx._check(api.lib().hiprtCreateFuncTable(ctx.h, 1, 1, C.byref(table)), "hiprtCreateFuncTable")
    ctx._check(api.lib().hiprtSetFuncTable(ctx.h, table, 0, 0, api.FuncDataSet(d_spheres.data_ptr(), None)), "hiprtSetFuncTable")

    def intersect_sphere(ray, geom_type, prim):  # tests/kernels/user_kernels.cu intersectSphere, in float32
        f = np.float32
        s = spheres[prim]
        oc = ray["origin"] - s[:3]
        d = ray["direction"]
        a = f(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        b = f(2.0) * f(oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2])
        c = f(oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2]) - s[3] * s[3]
        disc = b * b - f(4.0) * a * c
        if disc < 0:
            return None
        t = (-b - np.sqrt(disc, dtype=f)) / (f(2.0) * a)
        return None if (t < ray["minT"] or t > ray["maxT"]) else t

    rays = scenes.pinhole_rays(96, 96, (0.0, 0.5, 9.0), (0, 0, 0))
    t = ctx.torch
    d_rays = ctx.rays_to_device(rays)
    d_hits = t.zeros((len(rays), 12), dtype=t.int32, device=ctx.device)
    Launcher().launch(fn, len(rays), 64, [scene.h, table, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(len(rays))])
    t.cuda.synchronize()
    got = ctx.hits_to_host(d_hits)
    oscene = orc.Scene([orc.AabbGeometry(boxes), orc.Geometry(verts, idx)], [0, 1, 0, 1], frames, 1)
    ref = orc.trace_general(oscene, rays, intersect_sphere, tie_mode=1)
    hit = ref["primID"] != orc.INVALID
    assert hit.mean() > 0.05
    dt = np.abs(got["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-20)
    # fused multiply-adds in the device's sphere test move t by a few ulp and flip the odd grazing ray
    agree = (got["primID"] == ref["primID"]) & (got["instanceIDs"][:, 0] == ref["instanceIDs"][:, 0])
    assert agree.mean() > 0.995, agree.mean()
    both = agree & hit
    on_mesh = np.isin(got["instanceIDs"][:, 0], (1, 3))
    assert (dt[both & on_mesh] <= 5e-5).all(), dt[both & on_mesh].max()  # triangles: the scene bar (fma-built matrices, tests/test_gpu_scene.py)
    # spheres: t = (-b - sqrt(b^2 - 4ac)) / 2a in float32 cancels near the silhouette, so the bar is the forward error of that
    # formula, evaluated in float64 on the object-space ray (the device contracts it to fmas, the callback above does not)
    sel = np.nonzero(both & ~on_mesh)[0]
    assert len(sel) > 100
    inst = got["instanceIDs"][sel, 0].astype(np.int64)
    m = np.zeros((4, 4, 4))
    m[:, :3, :] = np.asarray(frames, np.float64).reshape(4, 4, 4)[:, :3, :]
    m[:, 3, 3] = 1.0
    inv = np.linalg.inv(m)[inst]
    o = np.einsum("nij,nj->ni", inv[:, :3, :3], rays["origin"][sel].astype(np.float64)) + inv[:, :3, 3]
    d = np.einsum("nij,nj->ni", inv[:, :3, :3], rays["direction"][sel].astype(np.float64))
    sp = spheres[got["primID"][sel]].astype(np.float64)
    oc = o - sp[:, :3]
    a, b, c = (d * d).sum(1), 2 * (oc * d).sum(1), (oc * oc).sum(1) - sp[:, 3] ** 2
    disc = np.maximum(b * b - 4 * a * c, 1e-300)
    t64 = (-b - np.sqrt(disc)) / (2 * a)
    eps = np.finfo(np.float32).eps
    # inputs (matrix entries, object-space ray) carry a few eps each; the discriminant amplifies them by (b^2 + 4a|c|) / disc
    bound = 1e-5 + 64 * eps * (b * b + 4 * a * np.abs(c)) / (2 * a * np.sqrt(disc) * np.abs(t64))
    for name, tt in (("device", got["t"][sel]), ("oracle", ref["t"][sel])):
        err = np.abs(tt - t64) / np.abs(t64)
        assert (err <= bound).all(), (name, float((err / bound).max()))
    # (camera 9 units away, radii ~0.1: b^2 and 4ac agree to ~3 digits, so the formula's own forward error is ~2e-3 for the
    # typical hit -- the bar is set by the test kernel's float32 sphere formula, not by the traversal)
    assert np.median(bound) < 1e-2
    api.lib().hiprtDestroyFuncTable(ctx.h, table)
    scene.destroy()
    g_mesh.destroy()
    ctx.destroy_geometry(g_spheres)


def test_library_kernels_break_exact_ties_towards_the_smaller_ids(ctx):
    """The four-way exact ties of tests/test_product_traversal_host.py::test_exact_ties_are_broken_towards_the_smaller_ids through
    the library's own trace kernels (csrc/trace_kernels.cu): doubled Cornell box, instanced twice under the same frame."""
    from tests import helpers as H

    verts, _, _ = H.cornell_box()
    doubled = np.concatenate([verts, verts])
    idx = np.arange(192, dtype=np.uint32).reshape(-1, 3)
    rays = H.cornell_rays(160, 160)
    d_rays = ctx.rays_to_device(rays)
    geom = ctx.build(to_dev(ctx, doubled), to_dev(ctx, idx), flags=api.FAST_BUILD | api.DISABLE_PAIRING)
    single = orc.Geometry(verts, np.arange(96, dtype=np.uint32).reshape(-1, 3)).trace(rays, tie_mode=1)
    hit = single["primID"] != orc.INVALID
    got = ctx.hits_to_host(ctx.trace(geom, d_rays))
    assert np.array_equal(got["primID"], single["primID"]) and (got["primID"][hit] < 32).all()
    frame = np.tile(np.asarray([0, 1, 0, 0.3, 1, 1, 1, 5, 0, 0, 0, 0], np.float32), (2, 1))
    scene = ctx.build_scene([geom, geom], frame, frame_type=api.FRAME_SRT)
    got = ctx.hits_to_host(ctx.trace(scene, d_rays))
    h = got["primID"] != api.INVALID
    assert h.mean() > 0.3 and (got["instanceIDs"][h, 0] == 0).all() and (got["primID"][h] < 32).all()
    scene.destroy()
    geom.destroy()
