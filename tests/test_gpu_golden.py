"""The CUDA path against the reference's own golden images (the same image formation as tests/test_oracle_golden.py, hits from
the GPU through the C-ABI instead of the oracle): cornellBox.obj as a two-level scene -- PrimIds, HitDist, Uvs, Normals,
PrimaryRay colours (through the fast, balanced and high-quality builders: BvhFast / BvhBalanced / BvhHighQCornellBox.png are
byte-identical to PrimaryRayCornellBox.png) -- the same scene with every instance translated / scaled / rotated (hiprtFrameSRT),
with instance masks (RayMaskCornellBox) and after hiprtCompactScene; MinimumCornellBox through the bare geometry with the any-hit
iterator of a user kernel: built, compacted (Compaction), saved and loaded (BvhIoApi), batch built (BatchCornellBox); Rebuild0/1
and Update0/1. Tolerance = the reference's (a pixel is bad when a channel differs by more than 10; < 0.3 % bad)."""
import ctypes as C

import numpy as np
import pytest

from hiprt_b200 import api
from oracle import binding as orc
from tests import helpers as H
from tests import test_oracle_golden as G
from tests.test_gpu_parity import to_dev
from tests.test_gpu_user_kernels import Launcher, build_kernels, dev_ptr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def gpu_obj_hits(ctx, srt_frame=None, flags=api.FAST_BUILD, masks=None, compact=False):
    d = np.load(H.os.path.join(H.GOLDEN, "cornell_obj.npz"))
    n = int(d["shape_count"])
    d_verts = to_dev(ctx, d["vertices"])
    geoms = [ctx.build(d_verts, to_dev(ctx, d[f"tris_{k}"]), flags=flags) for k in range(n)]
    if compact:  # hiprtCompactGeometry frees its input and hands back a new handle
        for g in geoms:
            g.h = ctx.compact_geometry(g.h)
    identity = [0, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0]  # hiprtFrameSRT: rotation (axis, angle 0), scale, translation, time
    frames = np.tile(np.asarray(identity if srt_frame is None else srt_frame, np.float32).reshape(1, 12), (n, 1))
    scene = ctx.build_scene(geoms, frames, frame_type=api.FRAME_SRT, masks=masks, flags=flags)
    if compact:
        scene.h = ctx.compact_scene(scene.h)
    hits = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(H.obj_camera_rays()))).copy()
    scene.destroy()
    for g in geoms:
        g.destroy()
    return hits


def test_cornell_obj_images_from_the_gpu(ctx, monkeypatch):
    hits = gpu_obj_hits(ctx)
    hit = hits["primID"] != api.INVALID
    assert hit.all()
    # reuse the oracle tests' image formation with the GPU's hits
    monkeypatch.setattr(G, "_obj_hits", lambda: (hits, hit))
    G.test_prim_ids_cornell_box_matches_reference_image()
    G.test_hit_dist_cornell_box_matches_reference_image()
    G.test_uvs_cornell_box_matches_reference_image()
    G.test_normals_cornell_box_matches_reference_image()
    G._check(G._colour_image(hits), "PrimaryRayCornellBox")


@pytest.mark.parametrize("name, frame", [
    ("TranslateCornellBox", [0, 0, 1, 0, 1, 1, 1, 0, 0, -5, 0, 0]),
    ("ScaleCornellBox", [0, 0, 1, 0, 2, 2, 2, 0, 0, 0, 0, 0]),
    ("RotateCornellBox", [0, 1, 0, 0.5, 1, 1, 1, 0, 0, -3, 0, 0]),
    ("BvhUpdateCornellBox", [0, 1, 0, G.BVH_UPDATE_ANGLE, 1, 1, 1, 0, 0, -3, 0, 0]),
])
def test_transformed_cornell_obj_images_from_the_gpu(ctx, name, frame):
    G._check(G._colour_image(gpu_obj_hits(ctx, frame)), name)


@pytest.mark.parametrize("name, flags", [("BvhBalancedCornellBox", api.BALANCED_BUILD), ("BvhHighQCornellBox", api.HIGH_QUALITY_BUILD)])
def test_cornell_obj_image_does_not_depend_on_the_builder(ctx, name, flags):
    # test/main.cpp:345-412: the same picture through every build flag (geometries and scene)
    G._check(G._colour_image(gpu_obj_hits(ctx, flags=flags)), name)


def test_ray_mask_cornell_obj_image_from_the_gpu(ctx):
    masks = np.array([0 if k % 2 == 0 else 0xFFFFFFFF for k in range(8)], np.uint32)  # hiprtTest.cpp:1215-1218
    G._check(G._colour_image(gpu_obj_hits(ctx, masks=masks)), "RayMaskCornellBox")


def test_compacted_geometries_and_scene_trace_like_the_originals(ctx):
    plain = gpu_obj_hits(ctx)
    compacted = gpu_obj_hits(ctx, compact=True)
    assert plain.tobytes() == compacted.tobytes()
    G._check(G._colour_image(compacted), "PrimaryRayCornellBox")


def all_hit_prims(ctx, fn_all, geom, d_rays, n, max_per=16):
    t = ctx.torch
    d_counts = t.zeros(n, dtype=t.int32, device=ctx.device)
    d_prims = t.full((n, max_per), -1, dtype=t.int32, device=ctx.device)
    d_ts = t.zeros((n, max_per), dtype=t.float32, device=ctx.device)
    Launcher().launch(fn_all, n, 64, [geom, dev_ptr(d_rays), C.c_uint32(n), C.c_uint32(max_per), dev_ptr(d_counts), dev_ptr(d_prims), dev_ptr(d_ts)])
    t.cuda.synchronize()
    assert int(d_counts.max()) <= max_per
    return d_prims.cpu().numpy().view(np.uint32)


def test_minimum_cornell_box_built_compacted_reloaded_and_batch_built(ctx, tmp_path):
    """test/main.cpp:532-795,929-1003 (MinimumCornellBox, Compaction, BatchCornellBox, BvhIoApi): the 32-triangle box, every hit
    along the ray through the any-hit iterator of a user kernel; the same picture from the geometry as built, after
    hiprtCompactGeometry, after hiprtSaveGeometry + hiprtLoadGeometry, and from a hiprtBuildGeometries batch build."""
    verts, _, _ = H.cornell_box()
    d_verts = to_dev(ctx, verts)
    d_idx = to_dev(ctx, np.arange(96, dtype=np.uint32).reshape(-1, 3))
    (fn_all,) = build_kernels(ctx, ["allHitsKernel"])
    rays = H.cornell_rays()
    d_rays, n = ctx.rays_to_device(rays), len(rays)

    def check(handle, name):
        ok, bad, mx = H.image_matches(G.minimum_cornell_image(all_hit_prims(ctx, fn_all, handle, d_rays, n)).reshape(512, 512, 4),
                                      H.load_golden_image(name))
        assert ok and bad == 0.0, (name, bad, mx)

    geom = ctx.build(d_verts, d_idx)
    check(geom.h, "MinimumCornellBox")
    closest = ctx.hits_to_host(ctx.trace(geom, d_rays)).copy()

    path = str(tmp_path / "geom.bin")
    ctx.save_geometry(geom.h, path)
    loaded = ctx.load_geometry(path)
    check(loaded, "BvhIoApi")
    assert ctx.hits_to_host(ctx.trace(loaded, d_rays)).tobytes() == closest.tobytes()
    ctx.destroy_geometry(loaded)

    geom.h = ctx.compact_geometry(geom.h)
    assert int(geom.header()["size"]) == 64 + int(geom.header()["boxNodeCount"]) * 128 + int(geom.header()["primNodeCount"]) * 64
    check(geom.h, "MinimumCornellBox")
    assert ctx.hits_to_host(ctx.trace(geom, d_rays)).tobytes() == closest.tobytes()
    geom.destroy()

    batch = ctx.build_geometries([ctx.mesh_input(d_verts, d_idx)], batch_max=512)
    check(batch[0].h, "MinimumCornellBox")
    ctx.destroy_geometries(batch)


@pytest.mark.parametrize("op, names", [(api.BUILD, ("Rebuild0", "Rebuild1")), (api.UPDATE, ("Update0", "Update1"))])
def test_rebuild_and_update_images_from_the_gpu(ctx, op, names):
    """test/main.cpp:2515-2653: one triangle; its vertices are overwritten in place and the geometry is rebuilt / refitted."""
    d_verts = to_dev(ctx, G.REBUILD_VERTICES[:3].copy())
    geom = ctx.build(d_verts, to_dev(ctx, np.array([[0, 1, 2]], np.uint32)))
    d_rays = ctx.rays_to_device(H.ortho_rays())
    shot = lambda: G.mesh_intersection_image(ctx.hits_to_host(ctx.trace(geom, d_rays))["primID"] != api.INVALID)
    assert G._check(shot(), names[0]) < 0.001
    d_verts.copy_(to_dev(ctx, G.REBUILD_VERTICES[3:].copy()))
    ctx.build_geometry(geom.inp, geom.h, geom.temp, geom.flags, op=op)
    assert G._check(shot(), names[1]) < 0.001
    geom.destroy()


def test_bounding_box_export_and_unimplemented_entry_points(ctx):
    """test/main.cpp:797-842 (BoundingBox): hiprtExportGeometryAabb of the Cornell box = the bounds of its vertices; the scene of
    the OBJ shapes, every instance translated, exports the translated bounds. Scene save / load are NotImplemented as in the
    reference (Context.cpp:855-860), and so is kernel linking from bitcode here."""
    verts, _, _ = H.cornell_box()
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, np.arange(96, dtype=np.uint32).reshape(-1, 3)))
    lo, hi = ctx.export_aabb(geom)
    assert np.array_equal(lo, verts.min(0)) and np.array_equal(hi, verts.max(0))
    geom.destroy()

    d = np.load(H.os.path.join(H.GOLDEN, "cornell_obj.npz"))
    n = int(d["shape_count"])
    d_verts = to_dev(ctx, d["vertices"])
    geoms = [ctx.build(d_verts, to_dev(ctx, d[f"tris_{k}"])) for k in range(n)]
    shift = np.array([1.0, -2.0, 3.0], np.float32)
    frames = np.tile(np.asarray([0, 0, 1, 0, 1, 1, 1, *shift, 0, 0], np.float32).reshape(1, 12), (n, 1))
    scene = ctx.build_scene(geoms, frames, frame_type=api.FRAME_SRT)
    used = d["vertices"][np.unique(np.concatenate([d[f"tris_{k}"].reshape(-1) for k in range(n)]))]
    lo, hi = ctx.export_aabb(scene)
    np.testing.assert_allclose(lo, used.min(0) + shift, rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(hi, used.max(0) + shift, rtol=1e-6, atol=1e-5)
    out = C.c_void_p()
    assert api.lib().hiprtSaveScene(ctx.h, scene.h, b"scene.bin") == api.ERR_NOT_IMPLEMENTED
    assert api.lib().hiprtLoadScene(ctx.h, C.byref(out), b"scene.bin") == api.ERR_NOT_IMPLEMENTED
    assert api.lib().hiprtSetLogLevel(ctx.h, 3) == 0
    scene.destroy()
    for g in geoms:
        g.destroy()
