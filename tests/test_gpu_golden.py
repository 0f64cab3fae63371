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


_KERNELS = {}


def kernels(ctx, func_names, name_sets=None, num_geom_types=0):
    """hiprtBuildTraceKernels once per (functions, function-table layout) for the whole module: every call compiles the test
    kernels' source with NVRTC."""
    key = (tuple(func_names), tuple(name_sets or ()), num_geom_types)
    if key not in _KERNELS:
        _KERNELS[key] = build_kernels(ctx, list(func_names), name_sets, num_geom_types=num_geom_types)
    return _KERNELS[key]


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


def same_hits(a, b):
    """Bit-identical hit records (a miss leaves uv / normal unwritten, as in the reference: only its primID is compared)."""
    hit = a["primID"] != api.INVALID
    return np.array_equal(a["primID"], b["primID"]) and all(a[f][hit].tobytes() == b[f][hit].tobytes() for f in ("t", "uv", "normal"))


def test_compacted_geometries_and_scene_trace_like_the_originals(ctx):
    plain = gpu_obj_hits(ctx)
    compacted = gpu_obj_hits(ctx, compact=True)
    assert same_hits(plain, compacted) and np.array_equal(plain["instanceIDs"][:, 0], compacted["instanceIDs"][:, 0])
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
    fn_all, _ = kernels(ctx, ("allHitsKernel", "aoKernel"))
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
    assert same_hits(ctx.hits_to_host(ctx.trace(loaded, d_rays)), closest)
    ctx.destroy_geometry(loaded)

    geom.h = ctx.compact_geometry(geom.h)
    assert int(geom.header()["size"]) == 64 + int(geom.header()["boxNodeCount"]) * 128 + int(geom.header()["primNodeCount"]) * 64
    check(geom.h, "MinimumCornellBox")
    assert same_hits(ctx.hits_to_host(ctx.trace(geom, d_rays)), closest)
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


# ---- custom primitives and multi-level scenes (test/main.cpp:1296-1385 CustomIntersection, :1485-1670 SceneIntersection,
# :1672-1900 SceneIntersectionMlas): three discs as an AABB list with a custom intersection function, three triangles, both
# instanced in one scene; that scene instanced twice in a top-level scene
CIRCLE_BOXES = np.array([[0.15, 0.40, 0.35, 0.60], [0.40, 0.40, 0.60, 0.60], [0.65, 0.40, 0.85, 0.60]], np.float32)
SCENE_TRIANGLES = np.array([[0.15, 0.40, 0], [0.35, 0.40, 0], [0.25, 0.60, 0], [0.40, 0.40, 0], [0.60, 0.40, 0], [0.50, 0.60, 0],
                            [0.65, 0.40, 0], [0.85, 0.40, 0], [0.75, 0.60, 0]], np.float32)


def unit_square_rays(w=512, h=512):
    x = (np.arange(w, dtype=np.float32) / np.float32(w))[None, :].repeat(h, 0)
    y = (np.arange(h, dtype=np.float32) / np.float32(h))[:, None].repeat(w, 1)
    o = np.stack([x, y, np.full_like(x, -1.0)], -1).reshape(-1, 3)
    return orc.make_rays(o, np.broadcast_to(np.array([0, 0, 1], np.float32), o.shape), max_t=1000.0)


class CircleFixture:
    """The discs' geometry (boxes at height z), the function table with their centres, and the compiled kernels."""

    def __init__(self, ctx, z, wide_stride):
        lo = np.concatenate([CIRCLE_BOXES[:, :2], np.full((3, 1), z, np.float32)], 1)
        hi = np.concatenate([CIRCLE_BOXES[:, 2:], np.full((3, 1), z, np.float32)], 1)
        pad = np.zeros((3, 1), np.float32)
        boxes = np.concatenate([lo, pad, hi, pad], 1) if wide_stride else np.concatenate([lo, hi], 1)  # float4 or float3 pairs
        self.ctx = ctx
        self.inp = ctx.aabb_input(to_dev(ctx, boxes), geom_type=0)
        self.temp = ctx.alloc_temp(ctx.geometry_temp_size(self.inp))
        self.geom = ctx.create_geometry(self.inp)
        ctx.build_geometry(self.inp, self.geom, self.temp)
        self.fn_scene, self.fn_closest = kernels(ctx, ("sceneAllHitsKernel", "customKernel"), [("intersectCircle", None)], num_geom_types=1)
        self.centres = to_dev(ctx, np.array([0.25, 0.5, 0.75], np.float32))
        self.table = C.c_void_p()
        ctx._check(api.lib().hiprtCreateFuncTable(ctx.h, 1, 1, C.byref(self.table)), "hiprtCreateFuncTable")
        ctx._check(api.lib().hiprtSetFuncTable(ctx.h, self.table, 0, 0, api.FuncDataSet(self.centres.data_ptr(), None)), "hiprtSetFuncTable")

    def close(self):
        api.lib().hiprtDestroyFuncTable(self.ctx.h, self.table)
        self.ctx.destroy_geometry(self.geom)


def scene_all_hits_image(ctx, fx, scene, max_per=8):
    """SceneIntersectionKernel's picture: every hit adds the colour of (innermost instance, primitive) to the pixel, in the
    image's own uint8 arithmetic (wraps)."""
    t = ctx.torch
    rays = unit_square_rays()
    n, d_rays = len(rays), ctx.rays_to_device(rays)
    d_counts = t.zeros(n, dtype=t.int32, device=ctx.device)
    d_out = [t.full((n, max_per), -1, dtype=t.int32, device=ctx.device) for _ in range(3)]
    Launcher().launch(fx.fn_scene, n, 64, [scene.h, fx.table, dev_ptr(d_rays), C.c_uint32(n), C.c_uint32(max_per), dev_ptr(d_counts)] + [dev_ptr(d) for d in d_out])
    t.cuda.synchronize()
    counts = d_counts.cpu().numpy()
    assert counts.max() <= max_per
    prims, inst0, inst1 = (d.cpu().numpy().view(np.uint32) for d in d_out)
    colours = np.array([[[1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, 1, 1], [1, 0, 1], [1, 1, 0]]], np.uint8) * np.uint8(255)
    img = np.zeros((n, 4), np.uint8)
    img[:, 3] = 255
    for k in range(max_per):
        valid = k < counts
        inner = np.where(inst1[:, k] != api.INVALID, inst1[:, k], inst0[:, k])
        c = colours[np.where(valid, inner, 0), np.where(valid, prims[:, k], 0)]
        img[:, :3] += np.where(valid[:, None], c, 0).astype(np.uint8)  # uint8 wrap-around, as `image[...] += color` does
    return img, counts, inst0, inst1


def test_custom_intersection_image_from_the_gpu(ctx):
    fx = CircleFixture(ctx, z=0.0, wide_stride=False)
    t = ctx.torch
    rays = unit_square_rays()
    n, d_rays = len(rays), ctx.rays_to_device(rays)
    d_hits = t.zeros((n, 12), dtype=t.int32, device=ctx.device)
    d_payload = t.zeros(1, dtype=t.int32, device=ctx.device)
    Launcher().launch(fx.fn_closest, n, 64, [fx.geom, fx.table, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(n), dev_ptr(d_payload)])
    t.cuda.synchronize()
    prim = ctx.hits_to_host(d_hits)["primID"]
    colours = np.array([[255, 0], [0, 255], [255, 255], [0, 0]], np.uint8)  # intersectCircle's payload colours; last = miss
    img = np.zeros((n, 4), np.uint8)
    img[:, :2] = colours[np.where(prim == api.INVALID, 3, prim)]
    img[:, 3] = 255
    assert (prim != api.INVALID).sum() > 3 * 5000  # three discs of radius 0.1 on a 512 x 512 unit square: ~8200 pixels each
    G._check(img, "CustomIntersection")
    fx.close()


def test_scene_intersection_and_multi_level_images_from_the_gpu(ctx):
    fx = CircleFixture(ctx, z=1.0, wide_stride=True)
    tris = ctx.build(to_dev(ctx, SCENE_TRIANGLES), to_dev(ctx, np.arange(9, dtype=np.uint32).reshape(3, 3)))
    srt = lambda dy: [0, 0, 1, 0, 1, 1, 1, 0, dy, 0, 0, 0]
    mid = ctx.build_scene([fx.geom, tris], np.array([srt(0.05), srt(-0.05)], np.float32), frame_type=api.FRAME_SRT)
    img, counts, inst0, inst1 = scene_all_hits_image(ctx, fx, mid)
    assert (inst1[counts > 0, 0] == api.INVALID).all()  # one level: only instanceIDs[0] is set
    G._check(img, "SceneIntersection")

    top = ctx.build_scene([mid, mid], np.array([srt(0.12), srt(-0.12)], np.float32), frame_type=api.FRAME_SRT,
                          instance_types=[api.INSTANCE_SCENE, api.INSTANCE_SCENE])
    img, counts, inst0, inst1 = scene_all_hits_image(ctx, fx, top)
    hit = counts > 0
    assert hit.any() and (inst0[hit, 0] <= 1).all() and (inst1[hit, 0] <= 1).all()  # top-level instance, then the one inside it
    G._check(img, "SceneIntersectionMlas")
    top.destroy()
    mid.destroy()
    tris.destroy()
    fx.close()


def test_ao_ray_cornell_box_image_from_the_gpu(ctx):
    """test/main.cpp:426-441 + AoRayKernel.h: the workload shape of the headline metric (primary closest-hit + incoherent
    any-hit AO rays) rendered by a user kernel through hiprtSceneTraversalClosest / AnyHit. 128 x 32 rays per pixel against the
    reference's 512 x 32: a pixel's standard error is <= 2 grey levels, the image tolerance (10) is 5 sigma."""
    t = ctx.torch
    d = np.load(H.os.path.join(H.GOLDEN, "cornell_obj.npz"))
    n = int(d["shape_count"])
    d_verts = to_dev(ctx, d["vertices"])
    geoms = [ctx.build(d_verts, to_dev(ctx, d[f"tris_{k}"])) for k in range(n)]
    scene = ctx.build_scene(geoms)
    emissive = sum(1 << k for k in range(n) if d[f"emission_{k}"].sum() > 0)
    assert emissive == 1  # the area light is shape 0
    _, fn = kernels(ctx, ("allHitsKernel", "aoKernel"))
    w = h = 512
    spp, ao_samples = 128, 32
    focal = np.float32(np.float32(0.024) / np.float32(2.0 * np.tan(np.float32(45.0 * np.pi / 180.0) / np.float32(2.0))))
    d_out = t.zeros(w * h, dtype=t.int32, device=ctx.device)
    Launcher().launch(fn, w * h, 64, [scene.h, C.c_uint32(w), C.c_uint32(h), C.c_float(0.0), C.c_float(2.5), C.c_float(5.8), C.c_float(float(focal)),
                                      C.c_uint32(spp), C.c_uint32(ao_samples), C.c_float(1.4), C.c_uint32(emissive), dev_ptr(d_out)])
    t.cuda.synchronize()
    grey = (d_out.cpu().numpy().astype(np.float64) / (spp * ao_samples) * 255.0).astype(np.uint8)
    img = np.zeros((w * h, 4), np.uint8)
    img[:, :3] = grey[:, None]
    img[:, 3] = 255
    golden = H.load_golden_image("AoRayCornellBox").reshape(-1, 4)
    bias = (grey.astype(np.int64) - golden[:, 0].astype(np.int64)).mean()
    assert abs(bias) < 0.4, bias
    G._check(img, "AoRayCornellBox")
    scene.destroy()
    for g in geoms:
        g.destroy()


def test_motion_blur_slerp_image_from_the_gpu(ctx):
    """test/main.cpp:2375-2513 + MotionBlurSlerpKernel: one custom sphere (geomType 1 of 2) instanced with two hiprtFrameMatrix
    key frames, the identity at time 0 and diag(-1, -1, 1) at time 1 -- half a turn about z that only the decomposition into a
    rotation + slerp renders as an arc (a linear blend of the matrices would collapse the sphere onto the axis)."""
    t = ctx.torch
    box = np.array([[-1.5, -0.5, -0.5, -0.5, 0.5, 0.5]], np.float32)
    inp = ctx.aabb_input(to_dev(ctx, box), geom_type=1)
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp))
    geom = ctx.create_geometry(inp)
    ctx.build_geometry(inp, geom, temp)
    m0, m1 = np.zeros(16, np.float32), np.zeros(16, np.float32)  # hiprtFrameMatrix: float matrix[3][4], float time, 3 pad
    m0[[0, 5, 10]] = 1.0
    m1[[0, 5]], m1[10], m1[12] = -1.0, 1.0, 1.0
    scene = ctx.build_scene([geom], np.stack([m0, m1]), frame_type=api.FRAME_MATRIX, transform_headers=np.array([[0, 2]], np.uint32))
    (fn,) = kernels(ctx, ("motionSamplesKernel",), [(None, None), ("intersectSphere", None)], num_geom_types=2)
    spheres = to_dev(ctx, np.array([[-1.0, 0.0, 0.0, 0.5]], np.float32))
    table = C.c_void_p()
    ctx._check(api.lib().hiprtCreateFuncTable(ctx.h, 2, 1, C.byref(table)), "hiprtCreateFuncTable")
    ctx._check(api.lib().hiprtSetFuncTable(ctx.h, table, 1, 0, api.FuncDataSet(spheres.data_ptr(), None)), "hiprtSetFuncTable")
    w = h = 512
    samples = 8192
    d_out = t.zeros(w * h, dtype=t.int32, device=ctx.device)
    Launcher().launch(fn, w * h, 64, [scene.h, table, C.c_uint32(w), C.c_uint32(h), C.c_uint32(samples), dev_ptr(d_out)])
    t.cuda.synchronize()
    frac = (d_out.cpu().numpy().astype(np.float32) / np.float32(samples))[:, None]
    colour = frac * np.array([0.9, 0.1, 0.1], np.float32) + (np.float32(1.0) - frac) * np.float32(0.1)
    img = np.zeros((w * h, 4), np.uint8)
    img[:, :3] = (np.power(colour, np.float32(1.0 / 2.2)) * np.float32(255)).astype(np.uint8)
    img[:, 3] = 255
    assert 0.05 < (frac > 0).mean() < 0.6
    G._check(img, "MotionBlurSlerp")
    api.lib().hiprtDestroyFuncTable(ctx.h, table)
    scene.destroy()
    ctx.destroy_geometry(geom)
