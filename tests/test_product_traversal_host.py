"""CPU-only: the PRODUCT's device-side traversal -- include/hiprt/hiprt_device.h, the traversal objects user kernels and the
library's generic trace kernels instantiate -- built for the host (tests/host/product_device_host.cpp: device qualifiers defined
away, intrinsics given their plain meaning, nothing of the header changed) and run on the CPU over the oracle's node arrays (the
layouts are the product's, include/hiprt/hiprt_b200_layout.h). Held to the oracle hit for hit -- primID / instanceID exact, t, uv and
normal BIT-identical (the triangle and slab arithmetic of both are bit-identical to the reference's, see
tests/test_oracle_vs_reference_math.py) -- and to the reference's golden images. A check of product source without a GPU; what the
GPU adds is the compiler's contraction and the kernels around the objects, which the -m gpu tests cover."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from hiprt_b200 import scenes
from oracle import binding as orc
from tests import helpers as H
from tests import test_oracle_golden as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INCLUDE = "/usr/local/cuda/include"
pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INCLUDE, "vector_types.h")),
                                reason="needs g++ and the CUDA toolkit's vector type headers")
CLOSEST, ANY = 2, 1


@pytest.fixture(scope="module")
def product(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("host") / "libproduct_device_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I", CUDA_INCLUDE,
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "product_device_host.cpp"), "-o", so])
    L = C.CDLL(so)
    L.prod_trace_geometry_all.restype = C.c_uint64
    return L


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def trace_geometry(L, og, rays, kind=CLOSEST):
    boxes, leaves = np.ascontiguousarray(og.box_nodes), np.ascontiguousarray(og.leaves)
    rays = np.ascontiguousarray(rays, orc.RAY_DTYPE)
    hits = np.zeros(len(rays), orc.HIT_DTYPE)
    L.prod_trace_geometry(p(boxes), p(leaves), len(boxes), len(leaves), p(rays), len(rays), p(hits), kind)
    return hits


def trace_scene(L, oscene, rays, kind=CLOSEST, ray_mask=0xFFFFFFFF):
    geoms = oscene.geometries
    gb = [np.ascontiguousarray(g.box_nodes) for g in geoms]
    gl = [np.ascontiguousarray(g.leaves) for g in geoms]
    boxes, leaves = np.ascontiguousarray(oscene.box_nodes), np.ascontiguousarray(oscene.leaves)
    geometry_of_leaf = np.ascontiguousarray(oscene.geometry_of_instance[leaves["bits"] & 0xFFFFFF], np.uint32)
    gb_ptrs = (C.c_void_p * len(geoms))(*[a.ctypes.data for a in gb])
    gl_ptrs = (C.c_void_p * len(geoms))(*[a.ctypes.data for a in gl])
    rays = np.ascontiguousarray(rays, orc.RAY_DTYPE)
    hits = np.zeros(len(rays), orc.HIT_DTYPE)
    L.prod_trace_scene(p(boxes), p(leaves), len(leaves), p(geometry_of_leaf), gb_ptrs, gl_ptrs, len(geoms), p(rays), len(rays), p(hits), kind, ray_mask)
    return hits


def assert_same_hits(got, ref, scene=False):
    hit = ref["primID"] != orc.INVALID
    assert np.array_equal(got["primID"], ref["primID"])
    if scene:
        assert np.array_equal(got["instanceIDs"][hit, 0], ref["instanceIDs"][hit, 0])
    for field in ("t", "uv", "normal"):
        assert np.array_equal(got[field][hit].view(np.uint32), ref[field][hit].view(np.uint32)), field


@pytest.mark.parametrize("mesh", ["cornell", "sphere", "soup"])
def test_geometry_traversal_objects_return_the_oracles_hits(product, mesh):
    if mesh == "cornell":
        verts, _, _ = H.cornell_box()
        og, rays = orc.Geometry(verts, np.arange(96, dtype=np.uint32).reshape(-1, 3)), H.cornell_rays(192, 192)
    elif mesh == "sphere":
        v, i = scenes.sphere_mesh(96, 48)
        og, rays = orc.Geometry(v, i), scenes.pinhole_rays(160, 160, (0, 0, 3), (0, 0, 0))
    else:
        v = scenes.random_soup(3000, seed=3)  # non-indexed: three vertices per triangle
        og, rays = orc.Geometry(v), scenes.random_rays(20000, [-1, -1, -1], [1, 1, 1], seed=5)
    ref = og.trace(rays, tie_mode=1)  # the order-independent closest-hit rule the product implements
    assert 0.05 < (ref["primID"] != orc.INVALID).mean()
    assert_same_hits(trace_geometry(product, og, rays), ref)
    occluded = trace_geometry(product, og, rays, ANY)["primID"] != orc.INVALID
    assert np.array_equal(occluded, ref["primID"] != orc.INVALID)


def test_any_hit_iterator_enumerates_the_oracles_hits(product):
    verts, _, _ = H.cornell_box()
    og = orc.Geometry(verts, np.arange(96, dtype=np.uint32).reshape(-1, 3))
    rays = H.cornell_rays()
    boxes, leaves = np.ascontiguousarray(og.box_nodes), np.ascontiguousarray(og.leaves)
    counts, prims = np.zeros(len(rays), np.uint32), np.full((len(rays), 16), orc.INVALID, np.uint32)
    product.prod_trace_geometry_all(p(boxes), p(leaves), len(boxes), len(leaves), p(rays), len(rays), 16, p(counts), p(prims))
    ocounts, oprims, _ = og.trace_all(rays, max_per_ray=16)
    assert np.array_equal(counts, ocounts) and np.array_equal(np.sort(prims, 1), np.sort(oprims, 1))
    # ... which is the reference's MinimumCornellBox picture
    ok, bad, mx = H.image_matches(G.minimum_cornell_image(prims).reshape(512, 512, 4), H.load_golden_image("MinimumCornellBox"))
    assert ok and bad == 0.0


@pytest.mark.parametrize("name, frame", [
    ("PrimaryRayCornellBox", None),
    ("RotateCornellBox", [0, 1, 0, 0.5, 1, 1, 1, 0, 0, -3, 0, 0]),
    ("ScaleCornellBox", [0, 0, 1, 0, 2, 2, 2, 0, 0, 0, 0, 0]),
])
def test_scene_traversal_objects_return_the_oracles_hits_and_the_reference_images(product, name, frame):
    oscene, _ = H.cornell_obj_scene(frame)
    rays = H.obj_camera_rays()
    got = trace_scene(product, oscene, rays)
    assert_same_hits(got, oscene.trace(rays, tie_mode=1), scene=True)
    G._check(G._colour_image(got), name)
    occluded = trace_scene(product, oscene, rays, ANY)["primID"] != orc.INVALID
    assert np.array_equal(occluded, got["primID"] != orc.INVALID)


def test_scene_traversal_respects_instance_masks(product):
    d = np.load(os.path.join(H.GOLDEN, "cornell_obj.npz"))
    n = int(d["shape_count"])
    geoms = [orc.Geometry(d["vertices"], d[f"tris_{k}"]) for k in range(n)]
    masks = np.array([0 if k % 2 == 0 else 0xFFFFFFFF for k in range(n)], np.uint32)
    oscene = orc.Scene(geoms, np.arange(n), None, 0, masks=masks)
    got = trace_scene(product, oscene, H.obj_camera_rays())
    G._check(G._colour_image(got), "RayMaskCornellBox")
    none = trace_scene(product, oscene, H.obj_camera_rays(64, 64), ray_mask=0)
    assert (none["primID"] == orc.INVALID).all()


# ---- custom primitives through the function table, scenes instanced in scenes -----------------------------------------------------
class HostObject(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("geomType", C.c_uint32), ("boxNodes", C.c_void_p), ("leaves", C.c_void_p), ("leafCount", C.c_uint32),
                ("pad", C.c_uint32), ("childOfLeaf", C.c_void_p), ("frames", C.c_void_p)]


def flatten(obj, table, keep):
    """Depth-first object table for prod_trace_general (children before parents); returns the index of `obj`."""
    if id(obj) in table["index"]:
        return table["index"][id(obj)]
    boxes, leaves = np.ascontiguousarray(obj.box_nodes), np.ascontiguousarray(obj.leaves)
    keep += [boxes, leaves]
    if isinstance(obj, orc.Scene):
        children = [flatten(g, table, keep) for g in obj.geometries]
        child_of_leaf = np.ascontiguousarray(np.asarray(children, np.uint32)[obj.geometry_of_instance[leaves["bits"] & 0xFFFFFF]])
        keep.append(child_of_leaf)
        rec = HostObject(2, 0, boxes.ctypes.data, leaves.ctypes.data, len(leaves), 0, child_of_leaf.ctypes.data, None)
    elif isinstance(obj, orc.AabbGeometry):
        rec = HostObject(1, obj.geom_type, boxes.ctypes.data, leaves.ctypes.data, len(leaves), 0, None, None)
    else:
        # (no user geometry type unless the test gives one: 0x7FFFFFFF << 1 | 1 is the "invalid" type of a plain triangle mesh)
        rec = HostObject(0, getattr(obj, "host_geom_type", 0x7FFFFFFF), boxes.ctypes.data, leaves.ctypes.data, len(leaves), 0, None, None)
    table["objects"].append(rec)
    table["index"][id(obj)] = len(table["objects"]) - 1
    return table["index"][id(obj)]


def trace_general(L, obj, rays, intersectors=(), data=(), max_per_ray=0):
    table, keep = {"objects": [], "index": {}}, []
    root = flatten(obj, table, keep)
    objects = (HostObject * len(table["objects"]))(*table["objects"])
    kinds = np.asarray(list(intersectors) or [0], np.uint32)
    arrays = [np.ascontiguousarray(a) for a in data] or [np.zeros(1, np.float32)]
    ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    rays = np.ascontiguousarray(rays, orc.RAY_DTYPE)
    n = len(rays)
    counts = np.zeros(n, np.uint32)
    hits = np.zeros((n, max(max_per_ray, 1)), orc.HIT_DTYPE)
    hits["primID"] = orc.INVALID
    L.prod_trace_general(objects, len(objects), root, len(kinds), p(kinds), ptrs, p(rays), n, C.c_float(0.0), max_per_ray, p(counts), p(hits))
    return hits[:, 0] if max_per_ray == 0 else (counts, hits)


def test_custom_primitives_and_multi_level_scenes_reproduce_the_reference_images(product):
    """CustomIntersection, SceneIntersection and SceneIntersectionMlas through the product's traversal objects on the CPU: the
    function table, custom leaves next to triangles, the re-entrant any-hit scene iterator and the instance stack."""
    from tests import test_oracle_custom as T

    rays = T.unit_square_rays()
    prim = trace_general(product, T.circles(0.0), rays, intersectors=[1], data=[T.CENTRES])["primID"]
    colours = np.array([[255, 0], [0, 255], [255, 255], [0, 0]], np.uint8)
    img = np.zeros((len(prim), 4), np.uint8)
    img[:, :2] = colours[np.where(prim == orc.INVALID, 3, prim)]
    img[:, 3] = 255
    assert G._check(img, "CustomIntersection") == 0.0

    srt = lambda dy: [0, 0, 1, 0, 1, 1, 1, 0, dy, 0, 0, 0]
    tris = orc.Geometry(T.TRIANGLES, np.arange(9, dtype=np.uint32).reshape(3, 3))
    mid = orc.Scene([T.circles(1.0), tris], [0, 1], np.array([srt(0.05), srt(-0.05)], np.float32), 0)
    counts, hits = trace_general(product, mid, rays, intersectors=[1], data=[T.CENTRES], max_per_ray=4)
    assert counts.max() <= 2
    assert G._check(T.scene_image(counts, hits), "SceneIntersection") == 0.0

    top = orc.Scene([mid], [0, 0], np.array([srt(0.12), srt(-0.12)], np.float32), 0)
    counts, hits = trace_general(product, top, rays, intersectors=[1], data=[T.CENTRES], max_per_ray=8)
    assert G._check(T.scene_image(counts, hits), "SceneIntersectionMlas") == 0.0
    # the same hits as the oracle's walker, level by level
    ocounts, ohits = orc.trace_general(top, rays, T.intersect_circle, max_per_ray=8)
    assert np.array_equal(counts, ocounts)
    key = lambda h: np.sort(h["instanceIDs"][:, :, 0].astype(np.int64) * 100 + h["instanceIDs"][:, :, 1].astype(np.int64) * 10 + h["primID"], 1)
    assert np.array_equal(key(hits), key(ohits))


def test_three_level_scene_closest_hits_equal_the_oracles(product):
    meshes = scenes.instanced_geometries(3)
    geoms = [orc.Geometry(v, i) for v, i in meshes]
    leaf = orc.Scene(geoms, [k % 3 for k in range(6)], scenes.random_matrix_frames(6, seed=5, extent=2.0, scale_range=(0.6, 1.2)), 1)
    mid = orc.Scene([leaf], [0] * 4, scenes.random_matrix_frames(4, seed=6, extent=6.0, scale_range=(0.8, 1.1)), 1)
    top = orc.Scene([mid], [0] * 3, scenes.random_matrix_frames(3, seed=7, extent=14.0, scale_range=(0.9, 1.0)), 1)
    rays = scenes.pinhole_rays(160, 160, (0.0, 4.0, 45.0), (0, 0, 0))
    got = trace_general(product, top, rays)
    ref = orc.trace_general(top, rays, tie_mode=1)
    hit = ref["primID"] != orc.INVALID
    assert hit.mean() > 0.03 and np.array_equal(got["primID"], ref["primID"])
    assert np.array_equal(got["instanceIDs"][hit], ref["instanceIDs"][hit])
    assert np.array_equal(got["t"][hit].view(np.uint32), ref["t"][hit].view(np.uint32))


def test_animated_instance_reproduces_the_motion_blur_slerp_image(product):
    """MotionBlurSlerp through the product's traversal objects on the CPU: a custom sphere (geometry type 1 of 2) under an ANIMATED
    instance node -- the node carries the key-frame range, the scene header the 64-byte key frames (here the reference's own
    decomposition of the identity and diag(-1, -1, 1), tests/golden/ref_math.npz), the instance box is the motion bound -- traced at
    128 evenly spaced times per pixel of every fourth row and column (the reference draws 512 random times: its own sampling noise
    is what remains)."""
    from tests.golden import make_ref_math as M
    from tests.test_oracle_ao import randf

    f32 = np.float32
    stored = np.ascontiguousarray(np.load(os.path.join(M.HERE, "ref_math.npz"))["motion_stored_12"])
    assert stored.shape == (2, 16) and stored[1, 13] == 1.0 and np.array_equal(stored[1, :4], np.array([0, 0, 1, 0], f32))
    sphere = orc.AabbGeometry(np.array([[-1.5, -0.5, -0.5, -0.5, 0.5, 0.5]], f32), geom_type=1)
    scene = orc.Scene([sphere], [0], None, 0)
    boxes, leaves = np.ascontiguousarray(scene.box_nodes), np.ascontiguousarray(scene.leaves)
    # animated instance node: {frameIndex, frameCount} in place of the matrix, neither static nor identity (bits 25, 26)
    leaves["invMatrix"][0, 0, :2] = np.array([0, 2], np.uint32).view(f32)
    leaves["bits"][0] &= ~np.uint32((1 << 25) | (1 << 26))
    bound = np.zeros((1, 6), f32)
    keys = np.zeros((2, 16), f32)
    keys[0, [0, 5, 10]] = 1.0
    keys[1, [0, 5]], keys[1, 10], keys[1, 12] = -1.0, 1.0, 1.0
    orc.lib().orc_math_motion_bounds(2, p(keys), 1, 1, p(sphere.aabbs), p(bound))
    boxes["box"][0, 0] = bound.reshape(2, 3)
    gboxes, gleaves = np.ascontiguousarray(sphere.box_nodes), np.ascontiguousarray(sphere.leaves)
    child = np.zeros(1, np.uint32)
    objects = (HostObject * 2)(HostObject(1, 1, gboxes.ctypes.data, gleaves.ctypes.data, 1, 0, None, None),
                               HostObject(2, 0, boxes.ctypes.data, leaves.ctypes.data, 1, 0, child.ctypes.data, stored.ctypes.data))
    kinds = np.array([0, 2], np.uint32)
    spheres = np.array([[-1.0, 0.0, 0.0, 0.5]], f32)
    ptrs = (C.c_void_p * 2)(None, spheres.ctypes.data)

    res, stride, samples = 512, 4, 128
    ys, xs = np.meshgrid(np.arange(0, res, stride, dtype=np.uint32), np.arange(0, res, stride, dtype=np.uint32), indexing="ij")
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    seed = H.tea16(xs, ys)[0].astype(np.uint32)  # MotionBlurSlerpKernel's jitter (HiprtTestKernel.h:373-383)
    dx = f32(2.0) * ((xs.astype(f32) + randf(seed)) / f32(res)) - f32(1.0)
    dy = f32(2.0) * ((ys.astype(f32) + randf(seed)) / f32(res)) - f32(1.0)
    d = np.stack([dx * f32(1.576494), dy * f32(1.576494), np.full_like(dx, -5.0)], -1).astype(f32)
    d = (d / np.sqrt((d * d).sum(1, dtype=f32), dtype=f32)[:, None]).astype(f32)
    rays = orc.make_rays(np.broadcast_to(np.array([0, 0, 5], f32), d.shape), d)
    n = len(rays)
    count = np.zeros(n, np.int64)
    counts, hits = np.zeros(n, np.uint32), np.zeros(n, orc.HIT_DTYPE)
    for k in range(samples):
        product.prod_trace_general(objects, 2, 1, 2, p(kinds), ptrs, p(rays), n, C.c_float((k + 0.5) / samples), 0, p(counts), p(hits))
        count += hits["primID"] != orc.INVALID
    frac = (count / samples).astype(f32)[:, None]
    colour = frac * np.array([0.9, 0.1, 0.1], f32) + (f32(1.0) - frac) * f32(0.1)
    img = (np.power(colour, f32(1.0 / 2.2)) * f32(255)).astype(np.uint8)
    golden = H.load_golden_image("MotionBlurSlerp")[::stride, ::stride, :3].reshape(-1, 3)
    diff = np.abs(img.astype(np.int64) - golden.astype(np.int64)).max(1)
    assert 0.3 < (count > 0).mean() < 0.5
    assert (diff > 10).mean() < 0.005, ((diff > 10).mean(), diff.max())


@pytest.mark.parametrize("dynamic", [0, 1])
@pytest.mark.parametrize("shared_size", [0, 1, 2, 16])
def test_global_and_dynamic_stacks_spill_and_refill_correctly(product, dynamic, shared_size):
    """hiprtGlobalStack / hiprtDynamicStack (hiprt_device.h; reference hiprt_device_impl.h:86-247): with a "shared" ring of 0, 1, 2
    or 16 entries in front of the global region, every push beyond the ring spills the oldest entry and every pop past it refills
    from the global region; the hits must be those of the private stack, bit for bit, and the dynamic variant must release its
    warp slot."""
    v, i = scenes.sphere_mesh(128, 64)
    og = orc.Geometry(v, i)
    rays = scenes.random_rays(6000, [-1.2, -1.2, -1.2], [1.2, 1.2, 1.2], seed=9)
    ref = trace_geometry(product, og, rays)
    boxes, leaves = np.ascontiguousarray(og.box_nodes), np.ascontiguousarray(og.leaves)
    hits = np.zeros(len(rays), orc.HIT_DTYPE)
    state = C.c_uint32(0)
    product.prod_trace_geometry_global_stack(p(boxes), p(leaves), len(boxes), len(leaves), p(rays), len(rays), p(hits), shared_size, 64, dynamic, C.byref(state))
    assert state.value == 1, state.value  # hiprtTraversalStateFinished (1) for every ray: no overflow (3), locks released
    assert (ref["primID"] != orc.INVALID).mean() > 0.3
    assert_same_hits(hits, ref)


def test_hit_filter_on_triangles_reproduces_the_cutout_image(product):
    """Cutout (test/main.cpp:1219-1290): the unit quad with a user geometry type, closest hit with the checkerboard hit filter in
    function-table slot 0 -- the filter runs inside the product's traversal, a rejected hit is not reported."""
    v = np.array([[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0]], np.float32)
    quad = orc.Geometry(v, np.array([[0, 1, 2], [0, 2, 3]], np.uint32))
    quad.host_geom_type = 0
    rays = H.ortho_rays()
    product.prod_set_filter(0, 1)
    try:
        filtered = trace_general(product, quad, rays, intersectors=[0], data=[np.zeros(1, np.float32)])
    finally:
        product.prod_set_filter(0, 0)
    plain = trace_general(product, quad, rays, intersectors=[0], data=[np.zeros(1, np.float32)])
    visible = filtered["primID"] != orc.INVALID
    assert (plain["primID"] != orc.INVALID).mean() > 0.99 and 0.4 < visible.mean() < 0.6
    img = np.zeros((len(rays), 4), np.uint8)
    img[:, :3] = np.where(visible, 255, 0)[:, None]
    img[:, 3] = 255
    G._check(img, "Cutout")


@pytest.mark.parametrize("two_levels", [0, 1])
def test_transform_queries_are_the_references(product, two_levels):
    """hiprtPoint/VectorObjectToWorld/WorldToObject of the device API (reference: hiprt_device.h:508-583) against the reference's
    SRTFrame::transform / invTransform / transformVector / invTransformVector under Transform::interpolateFrames, bit for bit
    (tests/golden/ref_math.npz), through the single-level signature and through the multi-level one with an identity outer level."""
    from tests.golden import make_ref_math as M

    golden = np.load(os.path.join(M.HERE, "ref_math.npz"))
    inputs = M.ref_math_inputs()
    pts = np.ascontiguousarray(inputs["motion_rays"].view(M.RAY).reshape(-1)["direction"][:100])
    for j in range(int(inputs["motion_count"])):
        stored = np.ascontiguousarray(golden[f"motion_stored_{j}"])
        for k, time in enumerate(M.POINT_TIMES):
            out = np.zeros((len(pts), 12), np.float32)
            product.prod_transform_queries(p(stored), len(stored), C.c_float(time), len(pts), p(pts), p(out), two_levels)
            assert np.array_equal(out.view(np.uint32), golden[f"motion_points_{j}"][k].view(np.uint32)), (j, time)


def test_exact_ties_are_broken_towards_the_smaller_ids(product):
    """"Ties broken deterministically": every triangle of the Cornell box exists twice (primitives p and p + 32 at the same place)
    and the doubled geometry is instanced twice under the same frame, so EVERY hit is a four-way exact tie in t. The product's
    traversal objects and the oracle's order-independent rule (min over t, instanceID, primID) must both report instance 0 and
    the lower primitive, whatever order the BVH visits them in."""
    verts, _, _ = H.cornell_box()
    doubled = np.concatenate([verts, verts])
    og = orc.Geometry(doubled, np.arange(192, dtype=np.uint32).reshape(-1, 3), allow_pairing=False)
    rays = H.cornell_rays(160, 160)
    single = orc.Geometry(verts, np.arange(96, dtype=np.uint32).reshape(-1, 3)).trace(rays, tie_mode=1)
    hit = single["primID"] != orc.INVALID
    for hits in (trace_geometry(product, og, rays), og.trace(rays, tie_mode=1)):
        assert np.array_equal(hits["primID"], single["primID"]) and (hits["primID"][hit] < 32).all()
        assert np.array_equal(hits["t"][hit].view(np.uint32), single["t"][hit].view(np.uint32))
    frame = np.tile(np.asarray([0, 1, 0, 0.3, 1, 1, 1, 5, 0, 0, 0, 0], np.float32), (2, 1))
    oscene = orc.Scene([og], [0, 0], frame, 0)
    one = orc.Scene([og], [0], frame[:1], 0).trace(rays, tie_mode=1)
    for hits in (trace_scene(product, oscene, rays), oscene.trace(rays, tie_mode=1)):
        h = hits["primID"] != orc.INVALID
        assert h.mean() > 0.3 and np.array_equal(hits["primID"], one["primID"])
        assert (hits["instanceIDs"][h, 0] == 0).all() and (hits["primID"][h] < 32).all()
