"""Pin the CPU oracle against the reference's own golden images (CPU only)."""
import numpy as np
import pytest

from oracle import binding as orc
from tests import helpers as H


def minimum_cornell_image(prims):
    """CornellBoxKernel (HiprtTestKernel.h:118-165): a third of the diffuse colour of every primitive along the ray, summed.
    prims: uint32 [rays, k] primitive ids of all hits of a ray, INVALID-padded."""
    _, mats, diffuse = H.cornell_box()
    col = (np.float32(1.0 / 3.0) * diffuse * np.float32(255)).astype(np.uint32)  # per material
    img = np.zeros((len(prims), 4), np.uint32)
    for k in range(prims.shape[1]):
        valid = prims[:, k] != orc.INVALID
        c = col[mats[np.where(valid, prims[:, k], 0)]]
        img[:, :3] += np.where(valid[:, None], c, 0)
    img = np.minimum(img, 255)
    img[:, 3] = 255
    return img


def test_minimum_cornell_box_matches_reference_image():
    # reference test: test/main.cpp:532-619 (any-hit iterator drained, colours summed)
    verts, _, _ = H.cornell_box()
    idx = np.arange(96, dtype=np.uint32).reshape(-1, 3)
    g = orc.Geometry(verts, idx)
    assert g.prim_count == 32  # iota indices share no vertex index: nothing pairs
    rays = H.cornell_rays()
    counts, prims, _ = g.trace_all(rays, max_per_ray=16)
    assert counts.max() <= 16
    img = minimum_cornell_image(prims)
    # CustomBvhImport.png and BvhIoApi.png are byte-identical copies of this image (tests/golden/make_golden.py)
    for name in ("MinimumCornellBox", "CustomBvhImport", "BvhIoApi"):
        ok, bad, mx = H.image_matches(img.reshape(512, 512, 4), H.load_golden_image(name))
        assert ok, (bad, mx)
        assert bad == 0.0, (bad, mx)


def test_mesh_intersection_matches_reference_image():
    # reference test: test/main.cpp:1016-1070 + MeshIntersectionKernel
    v = np.array([[0, 0, 0], [1, 0, 0], [0.5, 1, 0]], np.float32)
    g = orc.Geometry(v, np.array([[0, 1, 2]], np.uint32))
    assert g.prim_count == 1 and g.box_count == 1
    hits = g.trace(H.ortho_rays(), tie_mode=0)
    hit = hits["primID"] != orc.INVALID
    w = h = 512
    x = np.arange(w, dtype=np.float32)[None, :].repeat(h, 0).reshape(-1)
    y = np.arange(h, dtype=np.float32)[:, None].repeat(w, 1).reshape(-1)
    img = np.zeros((w * h, 4), np.uint8)
    img[:, 0] = np.where(hit, (x / np.float32(w) * np.float32(255)).astype(np.uint8), 0)
    img[:, 1] = np.where(hit, (y / np.float32(h) * np.float32(255)).astype(np.uint8), 0)
    img[:, 3] = 255
    ok, bad, mx = H.image_matches(img.reshape(h, w, 4), H.load_golden_image("MeshIntersection"))
    # the only differing pixels are the 255 rays that graze the edge x = 1 - y/2 exactly (the golden
    # image comes from hardware intersection, which drops them; Moller-Trumbore keeps u+v == 1)
    assert ok and bad < 0.001, (bad, mx)


def test_closest_hit_tie_modes_agree_off_ties():
    verts, _, _ = H.cornell_box()
    g = orc.Geometry(verts, np.arange(96, dtype=np.uint32).reshape(-1, 3))
    rays = H.cornell_rays(128, 128)
    a = g.trace(rays, tie_mode=0)
    b = g.trace(rays, tie_mode=1)
    same = a["primID"] == b["primID"]
    # where they differ the two candidates are at exactly the same distance
    assert np.all(a["t"][~same] == b["t"][~same])
    assert same.mean() > 0.99


# ---- the cornellBox.obj family (test/main.cpp:465-520, test/kernels/PrimaryRayKernel.h): closest-hit primary rays through a
# two-level scene, coloured by primitive / instance id, hit distance, barycentrics and normal. These images pin, pixel by
# pixel, WHICH primitive and instance the reference reports, its t, its uv convention and its normal orientation.
def _obj_hits():
    scene, geoms = H.cornell_obj_scene()
    hits = scene.trace(H.obj_camera_rays(), tie_mode=0)
    return hits, hits["primID"] != orc.INVALID


def _check(img, name, max_bad=0.003):
    ok, bad, mx = H.image_matches(img.reshape(512, 512, 4), H.load_golden_image(name), max_bad_frac=max_bad)
    assert ok, (name, bad, mx)
    return bad


def test_prim_ids_cornell_box_matches_reference_image():
    hits, hit = _obj_hits()
    prim, inst = hits["primID"], hits["instanceIDs"][:, 0]
    img = np.zeros((len(hits), 4), np.uint8)
    img[:, 0] = np.where(hit, H.tea16(prim, np.zeros_like(prim))[0] % 255, 0)
    img[:, 1] = np.where(hit, H.tea16(inst, np.zeros_like(inst))[0] % 255, 0)
    img[:, 2] = np.where(hit, H.tea16(inst, prim)[0] % 255, 0)
    img[:, 3] = 255
    bad = _check(img, "PrimIdsCornellBox")
    assert bad < 0.001, bad  # ids are all-or-nothing per pixel: only silhouette pixels may differ


def test_hit_dist_cornell_box_matches_reference_image():
    hits, hit = _obj_hits()
    v = np.clip((hits["t"] / np.float32(50.0) * np.float32(255)).astype(np.int64), 0, 255)
    img = np.zeros((len(hits), 4), np.uint8)
    img[:, :3] = np.where(hit, v, 0)[:, None]
    img[:, 3] = 255
    _check(img, "HitDistCornellBox")


def test_uvs_cornell_box_matches_reference_image():
    hits, hit = _obj_hits()
    img = np.zeros((len(hits), 4), np.uint8)
    img[:, 0] = np.where(hit, np.clip((hits["uv"][:, 0] * np.float32(255)).astype(np.int64), 0, 255), 0)
    img[:, 1] = np.where(hit, np.clip((hits["uv"][:, 1] * np.float32(255)).astype(np.int64), 0, 255), 0)
    img[:, 3] = 255
    _check(img, "UvsCornellBox")


def test_normals_cornell_box_matches_reference_image():
    hits, hit = _obj_hits()
    n = hits["normal"].astype(np.float32)
    ln = np.sqrt((n * n).sum(1))
    n = n / np.maximum(ln, 1e-30)[:, None]  # identity instance frames: object-to-world leaves the vector unchanged
    img = np.zeros((len(hits), 4), np.uint8)
    img[:, :3] = np.where(hit[:, None], ((n + np.float32(1.0)) * np.float32(0.5) * np.float32(255)).astype(np.int64), 0)
    img[:, 3] = 255
    _check(img, "NormalsCornellBox")


def test_pair_triangles_matches_reference_image():
    # test/main.cpp:1122-1215 + PairTrianglesKernel: three stacked triangles (iota indices: nothing pairs), any-hit iterator
    # drained, one colour channel per primitive
    v = np.array([[0, 0, 1], [1, 0, 1], [0.5, 1, 1], [0, 0, 0.5], [1, 0, 0.5], [0.5, 1, 0.5], [0, 0, 0], [1, 0, 0], [0.5, 1, 0]], np.float32)
    g = orc.Geometry(v, np.arange(9, dtype=np.uint32).reshape(3, 3))
    counts, prims, _ = g.trace_all(H.ortho_rays(), max_per_ray=4)
    assert counts.max() <= 3
    img = np.zeros((512 * 512, 4), np.uint32)
    for k in range(prims.shape[1]):
        valid = prims[:, k] != orc.INVALID
        for p in range(3):
            img[:, p] += np.where(valid & (prims[:, k] == p), 255, 0).astype(np.uint32)
    img = np.minimum(img, 255)
    img[:, 3] = 255
    _check(img.astype(np.uint8), "PairTriangles")


def test_scene_intersection_singleton_matches_reference_image():
    # test/main.cpp:1387-1480 + SceneIntersectionSingleton kernel: two triangles, one instance scaled by 0.5 (hiprtFrameSRT)
    s = 0.5
    v = np.array([[0.5, s, 0], [0.5 + s * 0.86, -s * 0.5, 0], [0.5 - s * 0.86, -s * 0.5, 0],
                  [-0.5, s, 0], [-0.5 + s * 0.86, -s * 0.5, 0], [-0.5 - s * 0.86, -s * 0.5, 0]], np.float32)
    g = orc.Geometry(v, np.arange(6, dtype=np.uint32).reshape(2, 3))
    frame = np.array([[0, 0, 1, 0, 0.5, 0.5, 0.5, 0, 0, 0, 0, 0]], np.float32)  # rotation (axis, angle 0), scale, translation, time
    scene = orc.Scene([g], [0], frame, 0)
    w = h = 512
    x = (np.arange(w, dtype=np.float32) / np.float32(w))[None, :].repeat(h, 0)
    y = (np.arange(h, dtype=np.float32) / np.float32(h))[:, None].repeat(w, 1)
    o = np.stack([x - np.float32(0.5), y - np.float32(0.5), np.full_like(x, -1.0)], -1).reshape(-1, 3)
    rays = orc.make_rays(o, np.broadcast_to(np.array([0, 0, 1], np.float32), o.shape), max_t=1000.0)
    hit = scene.trace(rays)["primID"] != orc.INVALID
    img = np.zeros((w * h, 4), np.uint8)
    img[:, 0] = np.where(hit, (x.reshape(-1) * np.float32(255)).astype(np.uint8), 0)
    img[:, 1] = np.where(hit, (y.reshape(-1) * np.float32(255)).astype(np.uint8), 0)
    img[:, 3] = 255
    _check(img, "SceneIntersectionSingleton")


def _colour_image(hits):
    hit = hits["primID"] != orc.INVALID
    colours = H.cornell_obj_face_colours()
    img = np.zeros((len(hits), 4), np.uint8)
    inst, prim = hits["instanceIDs"][:, 0], hits["primID"]
    for k, c in enumerate(colours):
        sel = hit & (inst == k)
        img[sel, :3] = (c[prim[sel]] * np.float32(255)).astype(np.uint8)
    img[:, 3] = 255
    return img


def _accumulated_angle(steps):
    a = np.float32(0.0)
    for _ in range(steps):
        a = np.float32(float(a) + 0.1)  # `float angle; angle += 0.1` of the reference's loop
    return float(a)


BVH_UPDATE_ANGLE = _accumulated_angle(7)


@pytest.mark.parametrize("name, frame", [
    ("PrimaryRayCornellBox", None),
    ("TranslateCornellBox", [0, 0, 1, 0, 1, 1, 1, 0, 0, -5, 0, 0]),
    ("ScaleCornellBox", [0, 0, 1, 0, 2, 2, 2, 0, 0, 0, 0, 0]),
    ("RotateCornellBox", [0, 1, 0, 0.5, 1, 1, 1, 0, 0, -3, 0, 0]),
    # test/main.cpp:301-343: the last of eight scenes rotated in steps of 0.1 (float accumulation of the double 0.1: 0.70000005)
    ("BvhUpdateCornellBox", [0, 1, 0, BVH_UPDATE_ANGLE, 1, 1, 1, 0, 0, -3, 0, 0]),
])
def test_instanced_cornell_box_colour_images_match_reference(name, frame):
    """test/main.cpp:252-300,465-475: every instance of the OBJ scene carries the same hiprtFrameSRT (translation / scale /
    rotation about y); painted with the face's diffuse colour. Pins the oracle's instance transform against the reference's."""
    scene, _ = H.cornell_obj_scene(frame)
    hits = scene.trace(H.obj_camera_rays(), tie_mode=0)
    _check(_colour_image(hits), name)


def test_ray_mask_cornell_box_matches_reference_image():
    """RayMaskCornellBox.png: the OBJ scene set up with enableRayMask (hiprtTest.cpp:1215-1218: every even shape gets instance
    mask 0, the others ~0) and traced with hiprtFullRayMask: masked instances must vanish (the light, the ceiling, the left wall
    and the short box), and what was behind them shows."""
    d = np.load(H.os.path.join(H.GOLDEN, "cornell_obj.npz"))
    n = int(d["shape_count"])
    geoms = [orc.Geometry(d["vertices"], d[f"tris_{k}"]) for k in range(n)]
    masks = np.array([0 if k % 2 == 0 else 0xFFFFFFFF for k in range(n)], np.uint32)
    scene = orc.Scene(geoms, np.arange(n), None, 0, masks=masks)
    hits = scene.trace(H.obj_camera_rays(), tie_mode=0)
    assert not np.isin(hits["instanceIDs"][hits["primID"] != orc.INVALID, 0], [0, 2, 4, 6]).any()
    _check(_colour_image(hits), "RayMaskCornellBox")


REBUILD_VERTICES = np.array([[0.15, 0.40, 0.0], [0.35, 0.40, 0.0], [0.25, 0.60, 0.0],
                             [0.65, 0.40, 0.0], [0.85, 0.40, 0.0], [0.75, 0.60, 0.0]], np.float32)


def mesh_intersection_image(hit, w=512, h=512):
    """MeshIntersectionKernel (HiprtTestKernel.h): red / green = pixel position where the ray hits, black elsewhere."""
    x = np.arange(w, dtype=np.float32)[None, :].repeat(h, 0).reshape(-1)
    y = np.arange(h, dtype=np.float32)[:, None].repeat(w, 1).reshape(-1)
    img = np.zeros((w * h, 4), np.uint8)
    img[:, 0] = np.where(hit, (x / np.float32(w) * np.float32(255)).astype(np.uint8), 0)
    img[:, 1] = np.where(hit, (y / np.float32(h) * np.float32(255)).astype(np.uint8), 0)
    img[:, 3] = 255
    return img


@pytest.mark.parametrize("names, first_vertex", [(("Rebuild0", "Update0"), 0), (("Rebuild1", "Update1"), 3)])
def test_rebuild_and_update_images_match_reference(names, first_vertex):
    """test/main.cpp:2515-2653: one triangle, its three vertices overwritten and the geometry rebuilt (Rebuild) or refitted
    (Update): both must show the triangle at its new place. The oracle has no update operation (a refit of a one-leaf tree is
    the rebuild); this pins the pictures the GPU tests hold hiprtBuildOperationUpdate to (tests/test_gpu_golden.py)."""
    g = orc.Geometry(REBUILD_VERTICES[first_vertex:first_vertex + 3], np.array([[0, 1, 2]], np.uint32))
    hit = g.trace(H.ortho_rays(), tie_mode=0)["primID"] != orc.INVALID
    for name in names:
        assert _check(mesh_intersection_image(hit), name) < 0.001


def test_shear_matches_reference_image():
    # test/main.cpp:1902-1993 + SceneIntersectionSingleton kernel: an indexed quad (two triangles: too few to pair) instanced with a
    # hiprtFrameMatrix that shears x by y: exercises the matrix -> scale / shear / rotation decomposition of the scene build
    off = 0.25
    v = np.array([[-off, -off, 0], [-off, off, 0], [off, off, 0], [off, -off, 0]], np.float32)
    g = orc.Geometry(v, np.array([[0, 1, 2], [0, 2, 3]], np.uint32))
    m = np.zeros((1, 16), np.float32)
    m[0, [0, 5, 10]] = 1.0
    m[0, 1] = 1.0  # matrix[0][1]
    scene = orc.Scene([g], [0], m, 1)
    w = h = 512
    x = (np.arange(w, dtype=np.float32) / np.float32(w))[None, :].repeat(h, 0)
    y = (np.arange(h, dtype=np.float32) / np.float32(h))[:, None].repeat(w, 1)
    o = np.stack([x - np.float32(0.5), y - np.float32(0.5), np.full_like(x, -1.0)], -1).reshape(-1, 3)
    rays = orc.make_rays(o, np.broadcast_to(np.array([0, 0, 1], np.float32), o.shape), max_t=1000.0)
    hit = scene.trace(rays)["primID"] != orc.INVALID
    img = np.zeros((w * h, 4), np.uint8)
    img[:, 0] = np.where(hit, (x.reshape(-1) * np.float32(255)).astype(np.uint8), 0)
    img[:, 1] = np.where(hit, (y.reshape(-1) * np.float32(255)).astype(np.uint8), 0)
    img[:, 3] = 255
    _check(img, "Shear")


def test_cutout_matches_reference_image():
    """test/main.cpp:1219-1290 + CutoutKernel / cutoutFilter (HiprtTestKernel.h:88-100,216-235): a unit quad of two triangles,
    closest hit with a hit filter that cuts a 16x16 checkerboard out of texture coordinates interpolated from hit.uv per
    primID. The filter only lets a hit through when uv follows the CALLER's vertex order -- this pins the barycentric convention
    (and primID assignment) the oracle reports, including for the second triangle of a quad."""
    v = np.array([[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0]], np.float32)
    g = orc.Geometry(v, np.array([[0, 1, 2], [0, 2, 3]], np.uint32))
    hits = g.trace(H.ortho_rays(), tie_mode=0)
    hit = hits["primID"] != orc.INVALID
    u, w_ = hits["uv"][:, 0], hits["uv"][:, 1]
    # texCoord[primID] = (1-u-v) * t0 + u * t1 + v * t2 with (t0,t1,t2) = ((0,0),(0,1),(1,1)) for triangle 0, ((0,0),(1,1),(1,0)) for 1
    tx = np.where(hits["primID"] == 0, w_, u + w_).astype(np.float32)
    ty = np.where(hits["primID"] == 0, u + w_, u).astype(np.float32)
    rejected = ((np.float32(16.0) * tx).astype(np.uint32) + (np.float32(16.0) * ty).astype(np.uint32)) & 1
    visible = hit & (rejected == 0)
    img = np.zeros((512 * 512, 4), np.uint8)
    img[:, :3] = np.where(visible, 255, 0)[:, None]
    img[:, 3] = 255
    _check(img, "Cutout")
