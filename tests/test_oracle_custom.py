"""CPU-only: the oracle's general traversal -- custom primitives through an intersection function, scenes instanced in scenes,
the any-hit iterator drained to completion -- against the reference's CustomIntersection.png, SceneIntersection.png and
SceneIntersectionMlas.png (test/main.cpp:1296-1385, 1485-1670, 1672-1900; kernels HiprtTestKernel.h:62-79, 238-262, 286-326), and
against the specialised two-level / geometry traversals it generalises."""
import numpy as np

from oracle import binding as orc
from tests import helpers as H
from tests import test_oracle_golden as G

CIRCLE_BOXES = np.array([[0.15, 0.40, 0.35, 0.60], [0.40, 0.40, 0.60, 0.60], [0.65, 0.40, 0.85, 0.60]], np.float32)
CENTRES = np.array([0.25, 0.5, 0.75], np.float32)
TRIANGLES = np.array([[0.15, 0.40, 0], [0.35, 0.40, 0], [0.25, 0.60, 0], [0.40, 0.40, 0], [0.60, 0.40, 0], [0.50, 0.60, 0],
                      [0.65, 0.40, 0], [0.85, 0.40, 0], [0.75, 0.60, 0]], np.float32)


def circles(z):
    lo = np.concatenate([CIRCLE_BOXES[:, :2], np.full((3, 1), z, np.float32)], 1)
    hi = np.concatenate([CIRCLE_BOXES[:, 2:], np.full((3, 1), z, np.float32)], 1)
    return orc.AabbGeometry(np.concatenate([lo, hi], 1), geom_type=0)


def intersect_circle(ray, geom_type, prim):
    """intersectCircle (HiprtTestKernel.h:62-79) in float32: a disc of radius 0.1 around (centre[prim], 0.5), seen along +z."""
    cx = CENTRES[prim] - ray["origin"][0]
    cy = np.float32(0.5) - ray["origin"][1]
    if np.sqrt(cx * cx + cy * cy, dtype=np.float32) < np.float32(0.1):
        return (np.float32(1.0) - ray["origin"][2]) / ray["direction"][2]
    return None


def unit_square_rays(w=512, h=512):
    x = (np.arange(w, dtype=np.float32) / np.float32(w))[None, :].repeat(h, 0)
    y = (np.arange(h, dtype=np.float32) / np.float32(h))[:, None].repeat(w, 1)
    o = np.stack([x, y, np.full_like(x, -1.0)], -1).reshape(-1, 3)
    return orc.make_rays(o, np.broadcast_to(np.array([0, 0, 1], np.float32), o.shape), max_t=1000.0)


def scene_image(counts, hits):
    """SceneIntersectionKernel: every hit adds the colour of (innermost instance, primitive) in the image's uint8 arithmetic."""
    colours = np.array([[[1, 0, 0], [0, 1, 0], [0, 0, 1]], [[0, 1, 1], [1, 0, 1], [1, 1, 0]]], np.uint8) * np.uint8(255)
    img = np.zeros((len(counts), 4), np.uint8)
    img[:, 3] = 255
    for k in range(hits.shape[1]):
        valid = k < counts
        ids = hits["instanceIDs"][:, k]
        inner = np.where(ids[:, 1] != orc.INVALID, ids[:, 1], ids[:, 0])
        c = colours[np.where(valid, inner, 0), np.where(valid, hits["primID"][:, k], 0)]
        img[:, :3] += np.where(valid[:, None], c, 0).astype(np.uint8)
    return img


def test_custom_intersection_matches_reference_image():
    hits = orc.trace_general(circles(0.0), unit_square_rays(), intersect_circle)
    prim = hits["primID"]
    colours = np.array([[255, 0], [0, 255], [255, 255], [0, 0]], np.uint8)
    img = np.zeros((len(prim), 4), np.uint8)
    img[:, :2] = colours[np.where(prim == orc.INVALID, 3, prim)]
    img[:, 3] = 255
    assert G._check(img, "CustomIntersection") == 0.0


def test_scene_intersection_and_multi_level_match_reference_images():
    srt = lambda dy: [0, 0, 1, 0, 1, 1, 1, 0, dy, 0, 0, 0]
    tris = orc.Geometry(TRIANGLES, np.arange(9, dtype=np.uint32).reshape(3, 3))
    mid = orc.Scene([circles(1.0), tris], [0, 1], np.array([srt(0.05), srt(-0.05)], np.float32), 0)
    rays = unit_square_rays()
    counts, hits = orc.trace_general(mid, rays, intersect_circle, max_per_ray=4)
    assert counts.max() <= 2 and (hits["instanceIDs"][counts > 0, 0, 1] == orc.INVALID).all()
    assert G._check(scene_image(counts, hits), "SceneIntersection") == 0.0

    top = orc.Scene([mid], [0, 0], np.array([srt(0.12), srt(-0.12)], np.float32), 0)
    counts, hits = orc.trace_general(top, rays, intersect_circle, max_per_ray=8)
    hit = counts > 0
    assert counts.max() <= 8 and (hits["instanceIDs"][hit, 0, 0] <= 1).all() and (hits["instanceIDs"][hit, 0, 1] <= 1).all()
    assert G._check(scene_image(counts, hits), "SceneIntersectionMlas") == 0.0


def same_records(a, b):
    return all(np.array_equal(a[f], b[f]) for f in ("instanceIDs", "primID", "uv", "normal", "t"))  # (not the padding word)


def test_general_traversal_agrees_with_the_specialised_ones():
    """Triangles only: the general walker must return the records of orc_scene_trace / orc_geometry_trace, and a scene wrapped
    in an identity top level the same hits one level deeper."""
    scene, geoms = H.cornell_obj_scene([0, 1, 0, 0.5, 1, 1, 1, 0, 0, -3, 0, 0])
    rays = H.obj_camera_rays(128, 128)
    for tie_mode in (0, 1):
        a = scene.trace(rays, tie_mode=tie_mode)
        b = orc.trace_general(scene, rays, tie_mode=tie_mode)
        assert same_records(a, b)
        g = geoms[6]
        assert same_records(g.trace(rays, tie_mode=tie_mode), orc.trace_general(g, rays, tie_mode=tie_mode))
    top = orc.Scene([scene], [0], None, 0)
    c = orc.trace_general(top, rays)
    a = scene.trace(rays)
    hit = a["primID"] != orc.INVALID
    assert np.array_equal(a["primID"], c["primID"]) and np.array_equal(a["t"][hit], c["t"][hit])
    assert (c["instanceIDs"][hit, 0] == 0).all() and np.array_equal(c["instanceIDs"][hit, 1], a["instanceIDs"][hit, 0])
    # every hit along the ray: the same set the geometry-level any-hit iterator enumerates
    verts, _, _ = H.cornell_box()
    box = orc.Geometry(verts, np.arange(96, dtype=np.uint32).reshape(-1, 3))
    box_rays = H.cornell_rays(128, 128)
    counts, prims, _ = box.trace_all(box_rays, max_per_ray=16)
    counts2, hits2 = orc.trace_general(box, box_rays, max_per_ray=16)
    assert np.array_equal(counts, counts2) and np.array_equal(np.sort(prims, 1), np.sort(hits2["primID"], 1))


def test_three_level_scene_equals_its_flattened_scene():
    """Scenes instanced in scenes, three levels deep, against the single-level scene of the composed transforms: the same rays
    hit, the same primitive, the instance ids compose, t agrees to float accuracy (the composed matrix rounds differently)."""
    from hiprt_b200 import scenes

    meshes = scenes.instanced_geometries(3)
    leaf_frames = scenes.random_matrix_frames(6, seed=5, extent=2.0, scale_range=(0.6, 1.2))
    mid_frames = scenes.random_matrix_frames(4, seed=6, extent=6.0, scale_range=(0.8, 1.1))
    top_frames = scenes.random_matrix_frames(3, seed=7, extent=14.0, scale_range=(0.9, 1.0))
    geoms = [orc.Geometry(v, i) for v, i in meshes]
    leaf = orc.Scene(geoms, [k % 3 for k in range(6)], leaf_frames, 1)
    mid = orc.Scene([leaf], [0] * 4, mid_frames, 1)
    top = orc.Scene([mid], [0] * 3, top_frames, 1)

    def matrix(f):
        m = np.eye(4)
        m[:3, :4] = f[:12].reshape(3, 4)
        return m

    flat_frames = np.zeros((72, 16), np.float32)
    for a in range(3):
        for b in range(4):
            for c in range(6):
                flat_frames[a * 24 + b * 6 + c, :12] = (matrix(top_frames[a]) @ matrix(mid_frames[b]) @ matrix(leaf_frames[c]))[:3].reshape(-1)
    flat = orc.Scene(geoms, [k % 6 % 3 for k in range(72)], flat_frames, 1)

    rays = scenes.pinhole_rays(160, 160, (0.0, 4.0, 45.0), (0, 0, 0))
    deep = orc.trace_general(top, rays, tie_mode=1)
    ref = flat.trace(rays, tie_mode=1)
    hit = ref["primID"] != orc.INVALID
    assert hit.mean() > 0.03 and np.array_equal(deep["primID"] != orc.INVALID, hit)
    ids = deep["instanceIDs"][hit]
    assert (ids[:, 3] == orc.INVALID).all()
    same = (deep["primID"][hit] == ref["primID"][hit]) & (ids[:, 0] * 24 + ids[:, 1] * 6 + ids[:, 2] == ref["instanceIDs"][hit, 0])
    assert same.mean() > 0.999
    np.testing.assert_allclose(deep["t"][hit][same], ref["t"][hit][same], rtol=2e-5)


def test_cutout_through_the_hit_filter_matches_reference_image():
    """The Cutout picture again (tests/test_oracle_golden.py evaluates the checkerboard after the fact), this time with the filter
    where the reference has it: inside the traversal, as a hit-filter function of a geometry with a user geometry type
    (cutoutFilter, HiprtTestKernel.h:88-100). A second quad behind the first shows through the holes only because a rejected hit
    does not shorten the ray."""
    v = np.array([[0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0]], np.float32)
    tri = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    calls = [0]

    def cutout(ray, geom_type, hit):
        calls[0] += 1
        u, w_ = hit["uv"]
        tx, ty = (w_, u + w_) if hit["primID"] == 0 else (u + w_, u)
        return bool((int(np.float32(16.0) * np.float32(tx)) + int(np.float32(16.0) * np.float32(ty))) & 1)

    g = orc.Geometry(v, tri)
    g.set_geom_type(0)
    rays = H.ortho_rays(256, 256)
    hits = orc.trace_general(g, rays, hit_filter=cutout, tie_mode=0)
    visible = hits["primID"] != orc.INVALID
    assert calls[0] >= (256 * 256) * 0.9
    img = np.zeros((256 * 256, 4), np.uint8)
    img[:, :3] = np.where(visible, 255, 0)[:, None]
    img[:, 3] = 255
    golden = H.load_golden_image("Cutout")[::2, ::2].reshape(-1, 4)  # the checkerboard cells are 32 pixels wide
    assert (np.abs(img.astype(int) - golden.astype(int)).max(1) > 10).mean() < 0.003

    # two layers: the far quad is reported exactly where the near one is cut out
    v2 = np.concatenate([v, v + np.array([0, 0, 0.5], np.float32)])
    g2 = orc.Geometry(v2, np.concatenate([tri, tri + 4]), allow_pairing=False)
    g2.set_geom_type(0)
    near_only = lambda ray, geom_type, hit: hit["primID"] < 2 and cutout(ray, geom_type, hit)
    both = orc.trace_general(g2, rays, hit_filter=near_only, tie_mode=1)
    far = (both["primID"] >= 2) & (both["primID"] != orc.INVALID)
    assert np.array_equal(far, ~visible & (both["primID"] != orc.INVALID))
    assert (both["primID"][visible] < 2).all() and (both["primID"] != orc.INVALID).mean() > 0.95
