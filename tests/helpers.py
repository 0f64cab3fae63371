"""Shared test helpers: ray generators restating the reference's test kernels, image diff."""
import os

import numpy as np

from oracle import binding as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32 = np.float32


# golden images the reference keeps as byte-identical copies of another one (tests/golden/make_golden.py asserts the equality):
# the same picture validated through another builder, the BVH import or the save / load path
GOLDEN_ALIASES = {"BvhFastCornellBox": "PrimaryRayCornellBox", "BvhBalancedCornellBox": "PrimaryRayCornellBox",
                  "BvhHighQCornellBox": "PrimaryRayCornellBox", "CustomBvhImport": "MinimumCornellBox",
                  "BvhIoApi": "MinimumCornellBox"}


def load_golden_image(name):
    return np.load(os.path.join(GOLDEN, "golden_images.npz"))[GOLDEN_ALIASES.get(name, name)]


def cornell_box():
    d = np.load(os.path.join(GOLDEN, "cornell_box.npz"))
    return d["vertices"], d["mat_indices"], d["diffuse"]


def cornell_rays(width=512, height=512):
    """Ray generator of CornellBoxKernel (reference test/kernels/HiprtTestKernel.h:118-134)."""
    x = np.arange(width, dtype=f32)[None, :].repeat(height, 0)
    y = np.arange(height, dtype=f32)[:, None].repeat(width, 1)
    dx = f32(2.0) * x / f32(width) - f32(1.0)
    dy = f32(2.0) * (f32(1.0) - y / f32(height)) - f32(1.0)
    d = np.stack([f32(-387.817566) * dx, f32(387.817566) * dy, np.full_like(dx, f32(1230.0))], -1)
    ln = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2], dtype=f32)
    d = (d / ln[..., None]).astype(f32)
    o = np.broadcast_to(np.array([278.0, 273.0, -900.0], f32), d.shape)
    return orc.make_rays(o.reshape(-1, 3), d.reshape(-1, 3))


def ortho_rays(width=512, height=512, max_t=orc.FLT_MAX):
    """Rays of MeshIntersectionKernel & friends (HiprtTestKernel.h:160-175): origin (x/w, y/h, -1), +z."""
    x = (np.arange(width, dtype=f32) / f32(width))[None, :].repeat(height, 0)
    y = (np.arange(height, dtype=f32) / f32(height))[:, None].repeat(width, 1)
    o = np.stack([x, y, np.full_like(x, f32(-1.0))], -1).reshape(-1, 3)
    d = np.broadcast_to(np.array([0, 0, 1], f32), o.shape)
    return orc.make_rays(o, d, max_t=max_t)


def image_matches(img, ref, max_bad_frac=0.003, tol=10):
    """The reference's image check (test/hiprtTest.cpp:943-1002): a pixel is bad when any
    channel differs by more than `tol`; pass when < 0.3 % of pixels are bad."""
    diff = np.abs(img.astype(np.int32) - ref.astype(np.int32))
    bad = (diff > tol).any(-1)
    return bad.mean() < max_bad_frac, float(bad.mean()), int(diff.max())


def tea16(v0, v1):
    """tea<16> of the reference's test/shared.h:117-131, vectorised over uint32 arrays."""
    v0 = np.asarray(v0, np.uint32).copy()
    v1 = np.asarray(v1, np.uint32).copy()
    s0 = np.uint32(0)
    with np.errstate(over="ignore"):
        for _ in range(16):
            s0 = np.uint32((int(s0) + 0x9E3779B9) & 0xFFFFFFFF)
            v0 += ((v1 << np.uint32(4)) + np.uint32(0xA341316C)) ^ (v1 + s0) ^ ((v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))
            v1 += ((v0 << np.uint32(4)) + np.uint32(0xAD90777D)) ^ (v0 + s0) ^ ((v0 >> np.uint32(5)) + np.uint32(0x7E95761E))
    return v0, v1


def validate_wide_bvh(boxes, leaves, prim_count_expected=None, duplicated_references=False):
    """Structural check of a 4-wide BVH read back from the device (any builder): every leaf is referenced
    exactly once, every child box equals the exact union of what is below it (boxes are fp32 min/max, so
    equality is exact), parent links are consistent. Returns the SAH-style cost sum(child areas)/root area.
    duplicated_references: a high-quality build with spatial splits -- a primitive may sit in several leaves, each with the
    part of its box the leaf is responsible for, and an inner child box is the union of the BINS' boxes, which contains
    (rather than equals) the union of the boxes below it."""
    INVALID = 0xFFFFFFFF
    child = boxes["child"]
    valid = child != INVALID
    ctype = child & 7
    is_box = valid & (ctype == 5)
    is_leaf = valid & (ctype != 5)
    assert np.array_equal(valid.sum(1), boxes["childCount"]), "childCount does not match the valid slots"
    # leaves: referenced exactly once
    leaf_addr = child[is_leaf] >> np.where(ctype[is_leaf] == 6, 4, 3)  # (instance nodes are addressed like box nodes)
    assert len(leaf_addr) == len(leaves) and np.array_equal(np.sort(leaf_addr), np.arange(len(leaves))), "leaves not referenced exactly once"
    # box children: referenced exactly once, never the root
    box_addr = child[is_box] >> 4
    assert np.array_equal(np.sort(box_addr), np.arange(1, len(boxes))), "box nodes not referenced exactly once"
    # parent links
    parent_of = np.full(len(boxes), INVALID, np.uint32)
    rows = np.nonzero(is_box)[0]
    parent_of[box_addr] = rows
    assert np.array_equal(parent_of, boxes["parentAddr"]), "parent links are wrong"
    # child boxes: exact unions
    lo = np.where(valid[..., None], boxes["box"][:, :, 0, :], np.inf).min(1)
    hi = np.where(valid[..., None], boxes["box"][:, :, 1, :], -np.inf).max(1)
    cb = boxes["box"][is_box]
    if duplicated_references:
        assert np.all(cb[:, 0] <= cb[:, 1]), "empty inner child box"
    else:
        assert np.array_equal(cb[:, 0], lo[box_addr]) and np.array_equal(cb[:, 1], hi[box_addr]), "inner child box is not the union of its children"
    if leaves.dtype.names and "v" in leaves.dtype.names:
        v = leaves["v"]
        lb = boxes["box"][is_leaf]
        if duplicated_references:
            assert np.all(lb[:, 0] >= v[leaf_addr].min(1)) and np.all(lb[:, 1] <= v[leaf_addr].max(1)) and np.all(lb[:, 0] <= lb[:, 1]), "leaf box leaves the box of its vertices"
        else:
            assert np.array_equal(lb[:, 0], v[leaf_addr].min(1)) and np.array_equal(lb[:, 1], v[leaf_addr].max(1)), "leaf box is not the box of its vertices"
        prims = np.concatenate([leaves["primIndex0"], leaves["primIndex1"][leaves["primIndex1"] != leaves["primIndex0"]]])
        if prim_count_expected is not None:
            if duplicated_references:
                assert np.array_equal(np.unique(prims), np.arange(prim_count_expected)), "triangles are not all covered"
            else:
                assert np.array_equal(np.sort(prims), np.arange(prim_count_expected)), "triangles are not covered exactly once"
    ext = np.maximum(boxes["box"][:, :, 1, :] - boxes["box"][:, :, 0, :], 0)
    area = 2 * (ext[..., 0] * ext[..., 1] + ext[..., 0] * ext[..., 2] + ext[..., 1] * ext[..., 2])
    root_ext = hi[0] - lo[0]
    root_area = 2 * (root_ext[0] * root_ext[1] + root_ext[0] * root_ext[2] + root_ext[1] * root_ext[2])
    return float(np.where(valid, area, 0).sum() / max(root_area, 1e-30))


def cornell_obj_scene(srt_frame=None):
    """The cornellBox.obj scene of the reference's ObjTestCases (hiprtTest.cpp:1026-1300): one geometry per OBJ shape, one
    instance per geometry, all with the identity frame or the same hiprtFrameSRT (12 floats: axis-angle rotation, scale,
    translation, time, pad); instanceID = shape index, primID = face index inside the shape."""
    d = np.load(os.path.join(GOLDEN, "cornell_obj.npz"))
    verts = d["vertices"]
    n = int(d["shape_count"])
    geoms = [orc.Geometry(verts, d[f"tris_{k}"]) for k in range(n)]
    frames = None if srt_frame is None else np.tile(np.asarray(srt_frame, np.float32).reshape(1, 12), (n, 1))
    scene = orc.Scene(geoms, np.arange(n), frames, 0)
    return scene, geoms


def cornell_obj_face_colours():
    """Diffuse colour (Kd of cornellBox.mtl) per (shape, face): list of float32 [faces, 3]."""
    d = np.load(os.path.join(GOLDEN, "cornell_obj.npz"))
    return [d[f"diffuse_{k}"] for k in range(int(d["shape_count"]))]


def obj_camera_rays(width=512, height=512, translation=(0.0, 2.5, 5.8), fov=45.0 * np.pi / 180.0):
    """generateRay of the reference's test/shared.h:151-168 with the TestCornellBox camera (hiprtTest.h:256-261: rotation
    angle 0, so holDir = +x, upDir = -y, viewDir = -z), pixel offset 0.5."""
    x = np.arange(width, dtype=f32)[None, :].repeat(height, 0)
    y = np.arange(height, dtype=f32)[:, None].repeat(width, 1)
    sx, sy = f32(0.024) * f32(width / height), f32(0.024)
    px = (x + f32(0.5)) / f32(width) - f32(0.5)
    py = (y + f32(0.5)) / f32(height) - f32(0.5)
    dz = f32(sy / f32(2.0 * np.tan(f32(fov) / f32(2.0))))
    d = np.stack([px * sx, -(py * sy), np.full_like(px, -dz)], -1).astype(f32)
    ln = np.sqrt((d * d).sum(-1, dtype=f32), dtype=f32)
    d = (d / ln[..., None]).astype(f32)
    o = np.broadcast_to(np.asarray(translation, f32), d.shape)
    return orc.make_rays(o.reshape(-1, 3), d.reshape(-1, 3))
