"""CPU-only checks of the oracle's auxiliary entry points (test infrastructure of the GPU parity tests)."""
import numpy as np

from hiprt_b200 import scenes
from oracle import binding as orc


def _scene(n_inst=12):
    meshes = scenes.instanced_geometries(3)
    ogeoms = [orc.Geometry(v, i) for v, i in meshes]
    frames = scenes.random_matrix_frames(n_inst, seed=7, extent=4.0)
    return orc.Scene(ogeoms, np.arange(n_inst) % 3, frames, 1), frames


def test_adopting_its_own_matrices_changes_nothing():
    scene, _ = _scene()
    rays = scenes.pinhole_rays(64, 64, (0.0, 1.0, 11.0), (0, 0, 0))
    before = scene.trace(rays)
    assert (before["primID"] != orc.INVALID).mean() > 0.02
    scene.adopt_transforms(scene.leaves)
    after = scene.trace(rays)
    for field in ("primID", "instanceIDs", "t", "uv", "normal"):  # (the padding word of hiprtHit is not initialised)
        assert np.array_equal(before[field], after[field]), field


def test_adopted_matrices_are_used():
    scene, _ = _scene()
    rays = scenes.pinhole_rays(64, 64, (0.0, 1.0, 11.0), (0, 0, 0))
    before = scene.trace(rays)
    leaves = scene.leaves.copy()
    leaves["invMatrix"][:, :, 3] += 100.0  # move every instance far away in object space
    scene.adopt_transforms(leaves)
    after = scene.trace(rays)
    assert (after["primID"] == orc.INVALID).all() and (before["primID"] != orc.INVALID).any()


def test_per_ray_visit_counts_add_up():
    verts, idx = scenes.sphere_mesh(64, 32)
    g = orc.Geometry(verts, idx)
    rays = scenes.pinhole_rays(48, 48, (0.0, 0.0, 3.0), (0, 0, 0))
    hits, totals = g.trace(rays, want_visits=True)
    bv, lv = g.trace_visits(rays)
    assert int(bv.sum()) == int(totals[0]) and int(lv.sum()) == int(totals[1])
    assert (bv >= 1).all()  # every ray visits at least the root
    assert ((lv > 0) | (hits["primID"] == orc.INVALID)).all()  # a hit needs a leaf visit


def test_box_unions_follow_cuda_min_max_on_signed_zeros():
    """CUDA's fminf(+0, -0) is -0 and fmaxf(+0, -0) is +0 whatever the argument order; the host libm returns the second argument
    (glibc 2.39), which once made the oracle's leaf child boxes differ from the GPU's in the sign bit of a zero
    (pole triangles of sphere_mesh(32, 16) without pairing). Every box union of the oracle must be order-independent."""
    pz, nz = np.float32(0.0), np.float32(-0.0)
    tris = []
    for zs in ((pz, nz, pz), (nz, pz, nz), (pz, pz, nz), (nz, nz, pz), (nz, pz, pz), (pz, nz, nz)):
        x = np.float32(len(tris))
        tris.append([[x, 0.0, zs[0]], [x + 0.5, 0.0, zs[1]], [x, 0.5, zs[2]]])
    verts = np.array(tris, np.float32).reshape(-1, 3)
    g = orc.Geometry(verts, None, allow_pairing=False)
    boxes = g.prim_boxes  # [P, 2, 3]
    assert np.signbit(boxes[:, 0, 2]).all(), "min z of {+0, -0} must be -0"
    assert not np.signbit(boxes[:, 1, 2]).any(), "max z of {+0, -0} must be +0"
    wide = g.box_nodes
    used = wide["child"] != orc.INVALID
    assert np.signbit(wide["box"][:, :, 0, 2][used]).all() and not np.signbit(wide["box"][:, :, 1, 2][used]).any()
    sb = g.scene_box
    assert np.signbit(sb[0, 2]) and not np.signbit(sb[1, 2])
    # the same triangles with their vertices rotated give bit-identical boxes
    g2 = orc.Geometry(verts.reshape(-1, 3, 3)[:, [1, 2, 0]].reshape(-1, 3), None, allow_pairing=False)
    assert g2.prim_boxes.tobytes() == boxes.tobytes()


def test_sphere_poles_unpaired_have_cuda_signed_zero_boxes():
    """The case the round-1 GPU run failed on: nodes 465/2, 485/0, 486/2 of sphere_mesh(32, 16) without pairing."""
    verts, idx = scenes.sphere_mesh(32, 16)
    g = orc.Geometry(verts, idx, allow_pairing=False)
    tri = verts[idx.astype(np.int64)]  # [T, 3, 3]
    lo, hi = g.prim_boxes[:, 0], g.prim_boxes[:, 1]
    zero_lo, zero_hi = lo == 0.0, hi == 0.0
    has_neg = (np.signbit(tri) & (tri == 0.0)).any(axis=1)  # [T, 3]: some vertex holds -0 in that component
    has_pos = (~np.signbit(tri) & (tri == 0.0)).any(axis=1)
    assert (np.signbit(lo)[zero_lo & has_neg]).all()
    assert (~np.signbit(hi)[zero_hi & has_pos]).all()
    assert (zero_lo & has_neg & has_pos).any(), "the mesh no longer exercises the {+0, -0} case"
