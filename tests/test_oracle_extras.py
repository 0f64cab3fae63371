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
