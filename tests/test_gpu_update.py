"""GPU tests of hiprtBuildOperationUpdate (refit): reference ResetCountersAndUpdateLeaves + FitBounds
(BvhBuilderKernels.h:550-814), tests Update (test/main.cpp:2585). Topology is kept, so the structure is checked for
validity (exact child-box unions, leaves holding the moved vertices) and the hits against the oracle built from
the moved input."""
import numpy as np
import pytest

from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests import helpers as H
from tests.test_gpu_parity import compare_hits, to_dev
from tests.test_gpu_scene import pin_transforms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("flags", [api.FAST_BUILD, api.BALANCED_BUILD])
@pytest.mark.parametrize("indexed", [True, False])
def test_geometry_update_refits_moved_vertices(ctx, flags, indexed):
    t = ctx.torch
    if indexed:
        verts, idx = scenes.sphere_mesh(120, 60)
    else:
        verts, idx = scenes.random_soup(20000, seed=8), None
    d_verts = to_dev(ctx, verts)
    d_idx = None if idx is None else to_dev(ctx, idx)
    geom = ctx.build(d_verts, d_idx, flags=flags)
    n_tris = len(verts) // 3 if idx is None else len(idx)
    boxes0, leaves0 = geom.nodes()
    for step in range(2):  # update twice: counters must come back to zero
        moved = verts.copy()
        moved[:, 0] *= 1.0 + 0.3 * (step + 1)
        moved[:, 1] += 0.2 * np.sin(5.0 * verts[:, 2] + step).astype(np.float32)
        d_verts.copy_(t.from_numpy(moved))
        ctx.build_geometry(geom.inp, geom.h, geom.temp, flags, op=api.UPDATE)
        boxes, leaves = geom.nodes()
        H.validate_wide_bvh(boxes, leaves, n_tris)
        assert np.array_equal(boxes["child"], boxes0["child"]), "update changed the topology"
        assert np.array_equal(leaves["primIndex0"], leaves0["primIndex0"]) and np.array_equal(leaves["flags"], leaves0["flags"])
        assert (boxes["updateCounter"] == 0).all()
        og = orc.Geometry(moved, idx)
        rays = scenes.random_rays(20000, moved.min(0) - 0.1, moved.max(0) + 0.1, seed=step)
        got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
        compare_hits(got, og.trace(rays, tie_mode=1))
    geom.destroy()


def test_update_of_singleton_and_batch_built_geometry(ctx):
    t = ctx.torch
    meshes = [scenes.random_soup(k, seed=k, size=0.3) for k in (1, 2, 37, 400)]
    d = [to_dev(ctx, v) for v in meshes]
    inputs = [ctx.mesh_input(x) for x in d]
    geoms = ctx.build_geometries(inputs, batch_max=512)
    for v, dv in zip(meshes, d):
        dv.copy_(t.from_numpy((v * 1.5 + 0.25).astype(np.float32)))
    arr, handles = geoms[0].batch
    import ctypes as C
    ctx._check(api.lib().hiprtBuildGeometries(ctx.h, api.UPDATE, len(geoms), arr, api.BuildOptions(api.FAST_BUILD, 512), api._ptr(geoms[0].temp),
                                              api._stream_handle(None), handles), "hiprtBuildGeometries(update)")
    for v, g in zip(meshes, geoms):
        moved = (v * 1.5 + 0.25).astype(np.float32)
        boxes, leaves = g.nodes()
        H.validate_wide_bvh(boxes, leaves, len(v) // 3)
        assert np.array_equal(np.sort(leaves["v"][:, :3].reshape(-1, 3), 0), np.sort(moved, 0))
    ctx.destroy_geometries(geoms)


def test_scene_update_with_new_frames(ctx):
    meshes = scenes.instanced_geometries(4)
    n_inst = 50
    geom_of = np.arange(n_inst) % 4
    frames = scenes.random_matrix_frames(n_inst, seed=1, extent=6.0)
    geoms = [ctx.build(to_dev(ctx, v), to_dev(ctx, i)) for v, i in meshes]
    ogeoms = [orc.Geometry(v, i) for v, i in meshes]
    scene = ctx.build_scene([geoms[g] for g in geom_of], frames)
    d_frames = scene.keep[1]
    new_frames = scenes.random_matrix_frames(n_inst, seed=2, extent=6.0)
    new_frames[:, [3, 7, 11]] = 0.7 * frames[:, [3, 7, 11]] + 0.3 * new_frames[:, [3, 7, 11]]  # instances drift, do not teleport
    d_frames.copy_(ctx.torch.from_numpy(new_frames))
    ctx.rebuild_scene(scene, op=api.UPDATE)
    boxes, leaves = scene.nodes()
    assert (boxes["updateCounter"] == 0).all()
    oscene = orc.Scene(ogeoms, geom_of, new_frames, 1)
    rays = scenes.pinhole_rays(160, 160, (0.0, 2.0, 17.0), (0, 0, 0))
    got = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays)))
    pin_transforms(scene, oscene)  # same world->object matrices on both sides: strict bar (tests/test_gpu_scene.py)
    ref = oscene.trace(rays)
    assert (ref["primID"] != orc.INVALID).mean() > 0.03
    compare_hits(got, ref, rel=1e-5)
    same = got["primID"] == ref["primID"]
    assert np.array_equal(got["instanceIDs"][same, 0], ref["instanceIDs"][same, 0])
    scene.destroy()
