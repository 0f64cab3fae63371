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


def _moved(verts, step):
    moved = verts.copy()
    moved[:, 0] *= 1.0 + 0.3 * (step + 1)
    moved[:, 1] += 0.2 * np.sin(5.0 * verts[:, 2] + step).astype(np.float32)
    return moved


@pytest.mark.parametrize("how", ["compacted", "saved+loaded"])
def test_update_after_compaction_or_reload_uses_the_header_pointers(ctx, how, tmp_path):
    """A compacted / reloaded structure keeps its node arrays right behind the USED box nodes, not where a fresh build of the same
    input would put them (64 + maxBoxNodeCount * 128): the refit must take the arrays from the device header like the reference
    (LbvhBuilder.h:445-452), else it reads garbage leaves and writes outside the compacted allocation."""
    t = ctx.torch
    verts, idx = scenes.sphere_mesh(96, 48)
    d_verts, d_idx = to_dev(ctx, verts), to_dev(ctx, idx)
    geom = ctx.build(d_verts, d_idx)
    before = geom.header().copy()
    geom.h = ctx.compact_geometry(geom.h)
    if how == "saved+loaded":  # the file keeps the compacted layout
        path = str(tmp_path / "geom.bin")
        ctx.save_geometry(geom.h, path)
        ctx.destroy_geometry(geom.h)
        geom.h = ctx.load_geometry(path)
    hdr = geom.header()
    assert int(hdr["primNodes"]) - int(hdr["boxNodes"]) < int(before["primNodes"]) - int(before["boxNodes"]), "nothing was compacted"
    # a guard allocation right after: an out-of-bounds refit would scribble over it
    guard = t.full((1 << 20,), 0x5A, dtype=t.uint8, device=ctx.device)
    boxes0, leaves0 = geom.nodes()
    for step in range(2):
        moved = _moved(verts, step)
        d_verts.copy_(t.from_numpy(moved))
        ctx.build_geometry(geom.inp, geom.h, geom.temp, api.FAST_BUILD, op=api.UPDATE)
        boxes, leaves = geom.nodes()
        H.validate_wide_bvh(boxes, leaves, len(idx))
        assert np.array_equal(boxes["child"], boxes0["child"]) and (boxes["updateCounter"] == 0).all()
        rays = scenes.random_rays(20000, moved.min(0) - 0.1, moved.max(0) + 0.1, seed=step)
        got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
        compare_hits(got, orc.Geometry(moved, idx).trace(rays, tie_mode=1))
    assert bool((guard == 0x5A).all())
    geom.destroy()


def test_update_of_an_imported_geometry(ctx):
    """Custom-BVH import sizes the storage from the node list; an update of it must find the arrays through the header too."""
    from tests.test_gpu_import import IMPORT, binary_tree_as_api_nodes

    t = ctx.torch
    n_tris = 5000
    verts = scenes.random_soup(n_tris, seed=21, extent=1.0, size=0.05)
    d_verts = to_dev(ctx, verts)
    built = ctx.build(d_verts)
    internal, leaves = binary_tree_as_api_nodes(built, n_tris)
    d_int = t.from_numpy(internal.view(np.uint8).reshape(-1)).to(ctx.device)
    d_leaf = t.from_numpy(leaves.view(np.uint8).reshape(-1)).to(ctx.device)
    inp = ctx.mesh_input(d_verts)
    inp.nodeList.internalNodes, inp.nodeList.leafNodes, inp.nodeList.nodeCount = d_int.data_ptr(), d_leaf.data_ptr(), n_tris
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp, IMPORT))
    handle = ctx.create_geometry(inp, IMPORT)
    ctx.build_geometry(inp, handle, temp, IMPORT)
    g = api.Geometry(ctx, handle, inp, temp, IMPORT)
    moved = (verts * np.float32(1.25) + np.float32(0.1)).astype(np.float32)
    d_verts.copy_(t.from_numpy(moved))
    ctx.build_geometry(inp, handle, temp, IMPORT, op=api.UPDATE)
    boxes, leaf_nodes = g.nodes()
    H.validate_wide_bvh(boxes, leaf_nodes, n_tris)
    rays = scenes.random_rays(20000, moved.min(0), moved.max(0), seed=4)
    got = ctx.hits_to_host(ctx.trace(g, ctx.rays_to_device(rays)))
    compare_hits(got, orc.Geometry(moved, None).trace(rays, tie_mode=1))
    g.destroy()
    built.destroy()


def test_scene_update_after_compaction(ctx):
    meshes = scenes.instanced_geometries(3)
    n_inst = 40
    geom_of = np.arange(n_inst) % 3
    frames = scenes.random_matrix_frames(n_inst, seed=11, extent=6.0)
    geoms = [ctx.build(to_dev(ctx, v), to_dev(ctx, i)) for v, i in meshes]
    ogeoms = [orc.Geometry(v, i) for v, i in meshes]
    scene = ctx.build_scene([geoms[g] for g in geom_of], frames)
    before = scene.header().copy()
    scene.h = ctx.compact_scene(scene.h)
    hdr = scene.header()
    assert int(hdr["primNodes"]) - int(hdr["boxNodes"]) < int(before["primNodes"]) - int(before["boxNodes"])
    guard = ctx.torch.full((1 << 20,), 0x5A, dtype=ctx.torch.uint8, device=ctx.device)
    new_frames = scenes.random_matrix_frames(n_inst, seed=12, extent=6.0)
    new_frames[:, [3, 7, 11]] = 0.7 * frames[:, [3, 7, 11]] + 0.3 * new_frames[:, [3, 7, 11]]
    scene.keep[1].copy_(ctx.torch.from_numpy(new_frames))
    ctx.rebuild_scene(scene, op=api.UPDATE)
    boxes, leaves = scene.nodes()
    assert (boxes["updateCounter"] == 0).all()
    oscene = orc.Scene(ogeoms, geom_of, new_frames, 1)
    rays = scenes.pinhole_rays(128, 128, (0.0, 2.0, 17.0), (0, 0, 0))
    got = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays)))
    pin_transforms(scene, oscene)
    ref = oscene.trace(rays)
    assert (ref["primID"] != orc.INVALID).mean() > 0.03
    compare_hits(got, ref, rel=1e-5)
    assert bool((guard == 0x5A).all())
    scene.destroy()
