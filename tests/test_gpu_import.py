"""Custom BVH import (hiprtBuildFlagBitCustomBvhImport; reference: BvhImporter.h:150-330, node types hiprt_types.h:385-435):
a caller-provided binary tree (internal nodes with own box + two typed child indices, leaves with box + primID, root first) is
collapsed to wide nodes by the same kernel as every other build. The binary tree fed in here is the LBVH this library built
for the same mesh (read back from the temporary buffer, root moved to position 0), so the imported structure must equal the
directly built one after canonicalisation, and hits must equal the oracle's."""
import numpy as np
import pytest

from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests import helpers as H
from tests.test_gpu_parity import compare_hits, to_dev

pytestmark = pytest.mark.gpu

INTERNAL_DTYPE = np.dtype([("aabbMin", "<f4", 3), ("aabbMax", "<f4", 3), ("childIndices", "<u4", 2), ("childNodeTypes", "<u4", 2), ("_pad", "<u4", 6)])
LEAF_DTYPE = np.dtype([("aabbMin", "<f4", 3), ("aabbMax", "<f4", 3), ("primID", "<u4"), ("_pad", "<u4")])
IMPORT = 3  # hiprtBuildFlagBitCustomBvhImport


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def binary_tree_as_api_nodes(geom, n):
    """(internal nodes, leaf nodes) of the last LBVH build of `geom` in the import format, root at position 0."""
    scratch = geom.temp_array("scratch", api.SCRATCH_DTYPE, n - 1).copy()
    refs = geom.temp_array("sortedRefs", api.REF_DTYPE, n).copy()
    return api_nodes_from_records(scratch, refs, n)


def api_nodes_from_records(scratch, refs, n):
    is_box = lambda c: (c & 7) == 5
    addr = lambda c: np.where(is_box(c) | ((c & 7) == 6), c >> 4, c >> 3)  # box and instance indices shift by 4 (BvhNode.h:88-103)
    children = np.stack([scratch["child0"], scratch["child1"]], 1)
    child_is_box = is_box(children)
    child_addr = addr(children)
    referenced = np.zeros(n - 1, bool)
    referenced[child_addr[child_is_box]] = True
    root = int(np.nonzero(~referenced)[0][0])
    # swap internal nodes `root` and 0
    perm = np.arange(n - 1)
    perm[0], perm[root] = root, 0  # new position k holds old node perm[k]
    new_of_old = np.empty(n - 1, np.int64)
    new_of_old[perm] = np.arange(n - 1)
    internal = np.zeros(n - 1, INTERNAL_DTYPE)
    internal["aabbMin"] = scratch["box"][perm, 0]
    internal["aabbMax"] = scratch["box"][perm, 1]
    ca, cb = child_addr[perm], child_is_box[perm]
    internal["childIndices"] = np.where(cb, new_of_old[np.minimum(ca, n - 2)], ca)
    internal["childNodeTypes"] = np.where(cb, 0, 1)  # hiprtBvhNodeTypeInternal = 0, Leaf = 1
    leaves = np.zeros(n, LEAF_DTYPE)
    leaves["aabbMin"], leaves["aabbMax"], leaves["primID"] = refs["box"][:, 0], refs["box"][:, 1], refs["primIndex"]
    return internal, leaves


@pytest.mark.parametrize("n_tris", [2, 3, 64, 5000, 200_000])
def test_import_of_own_lbvh_reproduces_the_built_structure(ctx, n_tris):
    verts = scenes.random_soup(n_tris, seed=n_tris, extent=1.0, size=0.05)  # non-indexed: one triangle per leaf
    d_verts = to_dev(ctx, verts)
    built = ctx.build(d_verts)
    internal, leaves = binary_tree_as_api_nodes(built, n_tris)
    d_int = ctx.torch.from_numpy(internal.view(np.uint8).reshape(-1)).to(ctx.device)
    d_leaf = ctx.torch.from_numpy(leaves.view(np.uint8).reshape(-1)).to(ctx.device)
    inp = ctx.mesh_input(d_verts)
    inp.nodeList.internalNodes = d_int.data_ptr()
    inp.nodeList.leafNodes = d_leaf.data_ptr()
    inp.nodeList.nodeCount = n_tris
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp, IMPORT))
    handle = ctx.create_geometry(inp, IMPORT)
    ctx.build_geometry(inp, handle, temp, IMPORT)
    imported = api.Geometry(ctx, handle, inp, temp, IMPORT)
    assert imported.build_error_flag() == 0
    hdr = imported.header()
    assert int(hdr["primNodeCount"]) == n_tris and int(hdr["boxNodeCount"]) == int(built.header()["boxNodeCount"])
    boxes, leaf_nodes = imported.nodes()
    H.validate_wide_bvh(boxes, leaf_nodes, n_tris)
    ib, il = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaf_nodes.view(orc.TRIPAIR_DTYPE))
    bb, bl = orc.canonicalize(*[a.view(d) for a, d in zip(built.nodes(), (orc.BOX4_DTYPE, orc.TRIPAIR_DTYPE))])
    assert ib.tobytes() == bb.tobytes() and il.tobytes() == bl.tobytes(), "imported structure differs from the directly built one"
    rays = scenes.random_rays(20_000, verts.min(0), verts.max(0), seed=3)
    got = ctx.hits_to_host(ctx.trace(imported, ctx.rays_to_device(rays)))
    compare_hits(got, orc.Geometry(verts, None).trace(rays, tie_mode=1))
    imported.destroy()
    built.destroy()


def test_import_with_a_primitive_referenced_by_two_leaves(ctx):
    """Spatial-split style input: more leaves than primitives (hiprt_types.h: 'Multiple primitives can be referenced from
    different leaves'); storage is sized from the node list."""
    verts = np.array([[0, 0, 0], [4, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1]], np.float32)  # long triangle 0, small triangle 1
    internal = np.zeros(2, INTERNAL_DTYPE)
    leaves = np.zeros(3, LEAF_DTYPE)
    # triangle 0 split at x = 2 into two leaf boxes, triangle 1 in its own leaf
    leaves["aabbMin"] = [[0, 0, 0], [2, 0, 0], [0, 0, 1]]
    leaves["aabbMax"] = [[2, 1, 0], [4, 0.5, 0], [1, 1, 1]]
    leaves["primID"] = [0, 0, 1]
    internal["aabbMin"], internal["aabbMax"] = [[0, 0, 0], [0, 0, 0]], [[4, 1, 1], [4, 1, 0]]
    internal["childIndices"] = [[1, 2], [0, 1]]
    internal["childNodeTypes"] = [[0, 1], [1, 1]]
    d_verts = to_dev(ctx, verts)
    d_int = ctx.torch.from_numpy(internal.view(np.uint8).reshape(-1)).to(ctx.device)
    d_leaf = ctx.torch.from_numpy(leaves.view(np.uint8).reshape(-1)).to(ctx.device)
    inp = ctx.mesh_input(d_verts)
    inp.nodeList.internalNodes, inp.nodeList.leafNodes, inp.nodeList.nodeCount = d_int.data_ptr(), d_leaf.data_ptr(), 3
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp, IMPORT))
    handle = ctx.create_geometry(inp, IMPORT)
    ctx.build_geometry(inp, handle, temp, IMPORT)
    g = api.Geometry(ctx, handle, inp, temp, IMPORT)
    assert int(g.header()["primNodeCount"]) == 3
    o = np.array([[0.5, 0.2, -1], [3.0, 0.1, -1], [0.2, 0.2, 2], [3.9, 0.9, -1]], np.float32)
    d = np.array([[0, 0, 1], [0, 0, 1], [0, 0, -1], [0, 0, 1]], np.float32)
    hits = ctx.hits_to_host(ctx.trace(g, ctx.rays_to_device(scenes.make_rays(o, d))))
    assert hits["primID"].tolist() == [0, 0, 1, api.INVALID]
    np.testing.assert_allclose(hits["t"][:3], [1.0, 1.0, 1.0], rtol=1e-6)
    g.destroy()


def test_scene_import_of_own_top_level_tree(ctx):
    """hiprtBuildScene with a caller-provided top-level tree (one leaf per instance)."""
    meshes = scenes.instanced_geometries(4)
    n_inst = 60
    geom_of = np.arange(n_inst) % 4
    frames = scenes.random_matrix_frames(n_inst, seed=9, extent=8.0)
    geoms = [ctx.build(to_dev(ctx, v), to_dev(ctx, i)) for v, i in meshes]
    scene = ctx.build_scene([geoms[g] for g in geom_of], frames)
    built_boxes, built_leaves = scene.nodes()
    rays = scenes.pinhole_rays(128, 128, (0.0, 3.0, 24.0), (0, 0, 0))
    want = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays)))
    assert (want["primID"] != api.INVALID).mean() > 0.05
    L = api.temp_layout(n_inst, 0)
    ctx.torch.cuda.synchronize()
    raw = scene.temp.cpu().numpy()
    scratch = raw[L["scratch"]: L["scratch"] + 32 * (n_inst - 1)].view(api.SCRATCH_DTYPE).copy()
    refs = raw[L["sortedRefs"]: L["sortedRefs"] + 32 * n_inst].view(api.REF_DTYPE).copy()
    internal, leaves = api_nodes_from_records(scratch, refs, n_inst)
    d_int = ctx.torch.from_numpy(internal.view(np.uint8).reshape(-1)).to(ctx.device)
    d_leaf = ctx.torch.from_numpy(leaves.view(np.uint8).reshape(-1)).to(ctx.device)
    scene.inp.nodeList.internalNodes, scene.inp.nodeList.leafNodes, scene.inp.nodeList.nodeCount = d_int.data_ptr(), d_leaf.data_ptr(), n_inst
    scene.flags = IMPORT
    ctx.rebuild_scene(scene)
    boxes, inst_leaves = scene.nodes()
    assert int(scene.header()["primNodeCount"]) == n_inst
    ib, il = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), inst_leaves.view(orc.INSTANCE_NODE_DTYPE))
    bb, bl = orc.canonicalize(built_boxes.view(orc.BOX4_DTYPE), built_leaves.view(orc.INSTANCE_NODE_DTYPE))
    assert ib.tobytes() == bb.tobytes() and il.tobytes() == bl.tobytes()
    got = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays)))
    assert np.array_equal(got["primID"], want["primID"]) and np.array_equal(got["instanceIDs"][:, 0], want["instanceIDs"][:, 0])
    hit = want["primID"] != api.INVALID
    assert np.array_equal(got["t"][hit], want["t"][hit])
    scene.destroy()
    for g in geoms:
        g.destroy()
