"""PreferHighQualityBuild on the GPU (hiprt_b200/csrc/sbvh_kernels.cuh: one cooperative launch builds the binned-SAH tree with
spatial splits) against the oracle's restatement of the same algorithm (oracle/hiprt_oracle.cpp, namespace sbvh; reference
SbvhBuilder.h:174-560, SbvhBuilderKernels.h:65-1077): the canonicalised wide nodes and leaves must be byte-identical, the hits
the ones of every other builder."""
import time

import numpy as np
import pytest

from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests import helpers as H
from tests.test_gpu_parity import compare_hits, to_dev
from tests.test_oracle_sbvh import CASES, long_thin_soup

pytestmark = pytest.mark.gpu

DISABLE_SPATIAL_SPLITS = 1 << 2


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def sbvh_state(geom):
    """sbvh::State of the last build (second 64-byte line of the temporary buffer)."""
    geom.ctx.torch.cuda.synchronize()
    w = geom.temp[64:128].cpu().numpy().view(np.uint32)
    return {"error": int(w[1]), "totalRefs": int(w[2]), "leafCount": int(w[3]), "nodeCount": int(w[4]), "levels": int(w[5])}


@pytest.mark.parametrize("name", list(CASES) + ["needles20k", "cornell"])
@pytest.mark.parametrize("spatial", [True, False])
def test_high_quality_build_equals_the_oracle(ctx, name, spatial):
    if name == "needles20k":
        verts, idx = long_thin_soup(20000)
    elif name == "cornell":
        verts, idx = H.cornell_box()[0], np.arange(96, dtype=np.uint32).reshape(-1, 3)
    else:
        verts, idx = CASES[name]()
    flags = api.HIGH_QUALITY_BUILD | (0 if spatial else DISABLE_SPATIAL_SPLITS)
    geom = ctx.build(to_dev(ctx, verts), None if idx is None else to_dev(ctx, idx), flags=flags)
    og = orc.Geometry(verts, idx, high_quality=True, spatial_splits=spatial)
    levels, refs, spatial_splits, unordered, inconsistent = (int(x) for x in og.sbvh_stats)
    st, hdr = sbvh_state(geom), geom.header()
    assert st["error"] == 0 and geom.build_error_flag() == 0
    assert st["leafCount"] == st["totalRefs"] == int(hdr["primNodeCount"]) and st["nodeCount"] == st["leafCount"] - 1
    boxes, leaves = geom.nodes()
    H.validate_wide_bvh(boxes, leaves, duplicated_references=True)
    if unordered == 0:  # (a split by arrival of more than two references has no unique outcome)
        assert (st["totalRefs"], st["levels"]) == (refs, levels)
        cb, cl = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaves.view(orc.TRIPAIR_DTYPE))
        assert cb.tobytes() == og.box_nodes.tobytes(), "wide nodes differ from the oracle"
        assert cl.tobytes() == og.leaves.tobytes(), "leaf nodes differ from the oracle"
    rays = scenes.random_rays(50000, verts.min(0), verts.max(0), seed=3)
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
    compare_hits(got, og.trace(rays, tie_mode=1))
    geom.destroy()


@pytest.mark.parametrize("name", ["sphere", "needles", "soup"])
def test_per_task_phases_by_single_threads_build_the_same_tree(ctx, name, monkeypatch):
    """Levels with many tasks run the split search with one thread per task instead of a sub-warp (sbvh_kernels.cuh,
    searchBinsSerial); HIPRT_B200_SBVH_SERIAL_MIN = 1 forces that form on every level: still the oracle's tree, byte for byte."""
    monkeypatch.setenv("HIPRT_B200_SBVH_SERIAL_MIN", "1")
    verts, idx = CASES[name]()
    geom = ctx.build(to_dev(ctx, verts), None if idx is None else to_dev(ctx, idx), flags=api.HIGH_QUALITY_BUILD)
    og = orc.Geometry(verts, idx, high_quality=True)
    assert sbvh_state(geom)["error"] == 0 and int(og.sbvh_stats[3]) == 0
    boxes, leaves = geom.nodes()
    cb, cl = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaves.view(orc.TRIPAIR_DTYPE))
    assert cb.tobytes() == og.box_nodes.tobytes() and cl.tobytes() == og.leaves.tobytes()
    geom.destroy()


def test_high_quality_build_of_an_aabb_list_equals_the_oracle(ctx):
    rng = np.random.default_rng(2)
    c = rng.random((3000, 3)).astype(np.float32)
    e = rng.lognormal(-3.5, 1.2, (3000, 3)).astype(np.float32)
    aabbs = np.concatenate([c - e, c + e], 1)
    inp = ctx.aabb_input(to_dev(ctx, aabbs), geom_type=0)
    temp = ctx.alloc_temp(ctx.geometry_temp_size(inp, api.HIGH_QUALITY_BUILD))
    g = ctx.create_geometry(inp, api.HIGH_QUALITY_BUILD)
    ctx.build_geometry(inp, g, temp, api.HIGH_QUALITY_BUILD)
    geom = api.Geometry(ctx, g, inp, temp, api.HIGH_QUALITY_BUILD)
    og = orc.AabbGeometry(aabbs, high_quality=True)
    assert sbvh_state(geom)["error"] == 0
    boxes, leaves = geom.nodes()
    assert len(leaves) > 3000  # references were duplicated
    cb, cl = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaves)
    assert cb.tobytes() == og.box_nodes.tobytes() and cl.tobytes() == og.leaves.tobytes()
    geom.destroy()


def test_compaction_and_update_after_a_high_quality_build(ctx):
    # reference: the TraceKernel unit test compacts a high-quality geometry (test/main.cpp:640-660); an update keeps the
    # topology, the duplicated leaves are rewritten from their primitives and the boxes refitted around them
    t = ctx.torch
    verts, idx = long_thin_soup(5000, seed=4)
    d_verts = to_dev(ctx, verts)
    geom = ctx.build(d_verts, None, flags=api.HIGH_QUALITY_BUILD)
    n_leaves = int(geom.header()["primNodeCount"])
    assert n_leaves > 5000
    rays = scenes.random_rays(30000, verts.min(0), verts.max(0), seed=5)
    d_rays = ctx.rays_to_device(rays)
    ref = orc.Geometry(verts, None).trace(rays, tie_mode=1)
    compare_hits(ctx.hits_to_host(ctx.trace(geom, d_rays)), ref)
    moved = (verts * np.float32(1.25) + np.float32(0.1)).astype(np.float32)
    d_verts.copy_(t.from_numpy(moved))
    ctx.build_geometry(geom.inp, geom.h, geom.temp, api.HIGH_QUALITY_BUILD, op=api.UPDATE)
    boxes, leaves = geom.nodes()
    assert len(leaves) == n_leaves
    H.validate_wide_bvh(boxes, leaves, 5000, duplicated_references=True)
    rays2 = scenes.random_rays(30000, moved.min(0), moved.max(0), seed=6)
    ref2 = orc.Geometry(moved, None).trace(rays2, tie_mode=1)
    compare_hits(ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays2))), ref2)
    geom.h = ctx.compact_geometry(geom.h)
    assert int(geom.header()["primNodeCount"]) == n_leaves
    compare_hits(ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays2))), ref2)
    geom.destroy()


def test_high_quality_scene_build_returns_the_fast_build_hits(ctx):
    """Scenes: the same builder over instance boxes; with spatial splits an instance may sit in several instance nodes (each with
    the part of its world box the leaf is responsible for), without them there is one node per instance. Update and compaction
    keep working on such a scene."""
    verts, idx = scenes.icosphere(3, 1.0)
    base = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    # long thin instances (stretched along one axis) between small ones: what instance splits are for
    frames = scenes.random_matrix_frames(3000, seed=7, extent=40.0)
    frames = frames.reshape(-1, 4, 4).copy()
    frames[::7, :3, 0] *= 25.0
    frames = frames.reshape(-1, 16)
    fast = ctx.build_scene([base] * len(frames), frames, api.FRAME_MATRIX)
    rays = scenes.random_rays(100000, np.full(3, -60.0, np.float32), np.full(3, 60.0, np.float32), seed=8)
    d_rays = ctx.rays_to_device(rays)
    a = ctx.hits_to_host(ctx.trace(fast, d_rays))
    assert (a["primID"] != api.INVALID).mean() > 0.02

    def same_hits(scene, exact=True):
        b = ctx.hits_to_host(ctx.trace(scene, d_rays))
        same = (a["primID"] == b["primID"]) & (a["instanceIDs"][:, 0] == b["instanceIDs"][:, 0])
        assert same.mean() > 0.999
        if exact:
            assert np.all(a["t"][~same] == b["t"][~same])  # (exact ties between overlapping instances may resolve either way)
            assert np.array_equal(a["t"][same], b["t"][same])
        else:
            # an update recomputes the world -> object matrices in another kernel than the build: the compiler contracts the
            # products of the inversion differently there, the matrices differ in the last bits (as build and update of the
            # reference do: two kernels), hit distances with them
            hit = same & (a["primID"] != api.INVALID)
            dt = np.abs(a["t"][hit] - b["t"][hit]) / np.maximum(np.abs(a["t"][hit]), 1e-20)
            assert dt.max() <= 1e-5

    for flags, split in ((api.HIGH_QUALITY_BUILD, True), (api.HIGH_QUALITY_BUILD | DISABLE_SPATIAL_SPLITS, False)):
        hq = ctx.build_scene([base] * len(frames), frames, api.FRAME_MATRIX, flags=flags)
        boxes, leaves = hq.nodes()
        if split:
            assert len(frames) < len(leaves) <= len(frames) + len(frames) // 2
            assert np.array_equal(np.unique(leaves["bits"] & 0xFFFFFF), np.arange(len(frames)))  # every instance is referenced
        else:
            assert len(leaves) == len(frames)
        H.validate_wide_bvh(boxes, leaves, duplicated_references=split)
        same_hits(hq)
        ctx.rebuild_scene(hq, op=api.UPDATE)  # refit in place: same instances, same frames
        same_hits(hq, exact=False)
        hq.h = ctx.compact_scene(hq.h)
        same_hits(hq, exact=False)
        hq.destroy()
    fast.destroy()
    base.destroy()


@pytest.mark.parametrize("name", ["sphere_1M", "terrain_2M"])
def test_large_high_quality_build(ctx, name):
    verts, idx = scenes.sphere_mesh(1000, 500) if name == "sphere_1M" else scenes.terrain_mesh(1000)
    d_v, d_i = to_dev(ctx, verts), to_dev(ctx, idx)
    fast = ctx.build(d_v, d_i)
    hq = ctx.build(d_v, d_i, flags=api.HIGH_QUALITY_BUILD)
    st = sbvh_state(hq)
    assert st["error"] == 0 and hq.build_error_flag() == 0
    ctx.torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.build_geometry(hq.inp, hq.h, hq.temp, api.HIGH_QUALITY_BUILD)
    ctx.torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    boxes, leaves = hq.nodes()
    cost_hq = H.validate_wide_bvh(boxes, leaves, len(idx), duplicated_references=True)
    cost_fast = H.validate_wide_bvh(*fast.nodes(), len(idx))
    print(f"{name}: high-quality build {ms:.2f} ms, {st['levels']} levels, {st['totalRefs']} references of {hq.ctx_prim_max()} max primitives, "
          f"cost {cost_hq:.2f} vs LBVH {cost_fast:.2f}")
    assert cost_hq < cost_fast
    lo, hi = verts.min(0), verts.max(0)
    d_rays = ctx.rays_to_device(scenes.random_rays(500000, lo - 0.2, hi + 0.2, seed=11))
    a = ctx.hits_to_host(ctx.trace(fast, d_rays))
    b = ctx.hits_to_host(ctx.trace(hq, d_rays))
    same = a["primID"] == b["primID"]
    assert np.all(a["t"][~same] == b["t"][~same]) and same.mean() > 0.999
    hit = same & (a["primID"] != api.INVALID)
    assert hit.mean() > 0.05 and np.array_equal(a["t"][hit], b["t"][hit]) and np.array_equal(a["uv"][hit], b["uv"][hit])
    fast.destroy(), hq.destroy()


def test_degenerate_inputs_terminate_and_stay_valid(ctx):
    """Inputs no plane can separate (1000 copies of one triangle: every split is by arrival, down to the depth bound), tiny
    geometries, zero-area triangles: the build ends, the tree is valid, rays find what the LBVH build finds."""
    one = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    cases = {
        "copies": np.tile(one, (1000, 1)),
        "two": np.concatenate([one, one + np.float32(2)]),
        "three": np.concatenate([one, one + np.float32(2), one * np.float32(0.5) - np.float32(1)]),
        "points": np.repeat(np.random.default_rng(3).random((500, 1, 3)).astype(np.float32), 3, axis=1).reshape(-1, 3),  # zero-area triangles
        "line": np.stack([np.linspace(0, 1, 900, dtype=np.float32), np.zeros(900, np.float32), np.zeros(900, np.float32)], 1),  # collinear
    }
    for name, verts in cases.items():
        verts = np.ascontiguousarray(verts, np.float32)
        hq = ctx.build(to_dev(ctx, verts), None, flags=api.HIGH_QUALITY_BUILD)
        st = sbvh_state(hq)
        assert st["error"] == 0 and hq.build_error_flag() == 0, (name, st)
        boxes, leaves = hq.nodes()
        H.validate_wide_bvh(boxes, leaves, len(verts) // 3, duplicated_references=True)
        fast = ctx.build(to_dev(ctx, verts), None)
        lo, hi = verts.min(0) - 0.5, verts.max(0) + 0.5
        d_rays = ctx.rays_to_device(scenes.random_rays(20000, lo, hi, seed=1))
        a, b = ctx.hits_to_host(ctx.trace(fast, d_rays)), ctx.hits_to_host(ctx.trace(hq, d_rays))
        assert np.array_equal(a["t"][a["primID"] != api.INVALID], b["t"][b["primID"] != api.INVALID]), name
        assert np.array_equal(a["primID"] != api.INVALID, b["primID"] != api.INVALID), name
        fast.destroy(), hq.destroy()


def test_several_high_quality_geometries_in_one_call_and_given_pairs(ctx):
    meshes = [scenes.sphere_mesh(40, 20), scenes.terrain_mesh(30), (scenes.random_soup(1500, seed=2), None)]
    inputs = [ctx.mesh_input(to_dev(ctx, v), None if i is None else to_dev(ctx, i)) for v, i in meshes]
    geoms = ctx.build_geometries(inputs, flags=api.HIGH_QUALITY_BUILD, batch_max=0)  # no batch builder: three SBVH builds, one temporary buffer
    for (v, i), g in zip(meshes, geoms):
        og = orc.Geometry(v, i, high_quality=True)
        boxes, leaves = g.nodes()
        cb, cl = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaves.view(orc.TRIPAIR_DTYPE))
        assert cb.tobytes() == og.box_nodes.tobytes() and cl.tobytes() == og.leaves.tobytes()
    ctx.destroy_geometries(geoms)
    # pairs given by the caller (hiprtTriangleMeshPrimitive::trianglePairIndices)
    verts, idx = scenes.sphere_mesh(40, 20)
    pairs = orc.Geometry(verts, idx).pairs
    g = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx), flags=api.HIGH_QUALITY_BUILD, pairs=to_dev(ctx, pairs))
    og = orc.Geometry(verts, idx, pairs=pairs, high_quality=True)
    boxes, leaves = g.nodes()
    cb, cl = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaves.view(orc.TRIPAIR_DTYPE))
    assert cb.tobytes() == og.box_nodes.tobytes() and cl.tobytes() == og.leaves.tobytes()
    g.destroy()
