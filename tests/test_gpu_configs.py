"""GPU parity at the FULL sizes of BASELINE.json's five configs (C1 lives in test_gpu_parity.py): the CUDA path through
the C-ABI against the CPU oracle on the same procedurally generated inputs, with the rays sampled so that the oracle
finishes in seconds. primID / instanceID exact (exact-t ties between different primitives counted and bounded), t and uv
within 1e-5; any-hit rays compared on occlusion."""
import numpy as np
import pytest

from hiprt_b200 import api, multi_gpu, scenes
from oracle import binding as orc
from tests import helpers as H
from tests.test_gpu_parity import check_build_against_oracle, compare_hits, to_dev
from tests.test_gpu_scene import pin_transforms

pytestmark = pytest.mark.gpu

W, HGT = 1920, 1080


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def sample(a, every, phase=0):
    return np.ascontiguousarray(a[phase::every])


def test_c2_sphere_1m_fast_build_primary_and_shadow(ctx):
    """C2: 1 000 000-triangle lat-long sphere, PreferFastBuild, 1920x1080 primary (closest) + shadow (any-hit) rays."""
    verts, idx = scenes.sphere_mesh(1000, 500)
    assert len(idx) == 1_000_000
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    og = orc.Geometry(verts, idx)
    assert int(geom.header()["primNodeCount"]) == og.prim_count == 500_000  # every quad pairs
    check_build_against_oracle(ctx, geom, og)  # sort, binary topology, wide nodes and leaves bit-exact for identical keys
    primary = scenes.pinhole_rays(W, HGT, (0.0, 0.0, 3.0), (0.0, 0.0, 0.0))[scenes.tile_order(W, HGT)]
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(primary)))
    sel = np.arange(3, len(primary), 11)
    ref = og.trace(primary[sel], tie_mode=1)
    assert (ref["primID"] != orc.INVALID).mean() > 0.25
    ties = compare_hits(got[sel], ref)
    assert ties <= 1e-3 * len(sel)
    shadow, _ = scenes.shadow_rays(primary, got, (2.0, 4.0, 3.0))
    occluded = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(shadow), kind=api.ANY_HIT))
    ssel = np.arange(1, len(shadow), 7)
    sref = og.trace(shadow[ssel], tie_mode=1)
    assert np.array_equal(occluded["primID"][ssel] != api.INVALID, sref["primID"] != orc.INVALID)
    assert 0.1 < (sref["primID"] != orc.INVALID).mean() < 0.9
    geom.destroy()


def test_c3_terrain_10m_balanced_build_ao_rays(ctx):
    """C3: 9 999 392-triangle fBm terrain, PreferBalancedBuild (H-PLOC), primary rays + 16-spp cosine-hemisphere AO rays."""
    verts, idx = scenes.terrain_mesh(2236)
    assert len(idx) == 9_999_392
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx), flags=api.BALANCED_BUILD)
    assert geom.build_error_flag() == 0
    og = orc.Geometry(verts, idx)  # hits do not depend on the builder
    assert int(geom.header()["primNodeCount"]) == og.prim_count
    primary = scenes.pinhole_rays(W, HGT, (0.5, 0.6, -0.6), (0.5, 0.0, 0.5))[scenes.tile_order(W, HGT)]
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(primary)))
    sel = np.arange(5, len(primary), 13)
    ref = og.trace(primary[sel], tie_mode=1)
    assert (ref["primID"] != orc.INVALID).mean() > 0.25
    ties = compare_hits(got[sel], ref)
    assert ties <= 1e-3 * len(sel)
    # AO: 16 rays per hit pixel of a pixel sample (incoherent, maxT = 0.05)
    pix = np.arange(2, len(primary), 97)
    ao = scenes.ao_rays(primary[pix], got[pix], spp=16)
    assert len(ao) > 100_000
    occluded = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(ao), kind=api.ANY_HIT))
    aref = og.trace(ao, tie_mode=1)
    assert np.array_equal(occluded["primID"] != api.INVALID, aref["primID"] != orc.INVALID)
    assert (aref["primID"] != orc.INVALID).mean() > 0.01
    # the same rays as closest-hit queries: exact primitives
    closest = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(ao)))
    compare_hits(closest, aref)
    geom.destroy()


def test_c4_two_level_scene_100_geometries_10k_instances(ctx):
    """C4: 100 geometries (128 .. 16 384 triangles), 10 000 instances with random rotation / scale / translation,
    hiprtBuildScene + scene traversal."""
    meshes = scenes.instanced_geometries(100)
    n_inst = 10_000
    geom_of = np.arange(n_inst) % len(meshes)
    frames = scenes.random_matrix_frames(n_inst, seed=42, extent=50.0)
    geoms = [ctx.build(to_dev(ctx, v), to_dev(ctx, i)) for v, i in meshes]
    ogeoms = [orc.Geometry(v, i) for v, i in meshes]
    scene = ctx.build_scene([geoms[g] for g in geom_of], frames)
    oscene = orc.Scene(ogeoms, geom_of, frames, 1)
    assert int(scene.header()["primNodeCount"]) == n_inst
    boxes, leaves = scene.nodes()
    assert np.array_equal(np.sort(leaves.view(orc.INSTANCE_NODE_DTYPE)["bits"] & 0xFFFFFF), np.arange(n_inst))
    pin_transforms(scene, oscene)  # matrices agree to 1e-5; the oracle then uses the GPU-built ones
    rays = scenes.pinhole_rays(W, HGT, (0.0, 15.0, 130.0), (0.0, 0.0, 0.0))[scenes.tile_order(W, HGT)]
    got = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays)))
    sel = np.arange(7, len(rays), 17)
    ref = oscene.trace(rays[sel])
    hit = ref["primID"] != orc.INVALID
    assert hit.mean() > 0.1
    ties = compare_hits(got[sel], ref)
    same = got["primID"][sel] == ref["primID"]
    assert np.array_equal(got["instanceIDs"][sel][same & hit, 0], ref["instanceIDs"][same & hit, 0])
    assert ties <= 2e-3 * len(sel)
    anyhit = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays[sel]), kind=api.ANY_HIT))
    assert np.array_equal(anyhit["primID"] != api.INVALID, hit)
    scene.destroy()
    for g in geoms:
        g.destroy()


def test_c5_batch_build_of_many_small_geometries(ctx):
    """C5 (build half): 100 000 geometries of 2 .. 512 triangles in one hiprtBuildGeometries call with
    batchBuildMaxPrimCount = 512; a sample of them is compared with the oracle, and the rank split used on N GPUs covers
    the list exactly once."""
    n_geoms = 100_000
    g = np.arange(n_geoms, dtype=np.uint64)
    sizes = (2 + ((g * 2654435761) % (2 ** 32) >> 7) % 511).astype(np.int64)
    assert sizes.min() >= 2 and sizes.max() <= 512
    rng = np.random.default_rng(5)
    total = int(sizes.sum())
    # one big vertex pool: geometry k owns triangles [first[k], first[k] + sizes[k]) (non-indexed soup in its own unit cell)
    first = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    centres = rng.random((total, 1, 3)).astype(np.float32)
    pool = (centres + (rng.random((total, 3, 3)).astype(np.float32) - 0.5) * 0.2).reshape(-1, 3)
    d_pool = to_dev(ctx, pool)
    inputs = [ctx.mesh_input(d_pool[3 * int(first[k]): 3 * int(first[k] + sizes[k])]) for k in range(n_geoms)]
    geoms = ctx.build_geometries(inputs, batch_max=512)
    ctx.torch.cuda.synchronize()
    for k in rng.choice(n_geoms, 48, replace=False):
        k = int(k)
        v = pool[3 * int(first[k]): 3 * int(first[k] + sizes[k])]
        hdr = geoms[k].header()
        assert int(hdr["primNodeCount"]) == sizes[k]
        boxes, leaves = geoms[k].nodes()
        H.validate_wide_bvh(boxes, leaves, int(sizes[k]))
        og = orc.Geometry(v, None, allow_pairing=False)
        rays = scenes.random_rays(512, v.min(0) - 0.05, v.max(0) + 0.05, seed=k)
        got = ctx.hits_to_host(ctx.trace(geoms[k], ctx.rays_to_device(rays)))
        compare_hits(got, og.trace(rays, tie_mode=1))
    # N-GPU split of the batch (hiprt_b200/multi_gpu.py): contiguous, complete, disjoint
    for world in (2, 8):
        parts = [multi_gpu.split_batch(n_geoms, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n_geoms
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
    ctx.destroy_geometries(geoms)
