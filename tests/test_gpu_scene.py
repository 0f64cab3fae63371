"""GPU parity tests of the two-level path: hiprtBuildScene + scene traversal against the CPU oracle
(oracle/hiprt_oracle.cpp orc_scene_*, restating InstanceList.h:104-121, BvhNode.h:1109-1182,
hiprt_device_impl.h:930-1204)."""
import numpy as np
import pytest

from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests.test_gpu_parity import compare_hits, to_dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def build_pair(ctx, meshes, geom_of_inst, frames, masks=None, frame_type=api.FRAME_MATRIX):
    geoms = [ctx.build(to_dev(ctx, v), to_dev(ctx, i)) for v, i in meshes]
    ogeoms = [orc.Geometry(v, i) for v, i in meshes]
    scene = ctx.build_scene([geoms[g] for g in geom_of_inst], frames, frame_type, masks)
    oscene = orc.Scene(ogeoms, geom_of_inst, frames, 1 if frame_type == api.FRAME_MATRIX else 0, masks)
    return geoms, scene, oscene


def check_scene_structure(scene, oscene):
    boxes, leaves = scene.nodes()
    cb, cl = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaves.view(orc.INSTANCE_NODE_DTYPE))
    ob, ol = oscene.box_nodes, oscene.leaves
    assert len(cb) == len(ob)
    # child indices, counts
    assert np.array_equal(cb["child"], ob["child"])
    assert np.array_equal(cb["childCount"], ob["childCount"])
    # instance boxes: transformed corners, same arithmetic -> a few ulp at most (fma contraction differs)
    valid = cb["child"] != orc.INVALID
    np.testing.assert_allclose(cb["box"][valid], ob["box"][valid], rtol=2e-6, atol=1e-5)
    assert np.array_equal(cl["bits"], ol["bits"])
    assert np.array_equal(cl["mask"], ol["mask"])
    np.testing.assert_allclose(cl["invMatrix"], ol["invMatrix"], rtol=1e-5, atol=1e-6)


# Barycentrics inside an instance are conditioned by (distance to the instance) / (triangle edge): the world ray is
# moved into object space with a matrix that the GPU and the CPU oracle (and the reference, whose build kernels use
# fast-math) compute with different rounding (fma contraction), so an object-space offset of a few 1e-6 turns into
# ~1e-4 in uv on the small triangles of these meshes, and into a few 1e-5 in t at grazing incidence. The tests
# therefore split the claim in two: (1) the world->object matrices agree to 1e-5 (check_scene_structure), and
# (2) with the SAME matrices (the oracle adopts the ones the GPU build wrote: pin_transforms) hits agree to the
# strict bar: primID / instanceID exact, t and uv within 1e-5. SCENE_UV_TOL is the loose bound used where the
# matrices are NOT shared (comparison with the compiled reference).
SCENE_UV_TOL = 1e-3


def matrices_by_instance(leaves):
    leaves = np.asarray(leaves).view(orc.INSTANCE_NODE_DTYPE)
    return leaves["invMatrix"][np.argsort(leaves["bits"] & 0xFFFFFF)]


def pin_transforms(scene, oscene):
    """Checks that both sides built the same world->object matrices up to rounding, then makes the oracle use the
    GPU's so that the traversal comparison is held to the strict bar."""
    _, leaves = scene.nodes()
    mine, theirs = matrices_by_instance(leaves), matrices_by_instance(oscene.leaves)
    # the translation column is -dot(row, t): its absolute rounding error scales with |t|
    np.testing.assert_allclose(mine, theirs, rtol=1e-5, atol=5e-7 * max(1.0, float(np.abs(theirs).max())))
    oscene.adopt_transforms(leaves)


def camera_rays(extent, n=192):
    return scenes.pinhole_rays(n, n, (0.0, 0.3 * extent, 2.6 * extent), (0, 0, 0))


def test_singleton_scene_identity(ctx):
    v, i = scenes.sphere_mesh(32, 16)
    frames = np.zeros((1, 16), np.float32)
    frames[0, [0, 5, 10]] = 1.0
    geoms, scene, oscene = build_pair(ctx, [(v, i)], [0], frames)
    rays = camera_rays(1.0)
    got = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays)))
    ref = oscene.trace(rays)
    compare_hits(got, ref)
    assert np.array_equal(got["instanceIDs"][:, 0], ref["instanceIDs"][:, 0])
    # and the scene gives what the bare geometry gives
    g = ctx.hits_to_host(ctx.trace(geoms[0], ctx.rays_to_device(rays)))
    assert np.array_equal(g["primID"], got["primID"])
    scene.destroy()


@pytest.mark.parametrize("n_inst", [2, 7, 300])
def test_instanced_scene_hits_match_oracle(ctx, n_inst):
    meshes = scenes.instanced_geometries(6)
    geom_of = np.arange(n_inst) % len(meshes)
    frames = scenes.random_matrix_frames(n_inst, seed=n_inst, extent=6.0)
    geoms, scene, oscene = build_pair(ctx, meshes, geom_of, frames)
    assert int(scene.header()["primNodeCount"]) == n_inst
    check_scene_structure(scene, oscene)
    rays = camera_rays(6.0)
    got = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays)))
    compare_hits(got, oscene.trace(rays), rel=1e-4, uv_tol=SCENE_UV_TOL)  # independently built matrices
    pin_transforms(scene, oscene)
    ref = oscene.trace(rays)
    hit = ref["primID"] != orc.INVALID
    assert hit.mean() > (0.05 if n_inst >= 100 else 0.005)
    ties = compare_hits(got, ref, rel=1e-5)
    same = got["primID"] == ref["primID"]
    assert np.array_equal(got["instanceIDs"][same & hit, 0], ref["instanceIDs"][same & hit, 0])
    assert ties <= 0.001 * len(rays)
    # any-hit agrees with closest-hit on occlusion
    anyhit = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays), kind=api.ANY_HIT))
    assert np.array_equal(anyhit["primID"] != api.INVALID, got["primID"] != api.INVALID)
    scene.destroy()


def test_ray_mask_culls_instances(ctx):
    meshes = scenes.instanced_geometries(2)
    n_inst = 40
    geom_of = np.arange(n_inst) % 2
    frames = scenes.random_matrix_frames(n_inst, seed=5, extent=5.0)
    masks = np.where(np.arange(n_inst) % 2 == 0, 0x1, 0x2).astype(np.uint32)
    geoms, scene, oscene = build_pair(ctx, meshes, geom_of, frames, masks)
    pin_transforms(scene, oscene)
    rays = camera_rays(5.0, 128)
    for ray_mask in (0x1, 0x2, 0x3):
        got = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays), mask=ray_mask))
        ref = oscene.trace(rays, ray_mask=ray_mask)
        compare_hits(got, ref, rel=1e-5)
        hit = got["primID"] != api.INVALID
        if ray_mask != 0x3:
            assert np.all(masks[got["instanceIDs"][hit, 0]] == ray_mask)
    scene.destroy()


def test_srt_frames(ctx):
    meshes = scenes.instanced_geometries(3)
    n_inst = 20
    rng = np.random.default_rng(3)
    frames = np.zeros((n_inst, 12), np.float32)  # rotation(axis,angle) scale translation time
    axis = rng.normal(size=(n_inst, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    frames[:, 0:3] = axis
    frames[:, 3] = rng.uniform(0, 2 * np.pi, n_inst)
    frames[:, 4:7] = rng.uniform(0.5, 1.5, (n_inst, 3))
    frames[:, 7:10] = rng.uniform(-4, 4, (n_inst, 3))
    geoms, scene, oscene = build_pair(ctx, meshes, np.arange(n_inst) % 3, frames, frame_type=api.FRAME_SRT)
    rays = camera_rays(4.0, 128)
    got = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays)))
    pin_transforms(scene, oscene)
    ref = oscene.trace(rays)
    compare_hits(got, ref, rel=1e-5)
    scene.destroy()


def test_integer_translations_are_exact(ctx):
    """Pure translations by small integers: the inverse matrix and the object-space ray are exact in fp32 when the
    ray origin is representable, so hits must agree with the oracle to the strict tolerance."""
    meshes = scenes.instanced_geometries(3)
    n_inst = 27
    frames = np.zeros((n_inst, 16), np.float32)
    frames[:, [0, 5, 10]] = 1.0
    g = np.stack(np.meshgrid(np.arange(3), np.arange(3), np.arange(3), indexing="ij"), -1).reshape(-1, 3)
    frames[:, [3, 7, 11]] = (g * 4 - 4).astype(np.float32)
    geoms, scene, oscene = build_pair(ctx, meshes, np.arange(n_inst) % 3, frames)
    rays = scenes.pinhole_rays(160, 160, (0.0, 2.0, 16.0), (0, 0, 0))
    got = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(rays)))
    ref = oscene.trace(rays)
    assert (ref["primID"] != orc.INVALID).mean() > 0.05
    compare_hits(got, ref, rel=1e-5)
    assert np.array_equal(got["instanceIDs"][:, 0], ref["instanceIDs"][:, 0])
    scene.destroy()
