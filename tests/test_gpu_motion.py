"""Motion blur and transform queries (reference: Transform.h:410-517, hiprt_device_impl.h:960-998,1775-1927).
Animated instances (several key frames) are traced from a user kernel at a given time; the expected pose is a numpy
restatement of the key-frame interpolation (linear scale / translation, shortest-path slerp), turned into a matrix frame of a
STATIC oracle scene. The arithmetic differs (quaternion inverse of two points vs inverse matrix), hence the looser t bound.
The geometry sits off-centre in object space, so a rotating instance sweeps an arc that leaves the union of its key poses:
hits at intermediate times also prove the motion bounds (key frames + 3 interpolated steps) of the scene builder."""
import ctypes as C

import numpy as np
import pytest

from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests.test_gpu_parity import compare_hits, to_dev
from tests.test_gpu_user_kernels import Launcher, build_kernels, dev_ptr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


from tests.motion_ref import pose_matrix, quat_from_axis_angle, quat_matrix, quat_mix  # noqa: E402,F401


INSTANCES = [
    [((0, 1, 0), 0.4, (1.0, 1.0, 1.0), (-6.0, 0.0, 0.0), 0.0)],                                                  # static
    [((0, 0, 1), 0.0, (1.0, 1.0, 1.0), (0.0, -1.0, 0.0), 0.0), ((0, 0, 1), 2.2, (1.5, 1.5, 1.5), (1.0, 1.0, 0.5), 1.0)],  # 2 keys
    [((1, 0, 0), 0.3, (0.8, 1.0, 1.2), (7.0, -2.0, 0.0), 0.0), ((0, 1, 0), 1.5, (1.0, 1.0, 1.0), (7.0, 0.0, 1.0), 0.5),
     ((0, 1, 1), 2.5, (1.2, 0.9, 1.0), (6.0, 2.0, -1.0), 1.0)],                                                  # 3 keys
]


def build_animated_scene(ctx):
    verts, idx = scenes.sphere_mesh(48, 24)
    verts = (verts * np.float32(0.8) + np.array([2.5, 0.0, 0.0], np.float32)).astype(np.float32)  # off-centre in object space
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    frames, headers = [], []
    for keys in INSTANCES:
        headers.append((len(frames), len(keys)))
        for axis, angle, scale, trans, time in keys:
            frames.append(list(axis) + [angle] + list(scale) + list(trans) + [time, 0.0])
    frames = np.asarray(frames, np.float32)
    scene = ctx.build_scene([geom] * len(INSTANCES), frames, frame_type=api.FRAME_SRT, transform_headers=np.asarray(headers, np.uint32))
    return geom, scene, (verts, idx)


def static_oracle_scene(mesh, time):
    og = orc.Geometry(*mesh)
    mats = np.zeros((len(INSTANCES), 16), np.float32)
    for k, keys in enumerate(INSTANCES):
        mats[k, :12] = pose_matrix(keys, time).reshape(-1)
    return orc.Scene([og], np.zeros(len(INSTANCES), np.uint32), mats, 1), og


@pytest.mark.parametrize("time", [0.0, 0.2, 0.5, 0.75, 1.0, 1.7])
def test_animated_instances_traced_at_time(ctx, time):
    geom, scene, mesh = build_animated_scene(ctx)
    (fn,) = build_kernels(ctx, ["motionSceneKernel"])
    rays = scenes.pinhole_rays(384, 256, (0.5, 1.0, 17.0), (0.5, 0.0, 0.0))
    d_rays = ctx.rays_to_device(rays)
    d_hits = ctx.torch.zeros((len(rays), 12), dtype=ctx.torch.int32, device=ctx.device)
    Launcher().launch(fn, len(rays), 64, [scene.h, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(len(rays)), C.c_float(time)])
    ctx.torch.cuda.synchronize()
    got = ctx.hits_to_host(d_hits)
    oscene, _ = static_oracle_scene(mesh, time)
    ref = oscene.trace(rays)
    hit = ref["primID"] != orc.INVALID
    assert hit.mean() > 0.015
    # every instance is seen at this time
    assert set(np.unique(ref["instanceIDs"][hit, 0])) == {0, 1, 2}
    silhouette = got["primID"] != ref["primID"]
    # the two sides compute the object-space ray differently: rays grazing a silhouette may fall on either side
    assert silhouette.mean() < 2e-3
    ok = ~silhouette
    compare_hits(got[ok], ref[ok], rel=2e-4, uv_tol=2e-2)  # uv: (distance to the instance) / (triangle edge) amplifies the 1e-6 object-space difference
    assert np.array_equal(got["instanceIDs"][ok & hit, 0], ref["instanceIDs"][ok & hit, 0])
    if time == 0.0:
        # the library's own scene kernel traces animated instances in their pose at time 0
        lib_hits = ctx.hits_to_host(ctx.trace(scene, d_rays))
        assert np.array_equal(lib_hits["primID"], got["primID"]) and np.array_equal(lib_hits["instanceIDs"][:, 0], got["instanceIDs"][:, 0])
    scene.destroy()
    geom.destroy()


def test_animated_instance_nodes_and_motion_bounds(ctx):
    geom, scene, mesh = build_animated_scene(ctx)
    boxes, leaves = scene.nodes()
    leaves = leaves.view(orc.INSTANCE_NODE_DTYPE)
    by_id = leaves[np.argsort(leaves["bits"] & 0xFFFFFF)]
    static_bit, identity_bit = (by_id["bits"] >> 25) & 1, (by_id["bits"] >> 26) & 1
    assert static_bit.tolist() == [1, 0, 0] and identity_bit.tolist() == [0, 0, 0]
    # animated nodes carry {frameIndex, frameCount} where the matrix would be (BvhNode.h:1152-1168)
    assert by_id["invMatrix"][1].view(np.uint32).reshape(-1)[:2].tolist() == [1, 2]
    assert by_id["invMatrix"][2].view(np.uint32).reshape(-1)[:2].tolist() == [3, 3]
    # the scene box contains the object at densely sampled times
    valid = boxes["child"] != orc.INVALID
    lo = boxes["box"][:, :, 0, :][valid].min(0)
    hi = boxes["box"][:, :, 1, :][valid].max(0)
    verts = mesh[0].astype(np.float64)
    for keys in INSTANCES[1:]:
        for time in np.linspace(0.0, 1.0, 41):
            m = pose_matrix(keys, float(time))
            w = verts @ m[:, :3].T + m[:, 3]
            slack = 0.06 * np.abs(hi - lo).max()  # 3 interpolation steps bound an arc only approximately (as in the reference)
            assert (w.min(0) >= lo - slack).all() and (w.max(0) <= hi + slack).all(), f"object leaves the scene box at time {time}"
    scene.destroy()
    geom.destroy()


def test_transform_queries(ctx):
    geom, scene, _ = build_animated_scene(ctx)
    (fn,) = build_kernels(ctx, ["transformQueryKernel"])
    rng = np.random.default_rng(2)
    pts = rng.uniform(-2, 2, (256, 3)).astype(np.float32)
    d_in = to_dev(ctx, pts)
    for inst, time in ((0, 0.0), (1, 0.5), (2, 0.3), (2, 0.9)):
        d_out = ctx.torch.zeros((len(pts) * 4, 3), dtype=ctx.torch.float32, device=ctx.device)
        Launcher().launch(fn, len(pts), 64, [scene.h, C.c_uint32(inst), dev_ptr(d_in), dev_ptr(d_out), C.c_uint32(len(pts)), C.c_float(time)])
        ctx.torch.cuda.synchronize()
        out = d_out.cpu().numpy().reshape(len(pts), 4, 3)
        m = pose_matrix(INSTANCES[inst], time)
        np.testing.assert_allclose(out[:, 0], pts @ m[:, :3].T + m[:, 3], rtol=2e-5, atol=2e-5)      # point object -> world
        np.testing.assert_allclose(out[:, 1], pts, rtol=2e-5, atol=2e-5)                                # and back
        normal_matrix = np.linalg.inv(m[:, :3]).T                                                       # vectors are normal-like
        np.testing.assert_allclose(out[:, 2], pts @ normal_matrix.T, rtol=5e-5, atol=5e-5)
        np.testing.assert_allclose(out[:, 3], pts, rtol=5e-5, atol=5e-5)
    scene.destroy()
    geom.destroy()


def test_transform_queries_of_a_scene_without_frames_are_the_identity(ctx):
    """A scene built without instanceFrames has a NULL frame array in its header, as the reference's arena hands out for an empty
    array (MemoryArena.h:38-40), and the transform queries treat that as the identity (Transform.h:421) -- found by the reference's
    own unit tests (NormalsCornellBox, ShadowRay / AoRayCornellBox call hiprtVectorObjectToWorld on such a scene)."""
    verts, idx = scenes.sphere_mesh(16, 8)
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    scene = ctx.build_scene([geom, geom], None)
    assert int(scene.header()["frames"]) == 0 and int(scene.header()["frameCount"]) == 0
    (fn,) = build_kernels(ctx, ["transformQueryKernel"])
    pts = np.random.default_rng(4).uniform(-2, 2, (128, 3)).astype(np.float32)
    d_in = to_dev(ctx, pts)
    for inst in (0, 1):
        d_out = ctx.torch.zeros((len(pts) * 4, 3), dtype=ctx.torch.float32, device=ctx.device)
        Launcher().launch(fn, len(pts), 64, [scene.h, C.c_uint32(inst), dev_ptr(d_in), dev_ptr(d_out), C.c_uint32(len(pts)), C.c_float(0.0)])
        ctx.torch.cuda.synchronize()
        out = d_out.cpu().numpy().reshape(len(pts), 4, 3)
        for k in range(4):
            assert np.array_equal(out[:, k], pts)
    compact = ctx.compact_scene(scene.h)
    scene.h = compact
    assert int(scene.header()["frames"]) == 0
    scene.destroy()
    geom.destroy()


def test_library_entry_point_traces_at_a_time(ctx):
    """hiprtB200TraceSceneAtTime (uniform time, and one time per ray) against the user kernel built on hiprtSceneTraversalClosest( ..., time )."""
    geom, scene, _ = build_animated_scene(ctx)
    t = ctx.torch
    (fn,) = build_kernels(ctx, ["motionSceneKernel"])
    rays = scenes.pinhole_rays(256, 160, (0.5, 1.0, 17.0), (0.5, 0.0, 0.0))
    d_rays = ctx.rays_to_device(rays)
    per_time = {}
    for time in (0.0, 0.35, 0.8):
        d_hits = t.zeros((len(rays), 12), dtype=t.int32, device=ctx.device)
        Launcher().launch(fn, len(rays), 64, [scene.h, dev_ptr(d_rays), dev_ptr(d_hits), C.c_uint32(len(rays)), C.c_float(time)])
        t.cuda.synchronize()
        want = ctx.hits_to_host(d_hits)
        got = ctx.hits_to_host(ctx.trace_scene_at_time(scene, d_rays, time=time))
        assert (want["primID"] != api.INVALID).mean() > 0.015
        assert np.array_equal(got["primID"], want["primID"]) and np.array_equal(got["instanceIDs"][:, 0], want["instanceIDs"][:, 0])
        hit = want["primID"] != api.INVALID
        assert np.array_equal(got["t"][hit], want["t"][hit])
        per_time[time] = want
    # one time per ray: ray k is traced at times[k]
    choice = np.arange(len(rays)) % 3
    times = np.array([0.0, 0.35, 0.8], np.float32)[choice]
    got = ctx.hits_to_host(ctx.trace_scene_at_time(scene, d_rays, times=to_dev(ctx, times)))
    for k, time in enumerate((0.0, 0.35, 0.8)):
        sel = choice == k
        assert np.array_equal(got["primID"][sel], per_time[time]["primID"][sel])
        assert np.array_equal(got["instanceIDs"][sel, 0], per_time[time]["instanceIDs"][sel, 0])
    # and any-hit agrees on occlusion
    occl = ctx.hits_to_host(ctx.trace_scene_at_time(scene, d_rays, time=0.35, kind=api.ANY_HIT))
    assert np.array_equal(occl["primID"] != api.INVALID, per_time[0.35]["primID"] != api.INVALID)
    scene.destroy()
    geom.destroy()
