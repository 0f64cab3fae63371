"""CPU-only: the oracle's C++ restatement of key-frame interpolation (oracle/hiprt_oracle.cpp: qmix, frameAt, frameInvTransform)
against the independent numpy restatement the GPU motion tests use -- two restatements of Transform.h / Quaternion.h that must agree."""
import numpy as np

from oracle import binding as orc
from tests.motion_ref import pose_matrix


def random_keys(rng, n):
    times = np.sort(rng.uniform(0.0, 1.0, n))
    times[0] = 0.0
    keys = []
    for k in range(n):
        axis = rng.normal(size=3)
        keys.append((tuple(axis), float(rng.uniform(0.0, 3.0)), tuple(rng.uniform(0.5, 2.0, 3)), tuple(rng.uniform(-5, 5, 3)), float(times[k])))
    return keys


def srt_rows(keys):
    return np.asarray([list(a) + [ang] + list(s) + list(t) + [time, 0.0] for a, ang, s, t, time in keys], np.float32)


def test_interpolated_pose_matches_numpy_restatement():
    rng = np.random.default_rng(0)
    for trial in range(40):
        keys = random_keys(rng, int(rng.integers(1, 5)))
        for time in (0.0, 0.1, 0.5, 0.9, 1.0, 1.5, float(rng.uniform(0, 1))):
            m, _ = orc.motion_pose(srt_rows(keys), time)
            np.testing.assert_allclose(m, pose_matrix(keys, time), rtol=2e-5, atol=2e-5)


def test_motion_ray_transform_is_the_inverse_of_the_pose():
    rng = np.random.default_rng(1)
    for trial in range(20):
        keys = random_keys(rng, 3)
        time = float(rng.uniform(0, 1))
        ray = orc.make_rays(rng.uniform(-3, 3, (1, 3)).astype(np.float32), rng.normal(size=(1, 3)).astype(np.float32))[0]
        m, local = orc.motion_pose(srt_rows(keys), time, ray)
        a, t = m[:, :3].astype(np.float64), m[:, 3].astype(np.float64)
        inv = np.linalg.inv(a)
        np.testing.assert_allclose(local["origin"], inv @ (ray["origin"] - t), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(local["direction"], inv @ ray["direction"], rtol=1e-4, atol=1e-4)
        assert local["maxT"] == ray["maxT"] and local["minT"] == ray["minT"]


def test_motion_blur_matches_reference_image():
    """The reference's MotionBlur test (test/main.cpp:1995-2180, MotionBlurKernel HiprtTestKernel.h:328-359): two single-triangle
    instances with 3 and 2 hiprtFrameSRT key frames (translation and rotation about z), 32 time samples per pixel, closest hit,
    colour per instance, gamma 2.2. Pins the oracle's key-frame interpolation (orc_motion_pose) against the reference's image."""
    from tests import helpers as H

    s, off, pi = 0.15, 0.3, np.float32(np.pi)
    tri = np.array([[s * np.sin(0.0), s * np.cos(0.0), 0.0], [s * np.sin(pi * 2 / 3), s * np.cos(pi * 2 / 3), 0.0],
                    [s * np.sin(pi * 4 / 3), s * np.cos(pi * 4 / 3), 0.0]], np.float32)
    geom = orc.Geometry(tri, np.array([[0, 1, 2]], np.uint32))
    srt = lambda tr, angle, time: [0, 0, 1, angle, 1, 1, 1, tr[0], tr[1], tr[2], time, 0]
    keys = [np.array([srt((-0.25, -off, 0), 0.0, 0.0), srt((0.0, -off, 0), 0.0, 0.35), srt((0.25, -off, 0), pi * 0.25, 1.0)], np.float32),
            np.array([srt((0.0, off, 0), 0.0, 0.0), srt((0.0, off, 0), pi * 0.5, 1.0)], np.float32)]
    w = h = 512
    x = (np.arange(w, dtype=np.float32) / np.float32(w))[None, :].repeat(h, 0) - np.float32(0.5)
    y = (np.arange(h, dtype=np.float32) / np.float32(h))[:, None].repeat(w, 1) - np.float32(0.5)
    o = np.stack([x, y, np.full_like(x, -1.0)], -1).reshape(-1, 3)
    d = np.broadcast_to(np.array([0, 0, 1], np.float32), o.shape)
    colours = np.array([[1.0, 0.0, 0.5], [0.0, 0.5, 1.0]], np.float32)
    acc = np.zeros((w * h, 3), np.float32)
    samples = 32
    for i in range(samples):
        time = np.float32(i) / np.float32(samples)
        best_t = np.full(w * h, np.inf, np.float32)
        best_inst = np.full(w * h, -1, np.int64)
        for inst in range(2):
            m, _ = orc.motion_pose(keys[inst], float(time))
            a_inv = np.linalg.inv(m[:, :3].astype(np.float64))
            oo = ((o.astype(np.float64) - m[:, 3].astype(np.float64)) @ a_inv.T).astype(np.float32)
            dd = (d.astype(np.float64) @ a_inv.T).astype(np.float32)
            hits = geom.trace(orc.make_rays(oo, dd), tie_mode=0)
            closer = (hits["primID"] != orc.INVALID) & (hits["t"] < best_t)
            best_t[closer], best_inst[closer] = hits["t"][closer], inst
        hit = best_inst >= 0
        acc[hit] += colours[best_inst[hit]]
    col = np.power(acc / np.float32(samples), np.float32(1.0 / 2.2))
    img = np.zeros((w * h, 4), np.uint8)
    img[:, :3] = (col * np.float32(255)).astype(np.uint8)
    img[:, 3] = 255
    ok, bad, mx = H.image_matches(img.reshape(h, w, 4), H.load_golden_image("MotionBlur"))
    assert ok, (bad, mx)


def test_motion_blur_slerp_matches_reference_image():
    """The reference's MotionBlurSlerp test (test/main.cpp:2375-2513, MotionBlurSlerpKernel HiprtTestKernel.h:361-405): a sphere
    under two hiprtFrameMatrix key frames, the identity and diag(-1, -1, 1) -- half a turn about z, which only the decomposition
    of the matrix into a rotation and the slerp between the key frames renders as an arc (and whose direction hangs on the
    quaternion coming out of the decomposition with w = +0 exactly). Every second pixel, with the pixel's own `tea<16>` / `lcg`
    sequence for the jitter and the 512 sample times; the pose is taken from the oracle on a grid of 1/2048 in time."""
    from tests import helpers as H
    from tests.test_oracle_ao import randf

    f32 = np.float32
    m0, m1 = np.zeros(16, f32), np.zeros(16, f32)
    m0[[0, 5, 10]] = 1.0
    m1[[0, 5]], m1[10], m1[12] = -1.0, 1.0, 1.0
    grid = 2048
    inverse = np.zeros((grid + 1, 3, 3), np.float64)  # world -> object rotation of the pose at k / grid (no scale, no translation)
    for k in range(grid + 1):
        m, _ = orc.motion_pose(np.stack([m0, m1]), k / grid, frame_type=1)
        assert np.allclose(m[:, 3], 0) and np.allclose(m[:, :3] @ m[:, :3].T, np.eye(3), atol=1e-5)
        inverse[k] = m[:, :3].T

    res, stride, samples = 512, 2, 512
    ys, xs = np.meshgrid(np.arange(0, res, stride, dtype=np.uint32), np.arange(0, res, stride, dtype=np.uint32), indexing="ij")
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    seed = H.tea16(xs, ys)[0].astype(np.uint32)
    dx = f32(2.0) * ((xs.astype(f32) + randf(seed)) / f32(res)) - f32(1.0)
    dy = f32(2.0) * ((ys.astype(f32) + randf(seed)) / f32(res)) - f32(1.0)
    d = np.stack([dx * f32(1.576494), dy * f32(1.576494), np.full_like(dx, -5.0)], -1).astype(np.float64)
    d /= np.linalg.norm(d, axis=1)[:, None]
    o = np.array([0.0, 0.0, 5.0])
    centre, radius = np.array([-1.0, 0.0, 0.0]), 0.5
    count = np.zeros(len(xs), np.int64)
    for _ in range(samples):
        k = np.rint(randf(seed).astype(np.float64) * grid).astype(np.int64)
        r = inverse[k]
        oc = np.einsum("nij,j->ni", r, o) - centre
        dd = np.einsum("nij,nj->ni", r, d)
        b = (oc * dd).sum(1)
        count += (b * b - ((oc * oc).sum(1) - radius * radius) >= 0) & (b < 0)
    frac = (count / samples).astype(f32)[:, None]
    colour = frac * np.array([0.9, 0.1, 0.1], f32) + (f32(1.0) - frac) * f32(0.1)
    img = (np.power(colour, f32(1.0 / 2.2)) * f32(255)).astype(np.uint8)
    golden = H.load_golden_image("MotionBlurSlerp")[::stride, ::stride, :3].reshape(-1, 3)
    diff = np.abs(img.astype(np.int64) - golden.astype(np.int64)).max(1)
    assert 0.3 < (count > 0).mean() < 0.5
    assert diff.max() <= 4, diff.max()  # measured: max 2, mean 0.015 -- the reference's sample sequence is reproduced exactly
