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
