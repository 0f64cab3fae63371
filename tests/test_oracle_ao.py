"""CPU-only: the oracle against the reference's AoRayCornellBox.png (test/main.cpp:426-441, test/kernels/AoRayKernel.h): primary
closest-hit rays with a random sub-pixel offset, 32 cosine-weighted any-hit rays of length 1.4 per hit (the kernel's hit filter
is compiled out: UseFilter = false, test/shared.h:35, so the function table is null); grey = fraction of unoccluded rays. This is the headline workload's shape
(primary + AO rays, BASELINE.json configs[2]) and the ray generator bench.py's AO workload restates (tea<16> / lcg of
test/shared.h:95-150).

The reference averages 512 x 32 rays per pixel; here every 8th pixel in x and y is estimated from the first 64 of its 512
primary samples (the same seeds), so the comparison is statistical: the standard error of a pixel is <= 0.5 / sqrt(2048) = 2.8
grey levels."""
import numpy as np

from oracle import binding as orc
from tests import helpers as H

f32 = np.float32


def lcg(seed):
    """test/shared.h:101-108: in-place linear congruential step on a uint32 array, returns the 24 low bits."""
    seed *= np.uint32(1103515245)
    seed += np.uint32(12345)
    return seed & np.uint32(0x00FFFFFF)


def randf(seed):
    return lcg(seed).astype(f32) / f32(0x01000000)


def normalize(v):
    return (v / np.sqrt((v * v).sum(-1, dtype=f32), dtype=f32)[..., None]).astype(f32)


def sample_hemisphere_cosine(n, seed):
    """test/shared.h:127-139."""
    phi = f32(2.0 * np.pi) * randf(seed)
    sin_theta_sqr = randf(seed)
    sin_theta = np.sqrt(sin_theta_sqr, dtype=f32)
    axis = np.where((np.abs(n[:, 0]) > f32(0.001))[:, None], np.array([0, 1, 0], f32), np.array([1, 0, 0], f32))
    t = normalize(np.cross(axis, n).astype(f32))
    s = np.cross(n, t).astype(f32)
    d = s * (np.cos(phi, dtype=f32) * sin_theta)[:, None] + t * (np.sin(phi, dtype=f32) * sin_theta)[:, None] \
        + n * np.sqrt(f32(1.0) - sin_theta_sqr, dtype=f32)[:, None]
    return normalize(d.astype(f32))


def test_ao_ray_cornell_box_matches_reference_image_within_sampling_noise():
    stride, spp, ao_samples, ao_radius, res = 8, 64, 32, f32(1.4), 512
    d = np.load(H.os.path.join(H.GOLDEN, "cornell_obj.npz"))
    n_shapes = int(d["shape_count"])
    scene, _ = H.cornell_obj_scene()
    occluders = scene  # any-hit = "is there a hit in (minT, maxT)": the closest-hit query answers it
    is_light = np.concatenate([d[f"emission_{k}"].sum(1) > 0 for k in range(n_shapes)])
    first_face = np.cumsum([0] + [len(d[f"tris_{k}"]) for k in range(n_shapes)])

    ys, xs = np.meshgrid(np.arange(0, res, stride, dtype=np.uint32), np.arange(0, res, stride, dtype=np.uint32), indexing="ij")
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    origin = np.array([0.0, 2.5, 5.8], f32)
    dz = f32(f32(0.024) / f32(2.0 * np.tan(f32(45.0 * np.pi / 180.0) / f32(2.0))))
    unoccluded = np.zeros(len(xs), np.float64)
    for p in range(spp):
        seed = H.tea16(xs + ys * np.uint32(res), np.full_like(xs, p))[0].astype(np.uint32)
        offset = randf(seed)  # generateRay( ..., isMultiSamples = true ): one offset for x and y
        px = (xs.astype(f32) + offset) / f32(res) - f32(0.5)
        py = (ys.astype(f32) + offset) / f32(res) - f32(0.5)
        direction = normalize(np.stack([px * f32(0.024), -(py * f32(0.024)), np.full_like(px, -dz)], -1).astype(f32))
        hits = scene.trace(orc.make_rays(np.broadcast_to(origin, direction.shape), direction), tie_mode=0)
        hit = hits["primID"] != orc.INVALID
        face = first_face[np.where(hit, hits["instanceIDs"][:, 0], 0)] + np.where(hit, hits["primID"], 0)
        shade = hit & ~is_light[face]
        ng = hits["normal"].astype(f32)
        ng = np.where(((direction * ng).sum(-1) > 0)[:, None], -ng, ng)
        ng = normalize(np.where(shade[:, None], ng, f32(1.0)))
        surface = (origin + (hits["t"] * f32(1.0 - 1.0e-2))[:, None] * direction).astype(f32)
        for _ in range(ao_samples):
            ao_dir = sample_hemisphere_cosine(ng, seed)
            blocked = occluders.trace(orc.make_rays(surface, ao_dir, max_t=ao_radius), tie_mode=0)["primID"] != orc.INVALID
            unoccluded += shade & ~blocked
    grey = (unoccluded / (spp * ao_samples) * 255.0).astype(np.uint8).astype(np.int64)
    golden = H.load_golden_image("AoRayCornellBox")[::stride, ::stride].reshape(-1, 4).astype(np.int64)
    assert (golden[:, 0] == golden[:, 1]).all() and (golden[:, 3] == 255).all()
    diff = grey - golden[:, 0]  # measured: mean |diff| 1.22, mean diff 0.06, max |diff| 11
    assert np.abs(diff).mean() < 2.0, np.abs(diff).mean()
    assert abs(diff.mean()) < 0.4, diff.mean()  # no bias: filter, ray length, normal orientation and sampler all agree
    assert (np.abs(diff) > 14).mean() < 0.003, (np.abs(diff) > 14).mean()  # 5 sigma, the reference's own bad-pixel budget
