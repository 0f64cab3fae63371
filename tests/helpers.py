"""Shared test helpers: ray generators restating the reference's test kernels, image diff."""
import os

import numpy as np

from oracle import binding as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32 = np.float32


def load_golden_image(name):
    return np.load(os.path.join(GOLDEN, "golden_images.npz"))[name]


def cornell_box():
    d = np.load(os.path.join(GOLDEN, "cornell_box.npz"))
    return d["vertices"], d["mat_indices"], d["diffuse"]


def cornell_rays(width=512, height=512):
    """Ray generator of CornellBoxKernel (reference test/kernels/HiprtTestKernel.h:118-134)."""
    x = np.arange(width, dtype=f32)[None, :].repeat(height, 0)
    y = np.arange(height, dtype=f32)[:, None].repeat(width, 1)
    dx = f32(2.0) * x / f32(width) - f32(1.0)
    dy = f32(2.0) * (f32(1.0) - y / f32(height)) - f32(1.0)
    d = np.stack([f32(-387.817566) * dx, f32(387.817566) * dy, np.full_like(dx, f32(1230.0))], -1)
    ln = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2], dtype=f32)
    d = (d / ln[..., None]).astype(f32)
    o = np.broadcast_to(np.array([278.0, 273.0, -900.0], f32), d.shape)
    return orc.make_rays(o.reshape(-1, 3), d.reshape(-1, 3))


def ortho_rays(width=512, height=512, max_t=orc.FLT_MAX):
    """Rays of MeshIntersectionKernel & friends (HiprtTestKernel.h:160-175): origin (x/w, y/h, -1), +z."""
    x = (np.arange(width, dtype=f32) / f32(width))[None, :].repeat(height, 0)
    y = (np.arange(height, dtype=f32) / f32(height))[:, None].repeat(width, 1)
    o = np.stack([x, y, np.full_like(x, f32(-1.0))], -1).reshape(-1, 3)
    d = np.broadcast_to(np.array([0, 0, 1], f32), o.shape)
    return orc.make_rays(o, d, max_t=max_t)


def image_matches(img, ref, max_bad_frac=0.003, tol=10):
    """The reference's image check (test/hiprtTest.cpp:943-1002): a pixel is bad when any
    channel differs by more than `tol`; pass when < 0.3 % of pixels are bad."""
    diff = np.abs(img.astype(np.int32) - ref.astype(np.int32))
    bad = (diff > tol).any(-1)
    return bad.mean() < max_bad_frac, float(bad.mean()), int(diff.max())


def tea16(v0, v1):
    """tea<16> of the reference's test/shared.h:117-131, vectorised over uint32 arrays."""
    v0 = np.asarray(v0, np.uint32).copy()
    v1 = np.asarray(v1, np.uint32).copy()
    s0 = np.uint32(0)
    with np.errstate(over="ignore"):
        for _ in range(16):
            s0 = np.uint32((int(s0) + 0x9E3779B9) & 0xFFFFFFFF)
            v0 += ((v1 << np.uint32(4)) + np.uint32(0xA341316C)) ^ (v1 + s0) ^ ((v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))
            v1 += ((v0 << np.uint32(4)) + np.uint32(0xAD90777D)) ^ (v0 + s0) ^ ((v0 >> np.uint32(5)) + np.uint32(0x7E95761E))
    return v0, v1
