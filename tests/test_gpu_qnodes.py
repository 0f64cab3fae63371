"""The quantised 64-byte traversal records (hiprt_b200/csrc/qnode.h) must not change a single bit of any hit.

A geometry built by hiprtBuildGeometry carries a Q region and the library kernels walk it; the SAME structure saved and loaded
again has none (files hold the public format only), so the library kernels walk its public 128-byte box nodes -- the round-1
kernel. Both handles are traced with the same rays: regular ones, and every kind the quantised walk hands to the exact path
(zero / negative-zero / denormal / huge direction components, origins on box planes, far-away and non-finite origins)."""
import numpy as np
import pytest

from hiprt_b200 import api, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return api.Context(0)


def to_dev(ctx, a):
    t = ctx.torch
    return t.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32 else np.ascontiguousarray(a)).to(ctx.device)


def _nasty_rays(lo, hi, n, seed):
    """Random rays plus the degenerate families, all aimed at the box [lo, hi]."""
    rng = np.random.default_rng(seed)
    rays = scenes.random_rays(n, lo, hi, seed=seed)
    ext = (hi - lo).astype(np.float32)
    k = n // 8
    d = rays["direction"]
    o = rays["origin"]
    # one / two zero direction components, negative zeros
    d[0 * k:1 * k, 0] = 0.0
    d[1 * k:2 * k, 1] = -0.0
    d[2 * k:3 * k, [0, 2]] = 0.0
    # denormal and tiny components, huge components
    d[np.arange(3 * k, 4 * k), rng.integers(0, 3, k)] = np.float32(1e-41)
    d[4 * k:4 * k + k // 2, 0] = np.float32(1e-25)
    d[4 * k + k // 2:5 * k, 1] = np.float32(3e24)
    # origins exactly on vertex coordinates of the mesh (box planes), axis-parallel directions
    sel = slice(5 * k, 6 * k)
    o[sel] = lo + ext * rng.integers(0, 9, (k, 3)).astype(np.float32) / np.float32(8.0)
    d[sel] = 0.0
    d[np.arange(5 * k, 6 * k), rng.integers(0, 3, k)] = rng.choice(np.array([-1.0, 1.0], np.float32), k)
    # far-away origins (beyond the Q grid's range) looking back at the scene, and a few non-finite ones
    sel = slice(6 * k, 6 * k + k // 2)
    o[sel] = (lo + hi) / 2 - d[sel] * np.float32(3e12)
    o[6 * k + k // 2:6 * k + k // 2 + 8, 0] = np.inf
    o[6 * k + k // 2 + 8:6 * k + k // 2 + 16, 1] = np.nan
    d[6 * k + k // 2 + 16:6 * k + k // 2 + 24, 2] = np.nan
    # short rays (AO-like) from inside the scene
    sel = slice(7 * k, 8 * k)
    rays["maxT"][sel] = np.float32(0.05) * ext.max()
    return rays


def _assert_same(a, b):
    assert np.array_equal(a["primID"], b["primID"])
    hit = a["primID"] != api.INVALID
    for f in ("t", "uv", "normal"):
        assert a[f][hit].tobytes() == b[f][hit].tobytes(), f


def _exact_copy(ctx, geom, tmp_path, name):
    path = str(tmp_path / name)
    ctx.save_geometry(geom.h, path)
    return api.Geometry(ctx, ctx.load_geometry(path), None, None, 0)


@pytest.mark.parametrize("mesh", ["terrain", "soup", "sphere", "flat"])
@pytest.mark.parametrize("flags", [api.FAST_BUILD, api.BALANCED_BUILD])
def test_quantised_walk_equals_exact_walk(ctx, tmp_path, mesh, flags):
    if mesh == "terrain":
        verts, idx = scenes.terrain_mesh(96)
    elif mesh == "sphere":
        verts, idx = scenes.sphere_mesh(64, 32)
    elif mesh == "flat":  # a planar grid: zero-extent nodes on one axis, coordinates on a power-of-two lattice
        g = np.arange(33, dtype=np.float32) / np.float32(32.0)
        x, z = np.meshgrid(g, g)
        verts = np.stack([x.ravel(), np.zeros(33 * 33, np.float32), z.ravel()], 1).astype(np.float32)
        q = np.arange(32 * 32)
        r, c = q // 32, q % 32
        a = (r * 33 + c).astype(np.uint32)
        idx = np.concatenate([np.stack([a, a + 1, a + 33], 1), np.stack([a + 1, a + 34, a + 33], 1)]).astype(np.uint32)
    else:
        verts, idx = scenes.random_soup(20000, seed=7).reshape(-1, 3), None
    d_verts = to_dev(ctx, verts)
    geom = ctx.build(d_verts, None if idx is None else to_dev(ctx, idx), flags=flags)
    exact = _exact_copy(ctx, geom, tmp_path, "g.bin")
    lo, hi = verts.min(0), verts.max(0)
    pad = np.maximum(hi - lo, np.float32(0.25))
    rays = _nasty_rays(lo - pad * np.float32(0.1), hi + pad * np.float32(0.1), 40000, seed=11)
    d_rays = ctx.rays_to_device(rays)
    a = ctx.hits_to_host(ctx.trace(geom, d_rays))
    b = ctx.hits_to_host(ctx.trace(exact, d_rays))
    assert (a["primID"] != api.INVALID).mean() > 0.1
    _assert_same(a, b)
    oa = ctx.hits_to_host(ctx.trace(geom, d_rays, kind=api.ANY_HIT))["primID"] != api.INVALID
    ob = ctx.hits_to_host(ctx.trace(exact, d_rays, kind=api.ANY_HIT))["primID"] != api.INVALID
    assert np.array_equal(oa, ob) and np.array_equal(oa, a["primID"] != api.INVALID)
    # the compact record formats take the same path
    t = ctx.torch
    rec = t.empty((len(rays), 4), dtype=t.int32, device=ctx.device)
    ctx.trace_format(geom, d_rays, api.FORMAT_COMPACT, rec, api.CLOSEST_HIT)
    rec = rec.cpu().numpy().view(api.COMPACT_DTYPE).reshape(-1)
    hit = a["primID"] != api.INVALID
    assert np.array_equal(rec["primID"], a["primID"]) and rec["t"][hit].tobytes() == a["t"][hit].tobytes()
    # the quantised boxes are conservative but tight: a few per cent more box visits, never fewer
    regular = scenes.random_rays(20000, lo, hi, seed=3)
    bq, lq = ctx.trace_stats(geom, ctx.rays_to_device(regular))
    be, le = ctx.trace_stats(exact, ctx.rays_to_device(regular))
    assert be <= bq <= be * 1.15 + 64, (bq, be)
    assert le <= lq, (lq, le)  # leaves reached >= leaves the exact walk tests
    exact.destroy()
    geom.destroy()


def test_q_region_follows_update_and_compaction(ctx, tmp_path):
    verts, idx = scenes.terrain_mesh(64)
    d_verts, d_idx = to_dev(ctx, verts), to_dev(ctx, idx)
    geom = ctx.build(d_verts, d_idx)
    rays = _nasty_rays(verts.min(0), verts.max(0) + np.float32(0.3), 30000, seed=5)
    d_rays = ctx.rays_to_device(rays)
    # refit with displaced vertices: the Q records are regenerated by the update
    moved = verts.copy()
    moved[:, 1] += (np.float32(0.15) * np.sin(np.float32(9.0) * moved[:, 0]) * np.cos(np.float32(7.0) * moved[:, 2])).astype(np.float32)
    d_verts.copy_(to_dev(ctx, moved))
    ctx.build_geometry(geom.inp, geom.h, geom.temp, geom.flags, op=api.UPDATE)
    exact = _exact_copy(ctx, geom, tmp_path, "u.bin")
    a = ctx.hits_to_host(ctx.trace(geom, d_rays))
    _assert_same(a, ctx.hits_to_host(ctx.trace(exact, d_rays)))
    assert (a["primID"] != api.INVALID).mean() > 0.1
    # compaction moves the structure; the Q region moves with it
    geom.h = ctx.compact_geometry(geom.h)
    _assert_same(ctx.hits_to_host(ctx.trace(geom, d_rays)), a)
    # and a second update after the compaction
    d_verts.copy_(to_dev(ctx, verts))
    ctx.build_geometry(geom.inp, geom.h, geom.temp, geom.flags, op=api.UPDATE)
    fresh = ctx.build(d_verts, d_idx)
    _assert_same(ctx.hits_to_host(ctx.trace(geom, d_rays)), ctx.hits_to_host(ctx.trace(_exact_copy(ctx, fresh, tmp_path, "f.bin"), d_rays)))
    exact.destroy()
    fresh.destroy()
    geom.destroy()
