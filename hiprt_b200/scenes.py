"""Procedural scenes and ray generators for the BASELINE.json configs.

Everything is generated with numpy on the host from fixed seeds, so the CPU oracle, the
reference and the CUDA path all see bit-identical inputs.
  C1  cornell box (tests/golden/cornell_box.npz) + CornellBoxKernel rays
  C2  lat-long sphere 1000x500 quads = 1 000 000 indexed triangles, pinhole primary + shadow rays
  C3  2236x2236-quad fBm heightfield = 9 999 392 triangles, primary + cosine-hemisphere AO rays
  C4  100 geometries x 10 000 matrix-frame instances
"""
import numpy as np

f32 = np.float32
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("minT", "<f4"), ("direction", "<f4", 3), ("maxT", "<f4")])
FLT_MAX = f32(3.402823466e38)


def make_rays(origins, directions, min_t=0.0, max_t=FLT_MAX):
    origins = np.asarray(origins, f32).reshape(-1, 3)
    directions = np.asarray(directions, f32).reshape(-1, 3)
    r = np.zeros(len(origins), RAY_DTYPE)
    r["origin"], r["direction"] = origins, directions
    r["minT"], r["maxT"] = min_t, max_t
    return r


# ---------------------------------------------------------------------------------------------
# meshes
def quad_grid_indices(nx, ny):
    """Two triangles per quad, the two halves of a quad consecutive (so pairing finds them)."""
    i = np.arange(nx, dtype=np.uint32)[None, :]
    j = np.arange(ny, dtype=np.uint32)[:, None]
    v00 = j * (nx + 1) + i
    v10 = v00 + 1
    v01 = v00 + (nx + 1)
    v11 = v01 + 1
    tris = np.stack([np.stack([v00, v10, v11], -1), np.stack([v00, v11, v01], -1)], 2)  # [ny,nx,2,3]
    return np.ascontiguousarray(tris.reshape(-1, 3).astype(np.uint32))


def grid_mesh(nx, ny, z=0.0):
    """Planar grid over [0,1]^2 (all coplanar: exercises the 2-D Morton branch)."""
    x = np.linspace(0, 1, nx + 1, dtype=f32)
    y = np.linspace(0, 1, ny + 1, dtype=f32)
    xx, yy = np.meshgrid(x, y)
    v = np.stack([xx, yy, np.full_like(xx, f32(z))], -1).reshape(-1, 3).astype(f32)
    return v, quad_grid_indices(nx, ny)


def sphere_mesh(n_lon=1000, n_lat=500, radius=1.0):
    """Lat-long sphere: n_lon x n_lat quads (C2: 1000 x 500 -> 1 000 000 triangles, 501 501 vertices)."""
    lon = np.linspace(0.0, 2.0 * np.pi, n_lon + 1)
    lat = np.linspace(0.0, np.pi, n_lat + 1)
    lo, la = np.meshgrid(lon, lat)
    v = np.stack([np.sin(la) * np.cos(lo), np.cos(la), np.sin(la) * np.sin(lo)], -1) * radius
    return np.ascontiguousarray(v.reshape(-1, 3).astype(f32)), quad_grid_indices(n_lon, n_lat)


def _hash2(ix, iy, seed):
    h = (ix.astype(np.uint64) * np.uint64(374761393) + iy.astype(np.uint64) * np.uint64(668265263) + np.uint64(seed) * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h = ((h ^ (h >> np.uint64(13))) * np.uint64(1274126177)) & np.uint64(0xFFFFFFFF)
    h = h ^ (h >> np.uint64(16))
    return (h & np.uint64(0xFFFFFF)).astype(np.float64) / float(1 << 24)


def _value_noise(x, y, seed):
    ix, iy = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
    fx, fy = x - ix, y - iy
    sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    a = _hash2(ix, iy, seed)
    b = _hash2(ix + 1, iy, seed)
    c = _hash2(ix, iy + 1, seed)
    d = _hash2(ix + 1, iy + 1, seed)
    return (a * (1 - sx) + b * sx) * (1 - sy) + (c * (1 - sx) + d * sx) * sy


def terrain_mesh(n=2236, seed=1234, amplitude=0.15, octaves=6):
    """fBm heightfield over [0,1]^2 with n x n quads (C3: n=2236 -> 9 999 392 triangles)."""
    x = np.linspace(0, 1, n + 1)
    xx, zz = np.meshgrid(x, x)
    h = np.zeros_like(xx)
    amp, freq, norm = 1.0, 4.0, 0.0
    for o in range(octaves):
        h += amp * _value_noise(xx * freq, zz * freq, seed + o)
        norm += amp
        amp *= 0.5
        freq *= 2.0
    h = (h / norm - 0.5) * 2.0 * amplitude
    v = np.stack([xx, h, zz], -1).reshape(-1, 3).astype(f32)
    return np.ascontiguousarray(v), quad_grid_indices(n, n)


def random_soup(n_tris, seed=0, extent=1.0, size=0.05):
    """Non-indexed triangle soup (vertices 3t..3t+2): never pairs."""
    rng = np.random.default_rng(seed)
    c = rng.random((n_tris, 1, 3)) * extent
    v = c + (rng.random((n_tris, 3, 3)) - 0.5) * size
    return np.ascontiguousarray(v.reshape(-1, 3).astype(f32))


def icosphere(subdiv=2, radius=1.0):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdiv):
        cache, verts, faces = {}, list(map(tuple, v)), []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = (np.array(verts[a]) + np.array(verts[b])) * 0.5
                m /= np.linalg.norm(m)
                verts.append(tuple(m))
                cache[k] = len(verts) - 1
            return cache[k]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            faces += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v, f = np.array(verts), np.array(faces, np.int64)
    return np.ascontiguousarray((v * radius).astype(f32)), np.ascontiguousarray(f.astype(np.uint32))


# ---------------------------------------------------------------------------------------------
# cameras and secondary rays (RNG: tea<16> + lcg of the reference's test/shared.h:103-131)
def tea16(v0, v1):
    v0 = np.asarray(v0, np.uint32).copy()
    v1 = np.asarray(v1, np.uint32).copy()
    s0 = 0
    with np.errstate(over="ignore"):
        for _ in range(16):
            s0 = (s0 + 0x9E3779B9) & 0xFFFFFFFF
            s = np.uint32(s0)
            v0 += ((v1 << np.uint32(4)) + np.uint32(0xA341316C)) ^ (v1 + s) ^ ((v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))
            v1 += ((v0 << np.uint32(4)) + np.uint32(0xAD90777D)) ^ (v0 + s) ^ ((v0 >> np.uint32(5)) + np.uint32(0x7E95761E))
    return v0, v1


def lcg(seed):
    with np.errstate(over="ignore"):
        seed = (np.uint32(1103515245) * seed + np.uint32(12345)).astype(np.uint32)
    return seed, seed & np.uint32(0x00FFFFFF)


def randf(seed):
    seed, v = lcg(seed)
    return seed, v.astype(f32) / f32(0x01000000)


def normalize(v):
    return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(f32)


def pinhole_rays(width, height, eye, look_at, up=(0, 1, 0), fov_deg=45.0):
    """Primary rays through pixel centres (offset 0.5 as the reference's generateRay, test/shared.h:151-168)."""
    eye = np.asarray(eye, np.float64)
    fwd = np.asarray(look_at, np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, np.float64))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    th = np.tan(np.radians(fov_deg) * 0.5)
    x = (np.arange(width) + 0.5) / width * 2 - 1
    y = 1 - (np.arange(height) + 0.5) / height * 2
    xx, yy = np.meshgrid(x * th * width / height, y * th)
    d = fwd[None, None, :] + xx[..., None] * right + yy[..., None] * upv
    d = normalize(d).reshape(-1, 3)
    o = np.broadcast_to(eye.astype(f32), d.shape)
    return make_rays(o, d)


def tile_order(width, height, tw=8, th=4):
    """Permutation of the row-major pixel indices into tw x th tiles (tile after tile, row-major inside a tile):
    32 consecutive rays cover an 8x4 pixel block, which is how the reference's tests launch their kernels
    (2-D thread blocks, test/hiprtTest.cpp:1510-1514), instead of a 32x1 strip."""
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    key = ((ys // th) * ((width + tw - 1) // tw) + xs // tw) * (tw * th) + (ys % th) * tw + xs % tw
    return np.argsort(key.reshape(-1), kind="stable")


def cornell_rays(width=512, height=512):
    """CornellBoxKernel ray generator (reference test/kernels/HiprtTestKernel.h:118-134)."""
    x = np.arange(width, dtype=f32)[None, :].repeat(height, 0)
    y = np.arange(height, dtype=f32)[:, None].repeat(width, 1)
    dx = f32(2.0) * x / f32(width) - f32(1.0)
    dy = f32(2.0) * (f32(1.0) - y / f32(height)) - f32(1.0)
    d = np.stack([f32(-387.817566) * dx, f32(387.817566) * dy, np.full_like(dx, f32(1230.0))], -1)
    ln = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2], dtype=f32)
    d = (d / ln[..., None]).astype(f32)
    o = np.broadcast_to(np.array([278.0, 273.0, -900.0], f32), d.shape)
    return make_rays(o.reshape(-1, 3), d.reshape(-1, 3))


def hit_points(rays, hits, back_off=1e-2):
    """World-space hit points pulled back along the ray by t*back_off (AoRayKernel.h:83)."""
    t = hits["t"].astype(f32) * f32(1.0 - back_off)
    return (rays["origin"] + rays["direction"] * t[:, None]).astype(f32)


def shadow_rays(rays, hits, light=(2.0, 4.0, 3.0)):
    """One any-hit ray per hit pixel toward a point light, maxT = 0.99 * distance."""
    ok = hits["primID"] != 0xFFFFFFFF
    p = hit_points(rays[ok], hits[ok], back_off=1e-3)
    d = np.asarray(light, f32)[None, :] - p
    dist = np.linalg.norm(d, axis=1).astype(f32)
    return make_rays(p, (d / dist[:, None]).astype(f32), 0.0, (dist * f32(0.99)).astype(f32)), ok


def ao_rays(rays, hits, spp=16, radius=0.05, pixel_ids=None):
    """`spp` cosine-hemisphere any-hit rays per hit pixel (sampleHemisphereCosine, test/shared.h:133-144). The sampler is seeded
    with tea<16>(pixel, sample): `pixel` is the ray's position in `rays`, or pixel_ids[position] when given (sharded frames seed
    by the global pixel index, so the rays do not depend on which rank owns the pixel)."""
    ok = np.nonzero(hits["primID"] != 0xFFFFFFFF)[0]
    p = hit_points(rays[ok], hits[ok])
    n = normalize(hits["normal"][ok].astype(np.float64))
    # face the normal against the incoming ray
    flip = (n * rays["direction"][ok]).sum(1) > 0
    n[flip] = -n[flip]
    pix = np.repeat((ok if pixel_ids is None else np.asarray(pixel_ids)[ok]).astype(np.uint32), spp)
    smp = np.tile(np.arange(spp, dtype=np.uint32), len(ok))
    seed, _ = tea16(pix, smp)
    seed, u0 = randf(seed)
    seed, u1 = randf(seed)
    phi = f32(2.0 * np.pi) * u0
    sin_t = np.sqrt(u1, dtype=f32)
    nn = np.repeat(n, spp, 0)
    axis = np.where((np.abs(nn[:, 0]) > 0.001)[:, None], np.array([0, 1, 0], f32), np.array([1, 0, 0], f32))
    t = normalize(np.cross(axis, nn))
    s = np.cross(nn, t).astype(f32)
    d = s * (np.cos(phi) * sin_t)[:, None] + t * (np.sin(phi) * sin_t)[:, None] + nn * np.sqrt(f32(1.0) - u1, dtype=f32)[:, None]
    return make_rays(np.repeat(p, spp, 0), normalize(d), 0.0, f32(radius))


def random_rays(n, box_lo, box_hi, seed=7):
    """Incoherent rays: origins on a sphere around the box, aimed at random points inside it."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(box_lo, np.float64), np.asarray(box_hi, np.float64)
    c, r = (lo + hi) * 0.5, np.linalg.norm(hi - lo) * 0.75 + 1e-3
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    o = c + dirs * r
    tgt = lo + rng.random((n, 3)) * (hi - lo)
    return make_rays(o.astype(f32), normalize(tgt - o))


# ---------------------------------------------------------------------------------------------
# two-level scenes (config C4: geometries instanced with random rigid + uniform-scale transforms)
def random_matrix_frames(n, seed=42, extent=50.0, scale_range=(0.5, 2.0)):
    """hiprtFrameMatrix[n] as float32 [n,16]: rows of the 3x4 object-to-world matrix, time, padding.
    Translation uniform in [-extent, extent]^3, rotation uniform (quaternion from 3 uniforms),
    uniform scale in scale_range."""
    rng = np.random.default_rng(seed)
    u = rng.random((n, 3))
    q = np.stack([np.sqrt(1 - u[:, 0]) * np.sin(2 * np.pi * u[:, 1]), np.sqrt(1 - u[:, 0]) * np.cos(2 * np.pi * u[:, 1]),
                  np.sqrt(u[:, 0]) * np.sin(2 * np.pi * u[:, 2]), np.sqrt(u[:, 0]) * np.cos(2 * np.pi * u[:, 2])], -1)
    x, y, z, w = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)
    s = rng.uniform(scale_range[0], scale_range[1], (n, 1, 1))
    t = rng.uniform(-extent, extent, (n, 3))
    out = np.zeros((n, 16), f32)
    m = np.concatenate([R * s, t[:, :, None]], 2)  # [n,3,4]
    out[:, :12] = m.reshape(n, 12).astype(f32)
    return out


def instanced_geometries(count=100, base_subdiv=None):
    """The C4 geometry set: geometry g is an icosphere (even g) or a lat-long sphere (odd g) whose triangle
    count grows with g mod 8 (about 128 * 2^(g mod 8) triangles)."""
    out = []
    for g in range(count):
        k = g % 8
        if g % 2 == 0:
            n_lon = max(4, int(round((64 * 2 ** k) ** 0.5)))
            out.append(sphere_mesh(2 * n_lon, n_lon // 2 + 1, radius=1.0 + 0.01 * g))
        else:
            n = max(3, int(round((64 * 2 ** k) ** 0.5)))
            v, i = grid_mesh(n, n)
            v = v.copy()
            v[:, 2] = 0.3 * np.sin(3.0 * v[:, 0] + g) * np.cos(2.0 * v[:, 1])  # wavy patch in [0,1]^2
            v[:, :2] = v[:, :2] * 2.0 - 1.0
            out.append((np.ascontiguousarray(v.astype(f32)), i))
    return out
