"""Golden vectors of the REFERENCE's own arithmetic, run on the CPU in this container: oracle/_ref/libref_host_math.so is
compiled from the reference's HIPRT_HOST_DEVICE headers where they lie (oracle/ref/ref_host_math.cpp, oracle/ref/Makefile), fed the
seeded inputs of ref_math_inputs() and its outputs are stored in tests/golden/ref_math.npz, so that the oracle stays pinned to
them where /root/reference does not exist (tests/test_oracle_vs_reference_math.py). Usage: python tests/golden/make_ref_math.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_host_math.so")

RAY = np.dtype([("origin", "<f4", 3), ("minT", "<f4"), ("direction", "<f4", 3), ("maxT", "<f4")])


def ref_math_inputs(seed=2024):
    """Seeded inputs, small enough to commit: generic values plus the edge cases that decide parity (cell boundaries of the Morton
    grid, flat extents, rays through edges and vertices, axis-parallel rays against flat boxes, half-turn and sheared frames)."""
    rng = np.random.default_rng(seed)
    f = np.float32
    d = {}
    # Morton codes: three extents (cubic, anisotropic, flat in z), positions on and next to cell boundaries
    pos = rng.uniform(0, 1, (3000, 3)).astype(f)
    k = rng.integers(0, 1024, (1000, 3)).astype(f) / f(1024)
    pos = np.concatenate([pos, k, np.nextafter(k, f(0)), np.nextafter(k, f(2)), np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5]], f)])
    d["morton_pos"] = pos
    d["morton_ext"] = np.array([[1, 1, 1], [3.0, 0.25, 1.5], [2.0, 1.0, 0.0], [1e-3, 5.0, 5.0]], f)
    # triangles and rays: random, plus rays aimed exactly at vertices, edge midpoints and the triangle plane from inside it
    n = 4000
    v = rng.uniform(-1, 1, (n, 3, 3)).astype(f)
    o = rng.uniform(-3, 3, (n, 3)).astype(f)
    w = rng.dirichlet([1, 1, 1], n).astype(f)
    w[: n // 4] = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0.5, 0.5, 0]], f)[rng.integers(0, 4, n // 4)]
    target = (v * w[:, :, None]).sum(1).astype(f)
    target[n // 2:] += rng.normal(0, 0.7, (n - n // 2, 3)).astype(f)  # half of the rays may miss
    rays = np.zeros(n, RAY)
    rays["origin"], rays["direction"] = o, target - o
    rays["maxT"] = np.where(rng.uniform(size=n) < 0.2, f(0.9), f(3.4e38))
    rays["minT"] = np.where(rng.uniform(size=n) < 0.1, f(0.5), f(0))
    d["tri_v"], d["tri_rays"] = v.reshape(n, 9), rays.view(np.uint8).reshape(n, 32)
    d["tri_flags"] = rng.integers(0, 64, n).astype(np.uint32)
    # slab test: random boxes, flat boxes, axis-parallel directions (infinite reciprocals), origins on box planes
    m = 4000
    lo = rng.uniform(-1, 0, (m, 3)).astype(f)
    hi = (lo + rng.uniform(0, 1.5, (m, 3))).astype(f)
    hi[: m // 8, 2] = lo[: m // 8, 2]
    bo = rng.uniform(-2, 2, (m, 3)).astype(f)
    bd = rng.normal(size=(m, 3)).astype(f)
    bd[m // 8: m // 4, rng.integers(0, 3)] = 0
    bo[m // 4: m // 4 + 200, 0] = lo[m // 4: m // 4 + 200, 0]
    with np.errstate(divide="ignore"):
        inv = (f(1) / bd).astype(f)
    d["slab_boxes"], d["slab_o"], d["slab_inv"] = np.concatenate([lo, hi], 1), bo, inv
    d["slab_maxt"] = np.where(rng.uniform(size=m) < 0.3, f(1.0), f(3.4e38)).astype(f)
    # frames: SRT (axis-angle incl. angle 0 and pi) and matrices (rotations, shears, reflections, the half turn, the identity)
    ns = 600
    axis = rng.normal(size=(ns, 3)).astype(f)
    angle = rng.uniform(-3.2, 3.2, ns).astype(f)
    angle[:20] = 0
    angle[20:40] = np.pi
    srt = np.zeros((ns, 12), f)
    srt[:, :3], srt[:, 3] = axis, angle
    srt[:, 4:7] = rng.uniform(0.3, 2.5, (ns, 3))
    srt[:, 7:10] = rng.uniform(-5, 5, (ns, 3))
    srt[:, 10] = rng.uniform(0, 1, ns)
    d["frames_srt"] = srt
    mats = np.zeros((ns, 16), f)
    a = rng.normal(size=(ns, 3, 3))
    a[:50] = np.linalg.qr(a[:50])[0]
    mats[:, :12] = np.concatenate([a, rng.uniform(-5, 5, (ns, 3, 1))], 2).reshape(ns, 12).astype(f)
    mats[0, :12] = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], f).reshape(-1)
    mats[1, :12] = np.array([[-1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 1, 0]], f).reshape(-1)
    mats[2, :12] = np.array([[1, 1, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], f).reshape(-1)
    mats[:, 12] = rng.uniform(0, 1, ns)
    d["frames_matrix"] = mats
    # triangle pairs: every way two triangles can share an edge (3 edges x 3 rotations of the first x 6 vertex orders of the second),
    # 12 random quads each, plus unpaired triangles (x == y)
    import itertools
    pv, pt, pl = [], [], []
    for edge, rot0, perm1 in itertools.product(range(3), range(3), itertools.permutations(range(3))):
        for _ in range(12):
            a = rng.uniform(-1, 1, (4, 3)).astype(f)  # 0,1,2 = first triangle, 3 = the vertex across the shared edge
            base = len(pv)
            pv.extend(a)
            first = np.roll(np.array([0, 1, 2]), rot0)
            shared = [(0, 1), (1, 2), (2, 0)][edge]
            second = np.array([shared[0], shared[1], 3])[list(perm1)]
            pt.append(base + first)
            pt.append(base + second)
            pl.append([len(pt) - 2, len(pt) - 1])
    for k in range(0, 60, 2):
        pl.append([k, k])
    d["pair_vertices"], d["pair_triangles"], d["pair_list"] = np.array(pv, f), np.array(pt, np.uint32), np.array(pl, np.uint32)
    # motion: key-frame sets (times ascending from 0), rays and times inside, on and beyond the key times
    sets = []
    for j in range(12):
        kind = j % 2
        cnt = int(rng.integers(1, 5))
        keys = (srt if kind == 0 else mats)[rng.choice(np.arange(40, ns), cnt, replace=False)].copy()
        times = np.sort(rng.uniform(0, 1, cnt)).astype(f)
        times[0] = 0
        keys[:, 10 if kind == 0 else 12] = times
        sets.append((kind, keys))
    half_turn = np.stack([mats[0], mats[1]]).copy()
    half_turn[:, 12] = [0.0, 1.0]
    sets.append((1, half_turn))
    d["motion_count"] = np.array(len(sets))
    for j, (kind, keys) in enumerate(sets):
        d[f"motion_kind_{j}"], d[f"motion_keys_{j}"] = np.array(kind), keys
    blo = rng.uniform(-2, 1, (200, 3)).astype(f)
    d["motion_boxes"] = np.concatenate([blo, blo + rng.uniform(0, 2, (200, 3)).astype(f)], 1)
    nm = 300
    mr = np.zeros(nm, RAY)
    mr["origin"], mr["direction"] = rng.uniform(-4, 4, (nm, 3)), rng.normal(size=(nm, 3))
    mr["maxT"], mr["minT"] = 100.0, 0.25
    d["motion_rays"] = mr.view(np.uint8).reshape(nm, 32)
    t = rng.uniform(-0.2, 1.2, nm).astype(f)
    t[:10] = [0, 1, 0.5, 0.25, 0.75, 1.5, -1, 0.999, 0.001, 0.35]
    d["motion_times"] = t
    return d


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def run(L, prefix, d):
    """Outputs of a library that exports <prefix>morton / triangle / slab / frames (and ref_motion_ray for the reference)."""
    out = {}
    pos = np.ascontiguousarray(d["morton_pos"])
    for e, ext in enumerate(d["morton_ext"]):
        codes = np.zeros(len(pos), np.uint32)
        getattr(L, prefix + "morton")(len(pos), p(pos), p(np.ascontiguousarray(ext)), p(codes))
        out[f"morton_{e}"] = codes
    n = len(d["tri_v"])
    hit, t, uv = np.zeros(n, np.uint8), np.zeros(n, np.float32), np.zeros((n, 2), np.float32)
    getattr(L, prefix + "triangle")(n, p(d["tri_v"]), p(d["tri_rays"]), p(d["tri_flags"]), p(hit), p(t), p(uv))
    out["tri_hit"], out["tri_t"], out["tri_uv"] = hit, np.where(hit, t, 0).astype(np.float32), np.where(hit[:, None], uv, 0).astype(np.float32)
    m = len(d["slab_boxes"])
    t01 = np.zeros((m, 2), np.float32)
    getattr(L, prefix + "slab")(m, p(d["slab_boxes"]), p(d["slab_o"]), p(d["slab_inv"]), p(d["slab_maxt"]), p(t01))
    out["slab_t01"] = t01
    pn = len(d["pair_list"])
    nodes, boxes = np.zeros((pn, 64), np.uint8), np.zeros((pn, 6), np.float32)
    getattr(L, prefix + "pair_nodes")(p(d["pair_vertices"]), len(d["pair_vertices"]), p(d["pair_triangles"]), len(d["pair_triangles"]),
                                      p(d["pair_list"]), pn, p(nodes), p(boxes))
    nodes[:, 48:52] = 0  # the padding word between the vertices and the primitive indices is not written
    out["pair_nodes"], out["pair_boxes"] = nodes, boxes
    for kind, name in ((0, "frames_srt"), (1, "frames_matrix")):
        fr = d[name]
        srt, inv = np.zeros((len(fr), 13), np.float32), np.zeros((len(fr), 12), np.float32)
        getattr(L, prefix + "frames")(len(fr), p(fr), kind, p(srt), p(inv))
        out[name + "_decomposed"], out[name + "_inverse"] = srt, inv
    return out


def run_motion_bounds(L, prefix, d):
    out = {}
    boxes = d["motion_boxes"]
    for j in range(int(d["motion_count"])):
        keys = np.ascontiguousarray(d[f"motion_keys_{j}"])
        res = np.zeros_like(boxes)
        getattr(L, prefix + "motion_bounds")(len(keys), p(keys), int(d[f"motion_kind_{j}"]), len(boxes), p(boxes), p(res))
        out[f"motion_bounds_{j}"] = res
    return out


POINT_TIMES = (0.0, 0.3, 0.5, 0.9, 1.0)


def run_reference_points(L, d):
    """Also stores, per key-frame set, the reference's decomposed key frames in the 64-byte layout of the scene storage
    (rotation 4, scale 3, shear 3, translation 3, time, 2 pad): the input of the product's traversal-side code."""
    out = {}
    pts = np.ascontiguousarray(d["motion_rays"].view(RAY).reshape(-1)["direction"][:100])
    for j in range(int(d["motion_count"])):
        keys, kind = np.ascontiguousarray(d[f"motion_keys_{j}"]), int(d[f"motion_kind_{j}"])
        srt, inv = np.zeros((len(keys), 13), np.float32), np.zeros((len(keys), 12), np.float32)
        L.ref_frames(len(keys), p(keys), kind, p(srt), p(inv))
        stored = np.zeros((len(keys), 16), np.float32)
        stored[:, :13] = srt
        stored[:, 13] = keys[:, 12 if kind == 1 else 10]
        out[f"motion_stored_{j}"] = stored
        res = np.zeros((len(POINT_TIMES), len(pts), 12), np.float32)
        for k, time in enumerate(POINT_TIMES):
            L.ref_motion_points(len(keys), p(keys), kind, C.c_float(time), len(pts), p(pts), p(res[k]))
        out[f"motion_points_{j}"] = res
    return out


def run_reference_motion(L, d):
    out = {}
    rays, times = d["motion_rays"], d["motion_times"]
    for j in range(int(d["motion_count"])):
        keys = np.ascontiguousarray(d[f"motion_keys_{j}"])
        res = np.zeros_like(rays)
        L.ref_motion_ray(len(keys), p(keys), int(d[f"motion_kind_{j}"]), len(rays), p(times), p(rays), p(res))
        out[f"motion_out_{j}"] = res
    return out


def main():
    L = C.CDLL(REF_LIB)
    d = ref_math_inputs()
    out = run(L, "ref_", d)
    out.update(run_reference_motion(L, d))
    out.update(run_motion_bounds(L, "ref_", d))
    out.update(run_reference_points(L, d))
    np.savez_compressed(os.path.join(HERE, "ref_math.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
