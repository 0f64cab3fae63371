"""ctypes/numpy binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg -- never by the product package hiprt_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

RAY_DTYPE = np.dtype(
    [("origin", "<f4", 3), ("minT", "<f4"), ("direction", "<f4", 3), ("maxT", "<f4")], align=False
)
HIT_DTYPE = np.dtype(
    [
        ("instanceIDs", "<u4", 4),
        ("primID", "<u4"),
        ("_pad", "<u4"),
        ("uv", "<f4", 2),
        ("normal", "<f4", 3),
        ("t", "<f4"),
    ],
    align=False,
)
BOX4_DTYPE = np.dtype(
    [
        ("child", "<u4", 4),
        ("box", "<f4", (4, 2, 3)),
        ("parentAddr", "<u4"),
        ("updateCounter", "<u4"),
        ("childCount", "<u4"),
        ("_pad", "<u4"),
    ]
)
TRIPAIR_DTYPE = np.dtype(
    [
        ("v", "<f4", (4, 3)),
        ("_pad", "<u4"),
        ("primIndex0", "<u4"),
        ("primIndex1", "<u4"),
        ("flags", "<u4"),
    ]
)
SCRATCH_DTYPE = np.dtype([("box", "<f4", (2, 3)), ("child0", "<u4"), ("child1", "<u4")])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 48
assert BOX4_DTYPE.itemsize == 128 and TRIPAIR_DTYPE.itemsize == 64 and SCRATCH_DTYPE.itemsize == 32

INVALID = 0xFFFFFFFF
FLT_MAX = np.float32(3.402823466e38)


# int fn(const hiprtRay* ray, uint32_t geomType, void* user, hiprtHit* hit): the oracle's stand-in for a hiprtFuncTable entry
CUSTOM_INTERSECT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p)
# int fn(const hiprtRay* ray, uint32_t geomType, void* user, const hiprtHit* hit): non-zero rejects the hit
HIT_FILTER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "hiprt_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.orc_pair_triangles.argtypes = [vp, u32, u32, vp]
        L.orc_pair_triangles.restype = u32
        L.orc_morton.argtypes = [vp, vp]
        L.orc_morton.restype = u32
        L.orc_geometry_create.argtypes = [vp, u32, u32, vp, u32, u32, vp, u32, i32]
        L.orc_geometry_create.restype = vp
        L.orc_geometry_destroy.argtypes = [vp]
        L.orc_geometry_rebuild_from_sorted.argtypes = [vp, vp, vp]
        L.orc_geometry_count.argtypes = [vp, i32]
        L.orc_geometry_count.restype = u32
        L.orc_geometry_read.argtypes = [vp, i32, vp]
        L.orc_lbvh_topology.argtypes = [vp, vp, u32, u32, vp]
        L.orc_lbvh_topology.restype = u32
        L.orc_geometry_trace.argtypes = [vp, vp, u32, vp, i32, vp]
        L.orc_geometry_trace_visits.argtypes = [vp, vp, u32, vp, vp]
        L.orc_trace_nodes.argtypes = [vp, vp, vp, u32, vp, i32]
        L.orc_geometry_trace_all.argtypes = [vp, vp, u32, vp, vp, vp, u32]
        L.orc_geometry_trace_all.restype = u64
        L.orc_canonicalize.argtypes = [vp, vp, u32, vp, vp, vp]
        L.orc_canonicalize.restype = u32
        L.orc_scene_create.argtypes = [vp, u32, vp, vp, i32, vp, u32]
        L.orc_scene_create.restype = vp
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_scene_count.argtypes = [vp, i32]
        L.orc_scene_count.restype = u32
        L.orc_scene_read.argtypes = [vp, i32, vp]
        L.orc_scene_trace.argtypes = [vp, vp, u32, vp, u32, i32]
        L.orc_scene_set_inv_matrices.argtypes = [vp, vp]
        L.orc_scene_build_flags.argtypes = [i32]
        L.orc_scene_build_flags.restype = None
        L.orc_scene_create_multilevel.argtypes = [vp, u32, vp, vp, i32, vp, u32, vp]
        L.orc_scene_create_multilevel.restype = vp
        L.orc_geometry_create_aabbs.argtypes = [vp, u32, u32, u32]
        L.orc_geometry_create_aabbs.restype = vp
        L.orc_geometry_create_sbvh.argtypes = [vp, u32, u32, vp, u32, u32, vp, u32, i32]
        L.orc_geometry_create_sbvh.restype = vp
        L.orc_geometry_create_aabbs_sbvh.argtypes = [vp, u32, u32, u32, i32]
        L.orc_geometry_create_aabbs_sbvh.restype = vp
        L.orc_trace_general.argtypes = [vp, i32, vp, u32, u32, i32, CUSTOM_INTERSECT_FN, HIT_FILTER_FN, vp, u32, vp, vp]
        L.orc_geometry_set_geom_type.argtypes = [vp, u32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_rays(origins, directions, min_t=0.0, max_t=FLT_MAX):
    origins = np.asarray(origins, np.float32).reshape(-1, 3)
    directions = np.asarray(directions, np.float32).reshape(-1, 3)
    r = np.zeros(len(origins), RAY_DTYPE)
    r["origin"], r["direction"] = origins, directions
    r["minT"], r["maxT"] = min_t, max_t
    return r


def pair_triangles(indices):
    idx = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
    out = np.zeros((len(idx), 2), np.uint32)
    n = lib().orc_pair_triangles(_p(idx), len(idx), 12, _p(out))
    return out[:n].copy()


def morton(pos, ext):
    pos = np.ascontiguousarray(pos, np.float32)
    ext = np.ascontiguousarray(ext, np.float32)
    return int(lib().orc_morton(_p(pos), _p(ext)))


class Geometry:
    """LBVH-built triangle geometry of the oracle."""

    def __init__(self, vertices, indices=None, pairs=None, allow_pairing=True, high_quality=False, spatial_splits=True):
        """high_quality: the SBVH restatement (PreferHighQualityBuild) instead of the LBVH; `sbvh_stats` then holds
        (levels, references incl. duplicates, spatial splits, splits by arrival, inconsistencies)."""
        self.vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        self.indices = None if indices is None else np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        self.pairs_in = None if pairs is None else np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
        ntri = 0 if self.indices is None else len(self.indices)
        self.high_quality = high_quality
        if high_quality:
            self.h = lib().orc_geometry_create_sbvh(
                _p(self.vertices), len(self.vertices), 12, _p(self.indices), ntri, 12,
                _p(self.pairs_in), 0 if self.pairs_in is None else len(self.pairs_in), int(allow_pairing) | (2 if spatial_splits else 0))
            return
        self.h = lib().orc_geometry_create(
            _p(self.vertices), len(self.vertices), 12, _p(self.indices), ntri, 12,
            _p(self.pairs_in), 0 if self.pairs_in is None else len(self.pairs_in), int(allow_pairing))

    sbvh_stats = property(lambda s: s._read(3, np.uint32, 5) if s.prim_count > 1 else np.zeros(5, np.uint32))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_geometry_destroy(self.h)
                self.h = None
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass

    def set_geom_type(self, geom_type):
        """User geometry type (hiprtGeometryBuildInput::geomType): makes the hit filter of trace_general apply to this geometry."""
        lib().orc_geometry_set_geom_type(self.h, geom_type)

    def _count(self, what):
        return int(lib().orc_geometry_count(self.h, what))

    prim_count = property(lambda s: s._count(0))
    box_count = property(lambda s: s._count(1))
    leaf_count = property(lambda s: s._count(2))
    root = property(lambda s: s._count(3))

    def _read(self, what, dtype, n):
        out = np.zeros(n, dtype)
        lib().orc_geometry_read(self.h, what, _p(out))
        return out

    pairs = property(lambda s: s._read(0, np.uint32, 2 * s.prim_count).reshape(-1, 2))
    prim_boxes = property(lambda s: s._read(1, np.float32, 6 * s.prim_count).reshape(-1, 2, 3))
    keys = property(lambda s: s._read(2, np.uint32, s.prim_count))
    sorted_keys = property(lambda s: s._read(3, np.uint32, s.prim_count))
    sorted_vals = property(lambda s: s._read(4, np.uint32, s.prim_count))
    scratch = property(lambda s: s._read(5, SCRATCH_DTYPE, s._count(4)))
    leaf_record_boxes = property(lambda s: s._read(11, np.float32, 6 * s._count(5)).reshape(-1, 2, 3))
    leaf_record_prims = property(lambda s: s._read(12, np.uint32, s._count(5)))
    box_nodes = property(lambda s: s._read(6, BOX4_DTYPE, s.box_count))
    leaves = property(lambda s: s._read(7, TRIPAIR_DTYPE, s.leaf_count))
    scene_box = property(lambda s: s._read(8, np.float32, 6).reshape(2, 3))
    prim_nodes = property(lambda s: s._read(9, TRIPAIR_DTYPE, s.prim_count))

    def rebuild_from_sorted(self, sorted_keys, sorted_vals):
        k = np.ascontiguousarray(sorted_keys, np.uint32)
        v = np.ascontiguousarray(sorted_vals, np.uint32)
        assert len(k) == len(v) == self.prim_count
        lib().orc_geometry_rebuild_from_sorted(self.h, _p(k), _p(v))

    def trace(self, rays, tie_mode=1, want_visits=False):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE)
        visits = np.zeros(2, np.uint32)
        lib().orc_geometry_trace(self.h, _p(rays), len(rays), _p(hits), tie_mode, _p(visits))
        return (hits, visits) if want_visits else hits

    def trace_visits(self, rays):
        """Per-ray (box visits, leaf visits) of the closest-hit traversal."""
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        bv = np.zeros(len(rays), np.uint32)
        lv = np.zeros(len(rays), np.uint32)
        lib().orc_geometry_trace_visits(self.h, _p(rays), len(rays), _p(bv), _p(lv))
        return bv, lv

    def trace_all(self, rays, max_per_ray=16):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        n = len(rays)
        counts = np.zeros(n, np.uint32)
        prims = np.full((n, max_per_ray), INVALID, np.uint32)
        ts = np.zeros((n, max_per_ray), np.float32)
        lib().orc_geometry_trace_all(self.h, _p(rays), n, _p(counts), _p(prims), _p(ts), max_per_ray)
        return counts, prims, ts


def motion_pose(srt_frames, time, ray=None, frame_type=0):
    """Key-frame interpolation of the oracle: (3x4 object->world matrix of the pose at `time`, world ray carried into the
    instance or None). srt_frames: float32 [n, 12] hiprtFrameSRT rows (frame_type 0) or [n, 16] hiprtFrameMatrix rows (1)."""
    f = np.ascontiguousarray(srt_frames, np.float32).reshape(-1, 16 if frame_type == 1 else 12)
    m = np.zeros(12, np.float32)
    rin = rout = None
    if ray is not None:
        rin = np.ascontiguousarray(ray, RAY_DTYPE).reshape(1)
        rout = np.zeros(1, RAY_DTYPE)
    L = lib()
    L.orc_motion_pose_typed.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_motion_pose_typed(_p(f), len(f), frame_type, C.c_float(time), _p(m), _p(rin), _p(rout))
    return m.reshape(3, 4), (None if rout is None else rout[0])


def lbvh_topology(sorted_keys, sorted_boxes, leaf_type=0):
    k = np.ascontiguousarray(sorted_keys, np.uint32)
    b = np.ascontiguousarray(sorted_boxes, np.float32).reshape(-1, 2, 3)
    out = np.zeros(max(len(k) - 1, 0), SCRATCH_DTYPE)
    root = lib().orc_lbvh_topology(_p(k), _p(b), len(k), leaf_type, _p(out))
    return out, int(root)


def trace_nodes(box_nodes, tri_nodes, rays, tie_mode=1):
    bn = np.ascontiguousarray(box_nodes)
    tn = np.ascontiguousarray(tri_nodes)
    rays = np.ascontiguousarray(rays, RAY_DTYPE)
    hits = np.zeros(len(rays), HIT_DTYPE)
    lib().orc_trace_nodes(_p(bn), _p(tn), _p(rays), len(rays), _p(hits), tie_mode)
    return hits


def canonicalize(box_nodes, leaves):
    """Renumber a 4-wide structure breadth-first (see orc_canonicalize)."""
    bn = np.ascontiguousarray(box_nodes, BOX4_DTYPE)
    lv = np.ascontiguousarray(leaves)
    out_b = np.zeros(len(bn), BOX4_DTYPE)
    out_l = np.zeros(len(lv), lv.dtype)
    nleaf = C.c_uint32(0)
    nb = lib().orc_canonicalize(_p(bn), _p(lv), lv.dtype.itemsize, _p(out_b), _p(out_l), C.byref(nleaf))
    return out_b[:nb], out_l[: nleaf.value]


INSTANCE_NODE_DTYPE = np.dtype([("invMatrix", "<f4", (3, 4)), ("pointer", "<u8"), ("mask", "<u4"), ("bits", "<u4")])


class AabbGeometry:
    """Custom primitives: an AABB list (float32 [n, 6] min xyz, max xyz) built by the oracle's LBVH pipeline with custom leaves.
    Hits come from the intersection function passed to trace_general."""

    def __init__(self, aabbs, geom_type=0, high_quality=False, spatial_splits=True):
        self.aabbs = np.ascontiguousarray(aabbs, np.float32).reshape(-1, 6)
        self.geom_type = geom_type
        if high_quality:
            self.h = lib().orc_geometry_create_aabbs_sbvh(_p(self.aabbs), len(self.aabbs), 24, geom_type, 2 if spatial_splits else 0)
        else:
            self.h = lib().orc_geometry_create_aabbs(_p(self.aabbs), len(self.aabbs), 24, geom_type)

    def _read(self, what, dtype, n):
        out = np.zeros(n, dtype)
        lib().orc_geometry_read(self.h, what, _p(out))
        return out

    box_nodes = property(lambda s: s._read(6, BOX4_DTYPE, int(lib().orc_geometry_count(s.h, 1))))
    leaves = property(lambda s: s._read(10, np.uint32, int(lib().orc_geometry_count(s.h, 2))))  # CustomNode: the primitive index, in leaf order

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_geometry_destroy(self.h)
                self.h = None
        except Exception:
            pass


def trace_general(obj, rays, intersect=None, ray_mask=0xFFFFFFFF, tie_mode=1, max_per_ray=0, hit_filter=None):
    """Closest hit (max_per_ray == 0: HIT_DTYPE [n]) or every hit along each ray in visiting order (max_per_ray > 0:
    (counts [n], HIT_DTYPE [n, max_per_ray])) through a Geometry, AabbGeometry or (multi-level) Scene. intersect(ray, geom_type,
    prim_id) -> None or t is the custom-primitive intersection function: `ray` is a RAY_DTYPE record in the primitive's object
    space. hit_filter(ray, geom_type, hit) -> True rejects the candidate hit (a HIT_DTYPE record); it is consulted for the
    geometries that have a user geometry type (Geometry.set_geom_type, AabbGeometry)."""
    rays = np.ascontiguousarray(rays, RAY_DTYPE)
    n = len(rays)

    def filter_callback(ray_ptr, geom_type, _user, hit_ptr):
        ray = np.ctypeslib.as_array(C.cast(ray_ptr, C.POINTER(C.c_uint8)), (32,)).view(RAY_DTYPE)[0]
        hit = np.ctypeslib.as_array(C.cast(hit_ptr, C.POINTER(C.c_uint8)), (48,)).view(HIT_DTYPE)[0]
        return int(bool(hit_filter(ray, geom_type, hit)))

    flt = HIT_FILTER_FN(filter_callback) if hit_filter is not None else C.cast(None, HIT_FILTER_FN)

    def callback(ray_ptr, geom_type, _user, hit_ptr):
        ray = np.ctypeslib.as_array(C.cast(ray_ptr, C.POINTER(C.c_uint8)), (32,)).view(RAY_DTYPE)[0]
        hit = np.ctypeslib.as_array(C.cast(hit_ptr, C.POINTER(C.c_uint8)), (48,)).view(HIT_DTYPE)
        t = intersect(ray, geom_type, int(hit["primID"][0]))
        if t is None:
            return 0
        hit["t"][0] = t
        return 1

    fn = CUSTOM_INTERSECT_FN(callback) if intersect is not None else C.cast(None, CUSTOM_INTERSECT_FN)
    is_scene = int(isinstance(obj, Scene))
    if max_per_ray == 0:
        hits = np.zeros(n, HIT_DTYPE)
        lib().orc_trace_general(obj.h, is_scene, _p(rays), n, ray_mask, tie_mode, fn, flt, None, 0, None, _p(hits))
        return hits
    counts = np.zeros(n, np.uint32)
    hits = np.zeros((n, max_per_ray), HIT_DTYPE)
    hits["primID"] = INVALID
    lib().orc_trace_general(obj.h, is_scene, _p(rays), n, ray_mask, tie_mode, fn, flt, None, max_per_ray, _p(counts), _p(hits))
    return counts, hits


class Scene:
    """Two-level scene of the oracle: instances of oracle Geometry objects, one frame per instance
    (frame_type 0: hiprtFrameSRT rows of 12 floats, 1: hiprtFrameMatrix rows of 16 floats)."""

    def __init__(self, geometries, geometry_of_instance, frames, frame_type=1, masks=None, high_quality=False, spatial_splits=True):
        """high_quality: the top level is built by the SBVH restatement over instance boxes (an instance may then sit in several
        instance nodes) instead of the LBVH."""
        self.geometries = list(geometries)
        self.geometry_of_instance = np.ascontiguousarray(geometry_of_instance, np.uint32)
        n = len(self.geometry_of_instance)
        self.frames = None if frames is None else np.ascontiguousarray(frames, np.float32).reshape(n, 16 if frame_type == 1 else 12)
        self.masks = None if masks is None else np.ascontiguousarray(masks, np.uint32)
        handles = (C.c_void_p * len(self.geometries))(*[g.h for g in self.geometries])
        # an entry of `geometries` may itself be a Scene: it is then instanced as hiprtInstanceTypeScene (multi-level instancing)
        nested = np.array([isinstance(g, Scene) for g in self.geometries], np.uint32)
        lib().orc_scene_build_flags((1 | (2 if spatial_splits else 0)) if high_quality else 0)
        self.h = lib().orc_scene_create_multilevel(handles, len(self.geometries), _p(self.geometry_of_instance), _p(self.frames),
                                                   frame_type, _p(self.masks), n, _p(nested))
        lib().orc_scene_build_flags(0)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    count = property(lambda s: int(lib().orc_scene_count(s.h, 0)))
    box_count = property(lambda s: int(lib().orc_scene_count(s.h, 1)))
    leaf_count = property(lambda s: int(lib().orc_scene_count(s.h, 2)))

    def _read(self, what, dtype, n):
        out = np.zeros(n, dtype)
        lib().orc_scene_read(self.h, what, _p(out))
        return out

    prim_boxes = property(lambda s: s._read(0, np.float32, 6 * s.count).reshape(-1, 2, 3))
    keys = property(lambda s: s._read(1, np.uint32, s.count))
    box_nodes = property(lambda s: s._read(2, BOX4_DTYPE, s.box_count))
    leaves = property(lambda s: s._read(3, INSTANCE_NODE_DTYPE, s.leaf_count))
    sorted_keys = property(lambda s: s._read(4, np.uint32, s.count))
    sorted_vals = property(lambda s: s._read(5, np.uint32, s.count))

    def adopt_transforms(self, instance_leaves):
        """Use the world->object matrices of externally built instance leaves (INSTANCE_NODE_DTYPE, any order)."""
        leaves = np.asarray(instance_leaves).view(INSTANCE_NODE_DTYPE)
        inv = np.zeros((self.count, 3, 4), np.float32)
        inv[leaves["bits"] & 0xFFFFFF] = leaves["invMatrix"]
        lib().orc_scene_set_inv_matrices(self.h, _p(inv))

    def trace(self, rays, ray_mask=0xFFFFFFFF, tie_mode=1):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE)
        lib().orc_scene_trace(self.h, _p(rays), len(rays), _p(hits), ray_mask, tie_mode)
        return hits
