"""ctypes binding of libhiprt_b200.so -- the HIPRT C API (include/hiprt/hiprt.h) plus the
hiprtB200* additions (include/hiprt/hiprt_b200.h).

This is harness plumbing for tests/ and bench.py: device memory and streams come from
PyTorch, every call goes through the C-ABI exactly as a C/C++ application would make it.
There is no CPU fallback: without the shared library (or without a CUDA device for the
calls that need one) this module raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhiprt_b200.so")

HIPRT_API_VERSION = 3001
INVALID = 0xFFFFFFFF

# enums (include/hiprt/hiprt_types.h)
SUCCESS, ERR_NOT_IMPLEMENTED, ERR_INTERNAL, ERR_OOM_HOST, ERR_OOM_DEVICE, ERR_API_VERSION, ERR_INVALID_PARAM = range(7)
BUILD, UPDATE = 1, 2
FAST_BUILD, BALANCED_BUILD, HIGH_QUALITY_BUILD, CUSTOM_IMPORT = 0, 1, 2, 3
DISABLE_PAIRING = 1 << 3
PRIM_TRIANGLE_MESH, PRIM_AABB_LIST = 0, 1
INSTANCE_GEOMETRY, INSTANCE_SCENE = 0, 1
FRAME_SRT, FRAME_MATRIX = 0, 1
ANY_HIT, CLOSEST_HIT = 1, 2
FORMAT_FULL, FORMAT_COMPACT, FORMAT_PRIM_ID = 0, 1, 2  # hiprtB200HitFormat
COMPACT_DTYPE = np.dtype([("primID", "<u4"), ("t", "<f4"), ("uv", "<f4", 2)])
FORMAT_WORDS = {0: 12, 1: 4, 2: 1}
DEVICE_AMD, DEVICE_NVIDIA = 0, 1

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("minT", "<f4"), ("direction", "<f4", 3), ("maxT", "<f4")])
HIT_DTYPE = np.dtype(
    [("instanceIDs", "<u4", 4), ("primID", "<u4"), ("_pad", "<u4"), ("uv", "<f4", 2), ("normal", "<f4", 3), ("t", "<f4")]
)
BOX4_DTYPE = np.dtype(
    [("child", "<u4", 4), ("box", "<f4", (4, 2, 3)), ("parentAddr", "<u4"), ("updateCounter", "<u4"), ("childCount", "<u4"), ("_pad", "<u4")]
)
TRIPAIR_DTYPE = np.dtype([("v", "<f4", (4, 3)), ("_pad", "<u4"), ("primIndex0", "<u4"), ("primIndex1", "<u4"), ("flags", "<u4")])
SCRATCH_DTYPE = np.dtype([("box", "<f4", (2, 3)), ("child0", "<u4"), ("child1", "<u4")])
REF_DTYPE = np.dtype([("box", "<f4", (2, 3)), ("primIndex", "<u4"), ("_pad", "<u4")])
INSTANCE_NODE_DTYPE = np.dtype([("invMatrix", "<f4", (3, 4)), ("pointer", "<u8"), ("mask", "<u4"), ("bits", "<u4")])
GEOM_HEADER_DTYPE = np.dtype(
    [("boxNodes", "<u8"), ("primNodes", "<u8"), ("size", "<u8"), ("referenceCount", "<u4"), ("boxNodeCount", "<u4"),
     ("primNodeCount", "<u4"), ("geomType", "<u4"), ("rtip", "<u4"), ("_pad", "<u4")]
)
SCENE_HEADER_DTYPE = np.dtype(
    [("boxNodes", "<u8"), ("primNodes", "<u8"), ("instances", "<u8"), ("frames", "<u8"), ("size", "<u8"),
     ("referenceCount", "<u4"), ("primCount", "<u4"), ("primNodeCount", "<u4"), ("boxNodeCount", "<u4"),
     ("frameCount", "<u4"), ("rtip", "<u4")]
)
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 48 and BOX4_DTYPE.itemsize == 128
assert GEOM_HEADER_DTYPE.itemsize == 48 and SCENE_HEADER_DTYPE.itemsize == 64 and INSTANCE_NODE_DTYPE.itemsize == 64

TEMP_LAYOUT_FIELDS = ["control", "pairs", "pairScan", "primRefs", "keys0", "vals0", "keys1", "vals1", "sortScratch",
                      "sortedRefs", "scratch", "flags", "taskQueue", "total"]


class ContextCreationInput(C.Structure):
    _fields_ = [("ctxt", C.c_void_p), ("device", C.c_int), ("deviceType", C.c_int)]


class BuildOptions(C.Structure):
    _fields_ = [("buildFlags", C.c_uint32), ("batchBuildMaxPrimCount", C.c_uint32)]


class TriangleMeshPrimitive(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertexCount", C.c_uint32), ("vertexStride", C.c_uint32),
                ("triangleIndices", C.c_void_p), ("triangleCount", C.c_uint32), ("triangleStride", C.c_uint32),
                ("trianglePairIndices", C.c_void_p), ("trianglePairCount", C.c_uint32)]


class AABBListPrimitive(C.Structure):
    _fields_ = [("aabbs", C.c_void_p), ("aabbCount", C.c_uint32), ("aabbStride", C.c_uint32)]


class _PrimitiveUnion(C.Union):
    _fields_ = [("triangleMesh", TriangleMeshPrimitive), ("aabbList", AABBListPrimitive)]


class BvhNodeList(C.Structure):
    _fields_ = [("internalNodes", C.c_void_p), ("leafNodes", C.c_void_p), ("nodeCount", C.c_uint32)]


class GeometryBuildInput(C.Structure):
    _fields_ = [("type", C.c_int), ("geomType", C.c_uint32), ("primitive", _PrimitiveUnion), ("nodeList", BvhNodeList),
                ("_pad", C.c_uint8 * 48)]


class SceneBuildInput(C.Structure):
    _fields_ = [("instances", C.c_void_p), ("instanceTransformHeaders", C.c_void_p), ("instanceFrames", C.c_void_p),
                ("instanceMasks", C.c_void_p), ("nodeList", BvhNodeList), ("instanceCount", C.c_uint32),
                ("frameCount", C.c_uint32), ("frameType", C.c_int), ("_pad", C.c_uint8 * 12)]


class Instance(C.Structure):
    _fields_ = [("type", C.c_uint32), ("_pad", C.c_uint32), ("pointer", C.c_void_p)]


class FuncDataSet(C.Structure):
    _fields_ = [("intersectFuncData", C.c_void_p), ("filterFuncData", C.c_void_p)]


class FuncNameSet(C.Structure):
    _fields_ = [("intersectFuncName", C.c_char_p), ("filterFuncName", C.c_char_p)]


class GlobalStackBufferInput(C.Structure):
    _fields_ = [("type", C.c_int), ("entryType", C.c_int), ("stackSize", C.c_uint32), ("threadCount", C.c_uint32)]


class GlobalStackBuffer(C.Structure):
    _fields_ = [("stackSize", C.c_uint32), ("stackCount", C.c_uint32), ("stackData", C.c_void_p)]


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


assert C.sizeof(GeometryBuildInput) == 128, C.sizeof(GeometryBuildInput)
assert C.sizeof(SceneBuildInput) == 80, C.sizeof(SceneBuildInput)
assert C.sizeof(TriangleMeshPrimitive) == 48

# name -> (restype, argtypes); C++ references are pointers at ABI level
_vp, _u32, _sz = C.c_void_p, C.c_uint32, C.c_size_t
_P = C.POINTER
SIGNATURES = {
    "hiprtCreateContext": [_u32, _P(ContextCreationInput), _P(_vp)],
    "hiprtDestroyContext": [_vp],
    "hiprtCreateGeometry": [_vp, _P(GeometryBuildInput), BuildOptions, _P(_vp)],
    "hiprtCreateGeometries": [_vp, _u32, _P(GeometryBuildInput), BuildOptions, _P(_P(_vp))],
    "hiprtDestroyGeometry": [_vp, _vp],
    "hiprtDestroyGeometries": [_vp, _u32, _P(_vp)],
    "hiprtBuildGeometry": [_vp, C.c_int, _P(GeometryBuildInput), BuildOptions, _vp, _vp, _vp],
    "hiprtBuildGeometries": [_vp, C.c_int, _u32, _P(GeometryBuildInput), BuildOptions, _vp, _vp, _P(_vp)],
    "hiprtGetGeometryBuildTemporaryBufferSize": [_vp, _P(GeometryBuildInput), BuildOptions, _P(_sz)],
    "hiprtGetGeometriesBuildTemporaryBufferSize": [_vp, _u32, _P(GeometryBuildInput), BuildOptions, _P(_sz)],
    "hiprtCompactGeometry": [_vp, _vp, _vp, _P(_vp)],
    "hiprtCompactGeometries": [_vp, _u32, _vp, _P(_vp), _P(_P(_vp))],
    "hiprtCreateScene": [_vp, _P(SceneBuildInput), BuildOptions, _P(_vp)],
    "hiprtCreateScenes": [_vp, _u32, _P(SceneBuildInput), BuildOptions, _P(_P(_vp))],
    "hiprtDestroyScene": [_vp, _vp],
    "hiprtDestroyScenes": [_vp, _u32, _P(_vp)],
    "hiprtBuildScene": [_vp, C.c_int, _P(SceneBuildInput), BuildOptions, _vp, _vp, _vp],
    "hiprtBuildScenes": [_vp, C.c_int, _u32, _P(SceneBuildInput), BuildOptions, _vp, _vp, _P(_vp)],
    "hiprtGetSceneBuildTemporaryBufferSize": [_vp, _P(SceneBuildInput), BuildOptions, _P(_sz)],
    "hiprtGetScenesBuildTemporaryBufferSize": [_vp, _u32, _P(SceneBuildInput), BuildOptions, _P(_sz)],
    "hiprtCompactScene": [_vp, _vp, _vp, _P(_vp)],
    "hiprtCompactScenes": [_vp, _u32, _vp, _P(_vp), _P(_P(_vp))],
    "hiprtCreateFuncTable": [_vp, _u32, _u32, _P(_vp)],
    "hiprtSetFuncTable": [_vp, _vp, _u32, _u32, FuncDataSet],
    "hiprtDestroyFuncTable": [_vp, _vp],
    "hiprtCreateGlobalStackBuffer": [_vp, _P(GlobalStackBufferInput), _P(GlobalStackBuffer)],
    "hiprtDestroyGlobalStackBuffer": [_vp, GlobalStackBuffer],
    "hiprtSaveGeometry": [_vp, _vp, C.c_char_p],
    "hiprtLoadGeometry": [_vp, _P(_vp), C.c_char_p],
    "hiprtSaveScene": [_vp, _vp, C.c_char_p],
    "hiprtLoadScene": [_vp, _P(_vp), C.c_char_p],
    "hiprtExportGeometryAabb": [_vp, _vp, _P(Float3), _P(Float3)],
    "hiprtExportSceneAabb": [_vp, _vp, _P(Float3), _P(Float3)],
    "hiprtBuildTraceKernels": [_vp, _u32, _P(C.c_char_p), C.c_char_p, C.c_char_p, _u32, _P(C.c_char_p), _P(C.c_char_p),
                               _u32, _P(C.c_char_p), _u32, _u32, _P(FuncNameSet), _P(_vp), _P(_vp), C.c_bool],
    "hiprtBuildTraceKernelsFromBitcode": [_vp, _u32, _P(C.c_char_p), C.c_char_p, C.c_char_p, _sz, _u32, _u32,
                                          _P(FuncNameSet), _P(_vp), C.c_bool],
    "hiprtSetCacheDirPath": [_vp, C.c_char_p],
    "hiprtSetLogLevel": [_vp, _u32],
    # additions (include/hiprt/hiprt_b200.h)
    "hiprtB200TraceGeometry": [_vp, _vp, _vp, _u32, _vp, C.c_int, _vp],
    "hiprtB200TraceScene": [_vp, _vp, _vp, _u32, _vp, C.c_int, _u32, _vp],
    "hiprtB200TraceSceneAtTime": [_vp, _vp, _vp, _vp, C.c_float, _u32, _vp, C.c_int, _u32, _vp],
    "hiprtB200TraceGeometryHost": [_vp, _vp, _vp, _u32, _vp, C.c_int, _vp],
    "hiprtB200TraceSceneHost": [_vp, _vp, _vp, _u32, _vp, C.c_int, _u32, _vp],
    "hiprtB200TraceGeometryFormat": [_vp, _vp, _vp, _u32, _vp, C.c_int, C.c_int, _vp],
    "hiprtB200TraceGeometryFormatHost": [_vp, _vp, _vp, _u32, _vp, C.c_int, C.c_int, _vp],
    "hiprtB200TraceGeometryIndirect": [_vp, _vp, _vp, _u32, _vp, _vp, C.c_int, C.c_int, _vp],
    "hiprtB200HitFormatSize": [C.c_int],
    "hiprtB200TraceGeometryStats": [_vp, _vp, _vp, _u32, _vp, C.c_int, _vp, _vp],
    "hiprtB200GetGeometrySize": [_vp, _vp, _P(_sz)],
    "hiprtB200AllocateGeometryBlob": [_vp, _sz, _P(_vp)],
    "hiprtB200RelocateGeometry": [_vp, _vp, _vp],
    "hiprtB200SharedAlloc": [_vp, _sz, _P(_vp), _vp],
    "hiprtB200SharedFree": [_vp, _vp],
    "hiprtB200SharedOpen": [_vp, _vp, _P(_vp)],
    "hiprtB200SharedClose": [_vp, _vp],
    "hiprtB200ResolveAoFrame": [_vp, _vp, C.c_int, _vp, _vp, _vp, _u32, _u32, _vp, _vp, _u32, _vp, _vp],
    "hiprtB200WaitFlags": [_vp, _vp, _u32, _u32, _u32, _vp, _vp],
    "hiprtB200GetTempLayout": [_sz, _sz, _P(_sz), _u32],
    "hiprtB200GetHighQualityTempLayout": [_sz, _sz, C.c_int, _P(_sz), _u32],
    "hiprtB200GetDeviceHeaders": [_P(_P(C.c_char_p)), _P(_P(C.c_char_p))],
    "hiprtB200LinkTraceKernels": [C.c_int, C.c_int, C.c_char_p, C.c_char_p, _sz, _u32, _u32, _P(FuncNameSet), _vp, _P(_sz), C.c_char_p, _sz],
    "hiprtB200GetLastErrorString": [_vp],
    "hiprtB200CopyToHost": [_vp, _vp, _vp, _sz],
    "hiprtB200EnableBuildProfile": [_vp, C.c_int],
    "hiprtB200GetBuildProfile": [_vp, _P(_u32), _P(C.c_char_p), _P(C.c_float), _u32],
}
HIPRT_API_FUNCTIONS = [n for n in SIGNATURES if not n.startswith("hiprtB200")]

_lib = None


def lib():
    """Load libhiprt_b200.so; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        L.hiprtB200GetDeviceHeaders.restype = C.c_uint32
        L.hiprtB200HitFormatSize.restype = C.c_size_t
        L.hiprtB200GetLastErrorString.restype = C.c_char_p
        _lib = L
    return _lib


class HiprtError(RuntimeError):
    def __init__(self, code, what, detail=""):
        super().__init__(f"{what} failed with hiprtError {code}" + (f": {detail}" if detail else ""))
        self.code = code


def temp_layout(prim_count, paired_triangles=0):
    out = (C.c_size_t * 14)()
    lib().hiprtB200GetTempLayout(prim_count, paired_triangles, out, 14)
    return dict(zip(TEMP_LAYOUT_FIELDS, [int(v) for v in out]))


def high_quality_temp_layout(prim_count, paired_triangles=0, spatial_splits=True):
    out = (C.c_size_t * 8)()
    lib().hiprtB200GetHighQualityTempLayout(C.c_size_t(prim_count), C.c_size_t(paired_triangles), int(spatial_splits), out, 8)
    return dict(zip(["control", "state", "primRefs", "nodes", "leaves", "maxRefs", "taskQueue", "total"], [int(v) for v in out]))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_handle(stream):
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class Context:
    """hiprtContext on one CUDA device (the calling rank's GPU)."""

    def __init__(self, device=0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("hiprt_b200 needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)  # make sure the primary context exists
        inp = ContextCreationInput(None, device, DEVICE_NVIDIA)
        h = C.c_void_p()
        self._check(lib().hiprtCreateContext(HIPRT_API_VERSION, C.byref(inp), C.byref(h)), "hiprtCreateContext", None)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().hiprtDestroyContext(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: ctypes may already be gone
            pass

    def _check(self, code, what, ctx=True):
        if code != SUCCESS:
            detail = ""
            if ctx and getattr(self, "h", None):
                detail = (lib().hiprtB200GetLastErrorString(self.h) or b"").decode()
            raise HiprtError(code, what, detail)

    # ---- geometry ---------------------------------------------------------------------------
    def mesh_input(self, vertices, indices=None, pairs=None, geom_type=INVALID, vertex_stride=None, triangle_stride=12):
        """hiprtGeometryBuildInput for device tensors (float32 vertices [V,3+], uint32/int32 indices [T,3])."""
        inp = GeometryBuildInput()
        inp.type = PRIM_TRIANGLE_MESH
        inp.geomType = geom_type
        m = inp.primitive.triangleMesh
        m.vertices = vertices.data_ptr()
        m.vertexCount = vertices.shape[0]
        m.vertexStride = vertex_stride if vertex_stride is not None else vertices.stride(0) * 4
        if indices is not None:
            m.triangleIndices = indices.data_ptr()
            m.triangleCount = indices.shape[0]
            m.triangleStride = triangle_stride
        if pairs is not None:
            m.trianglePairIndices = pairs.data_ptr()
            m.trianglePairCount = pairs.shape[0]
        inp._keep = (vertices, indices, pairs)
        return inp

    def aabb_input(self, aabbs, geom_type=0):
        """aabbs: float32 device tensor [N, 6] (min xyz, max xyz) or [N, 8]."""
        inp = GeometryBuildInput()
        inp.type = PRIM_AABB_LIST
        inp.geomType = geom_type
        inp.primitive.aabbList.aabbs = aabbs.data_ptr()
        inp.primitive.aabbList.aabbCount = aabbs.shape[0]
        inp.primitive.aabbList.aabbStride = aabbs.shape[1] * 4
        inp._keep = (aabbs,)
        return inp

    def geometry_temp_size(self, inp, flags=FAST_BUILD, batch_max=0):
        sz = C.c_size_t()
        self._check(lib().hiprtGetGeometryBuildTemporaryBufferSize(self.h, C.byref(inp), BuildOptions(flags, batch_max), C.byref(sz)),
                    "hiprtGetGeometryBuildTemporaryBufferSize")
        return sz.value

    def create_geometry(self, inp, flags=FAST_BUILD, batch_max=0):
        g = C.c_void_p()
        self._check(lib().hiprtCreateGeometry(self.h, C.byref(inp), BuildOptions(flags, batch_max), C.byref(g)), "hiprtCreateGeometry")
        return g

    def build_geometry(self, inp, geom, temp, flags=FAST_BUILD, batch_max=0, op=BUILD, stream=None):
        self._check(lib().hiprtBuildGeometry(self.h, op, C.byref(inp), BuildOptions(flags, batch_max), _ptr(temp),
                                             _stream_handle(stream), geom), "hiprtBuildGeometry")

    def destroy_geometry(self, geom):
        self._check(lib().hiprtDestroyGeometry(self.h, geom), "hiprtDestroyGeometry")

    def compact_geometry(self, geom, stream=None):
        """hiprtCompactGeometry: moves the geometry into an allocation of exactly the used size; `geom` (raw handle) is freed."""
        out = C.c_void_p()
        self._check(lib().hiprtCompactGeometry(self.h, _stream_handle(stream), geom, C.byref(out)), "hiprtCompactGeometry")
        return out

    def compact_scene(self, scene, stream=None):
        out = C.c_void_p()
        self._check(lib().hiprtCompactScene(self.h, _stream_handle(stream), scene, C.byref(out)), "hiprtCompactScene")
        return out

    def save_geometry(self, geom, filename):
        self._check(lib().hiprtSaveGeometry(self.h, geom, os.fsencode(filename)), "hiprtSaveGeometry")

    def load_geometry(self, filename):
        out = C.c_void_p()
        self._check(lib().hiprtLoadGeometry(self.h, C.byref(out), os.fsencode(filename)), "hiprtLoadGeometry")
        return out

    def export_aabb(self, geom_or_scene):
        """hiprtExportGeometryAabb / hiprtExportSceneAabb -> (min xyz, max xyz) as float32 arrays."""
        lo, hi = Float3(), Float3()
        is_scene = isinstance(geom_or_scene, Scene)
        fn = lib().hiprtExportSceneAabb if is_scene else lib().hiprtExportGeometryAabb
        handle = geom_or_scene.h if hasattr(geom_or_scene, "h") else geom_or_scene
        self._check(fn(self.h, handle, C.byref(lo), C.byref(hi)), "hiprtExportSceneAabb" if is_scene else "hiprtExportGeometryAabb")
        return np.array([lo.x, lo.y, lo.z], np.float32), np.array([hi.x, hi.y, hi.z], np.float32)

    def enable_build_profile(self, on=True):
        self._check(lib().hiprtB200EnableBuildProfile(self.h, int(on)), "hiprtB200EnableBuildProfile")

    def build_profile(self):
        """[(phase name, milliseconds)] of the most recent build (needs enable_build_profile)."""
        n = C.c_uint32()
        names = (C.c_char_p * 20)()
        ms = (C.c_float * 20)()
        self._check(lib().hiprtB200GetBuildProfile(self.h, C.byref(n), names, ms, 20), "hiprtB200GetBuildProfile")
        return [(names[i].decode(), float(ms[i])) for i in range(n.value)]

    def alloc_temp(self, nbytes):
        return self.torch.empty(max(int(nbytes), 64), dtype=self.torch.uint8, device=self.device)

    def build(self, vertices, indices=None, flags=FAST_BUILD, pairs=None, geom_type=INVALID):
        """Convenience: create + build a triangle geometry. Returns a Geometry wrapper."""
        inp = self.mesh_input(vertices, indices, pairs, geom_type)
        temp = self.alloc_temp(self.geometry_temp_size(inp, flags))
        g = self.create_geometry(inp, flags)
        self.build_geometry(inp, g, temp, flags)
        return Geometry(self, g, inp, temp, flags)

    def build_geometries(self, inputs, flags=FAST_BUILD, batch_max=512, stream=None):
        """hiprtCreateGeometries + hiprtBuildGeometries for a list of build inputs: inputs with at most `batch_max`
        primitives go through the batch builder (one launch for all of them). Returns a list of Geometry wrappers
        that share one device allocation and one temporary buffer."""
        n = len(inputs)
        arr = (GeometryBuildInput * n)(*inputs)
        opts = BuildOptions(flags, batch_max)
        sz = C.c_size_t()
        self._check(lib().hiprtGetGeometriesBuildTemporaryBufferSize(self.h, n, arr, opts, C.byref(sz)), "hiprtGetGeometriesBuildTemporaryBufferSize")
        temp = self.alloc_temp(sz.value)
        handles = (C.c_void_p * n)()
        base = C.addressof(handles)
        slots = (C.c_void_p * n)(*[base + 8 * i for i in range(n)])
        self._check(lib().hiprtCreateGeometries(self.h, n, arr, opts, C.cast(slots, _P(_P(_vp)))), "hiprtCreateGeometries")
        self._check(lib().hiprtBuildGeometries(self.h, BUILD, n, arr, opts, _ptr(temp), _stream_handle(stream), handles), "hiprtBuildGeometries")
        out = [Geometry(self, C.c_void_p(handles[i]), inputs[i], temp, flags) for i in range(n)]
        for g in out:
            g.batch = (arr, handles)
        return out

    def rebuild_geometries(self, geoms, batch_max=512, stream=None):
        arr, handles = geoms[0].batch
        self._check(lib().hiprtBuildGeometries(self.h, BUILD, len(geoms), arr, BuildOptions(geoms[0].flags, batch_max), _ptr(geoms[0].temp),
                                               _stream_handle(stream), handles), "hiprtBuildGeometries")

    def destroy_geometries(self, geoms):
        n = len(geoms)
        handles = (C.c_void_p * n)(*[g.h.value for g in geoms])
        self._check(lib().hiprtDestroyGeometries(self.h, n, handles), "hiprtDestroyGeometries")
        for g in geoms:
            g.h = None

    # ---- scenes -------------------------------------------------------------------------------
    def build_scene(self, instance_handles, frames=None, frame_type=FRAME_MATRIX, masks=None, flags=FAST_BUILD,
                    instance_types=None, transform_headers=None, stream=None):
        """Create + build a scene. instance_handles: device handles (Geometry/Scene wrappers or raw pointers), one per
        instance; frames: host float32 [F,16] (hiprtFrameMatrix) or [F,12] (hiprtFrameSRT), frame k -> instance k unless
        transform_headers (uint32 [I,2] = frameIndex, frameCount) is given; masks: uint32 [I]."""
        t = self.torch
        n = len(instance_handles)
        inst = np.zeros(n, np.dtype([("type", "<u4"), ("_pad", "<u4"), ("pointer", "<u8")]))
        for k, h in enumerate(instance_handles):
            h = h.h if hasattr(h, "h") else h
            inst["pointer"][k] = int(h.value if hasattr(h, "value") else h)
        if instance_types is not None:
            inst["type"] = np.asarray(instance_types, np.uint32)
        d_inst = t.from_numpy(inst.view(np.uint8).reshape(-1)).to(self.device)
        inp = SceneBuildInput()
        inp.instances = d_inst.data_ptr()
        inp.instanceCount = n
        keep = [d_inst]
        if frames is not None:
            f = np.ascontiguousarray(frames, np.float32).reshape(-1, 16 if frame_type == FRAME_MATRIX else 12)
            d_frames = t.from_numpy(f).to(self.device)
            inp.instanceFrames = d_frames.data_ptr()
            inp.frameCount = len(f)
            keep.append(d_frames)
        inp.frameType = frame_type
        if masks is not None:
            d_masks = t.from_numpy(np.ascontiguousarray(masks, np.uint32).view(np.int32)).to(self.device)
            inp.instanceMasks = d_masks.data_ptr()
            keep.append(d_masks)
        if transform_headers is not None:
            d_hdr = t.from_numpy(np.ascontiguousarray(transform_headers, np.uint32).view(np.int32)).to(self.device)
            inp.instanceTransformHeaders = d_hdr.data_ptr()
            keep.append(d_hdr)
        opts = BuildOptions(flags, 0)
        sz = C.c_size_t()
        self._check(lib().hiprtGetSceneBuildTemporaryBufferSize(self.h, C.byref(inp), opts, C.byref(sz)), "hiprtGetSceneBuildTemporaryBufferSize")
        temp = self.alloc_temp(sz.value)
        s = C.c_void_p()
        self._check(lib().hiprtCreateScene(self.h, C.byref(inp), opts, C.byref(s)), "hiprtCreateScene")
        self._check(lib().hiprtBuildScene(self.h, BUILD, C.byref(inp), opts, _ptr(temp), _stream_handle(stream), s), "hiprtBuildScene")
        keep.append(list(instance_handles))
        return Scene(self, s, keep, inp, temp, flags)

    def rebuild_scene(self, scene, op=BUILD, stream=None):
        self._check(lib().hiprtBuildScene(self.h, op, C.byref(scene.inp), BuildOptions(scene.flags, 0), _ptr(scene.temp),
                                          _stream_handle(stream), scene.h), "hiprtBuildScene")

    # ---- tracing ------------------------------------------------------------------------------
    def trace(self, geom_or_scene, rays, hits=None, kind=CLOSEST_HIT, mask=0xFFFFFFFF, stream=None):
        """rays: uint8/float32 device tensor holding hiprtRay[n]; returns device tensor of hiprtHit[n] (as [n,12] int32)."""
        t = self.torch
        n = rays.numel() * rays.element_size() // 32
        if hits is None:
            hits = t.empty((n, 12), dtype=t.int32, device=self.device)
        handle = geom_or_scene.h if hasattr(geom_or_scene, "h") else geom_or_scene
        if isinstance(geom_or_scene, Scene):
            self._check(lib().hiprtB200TraceScene(self.h, handle, _ptr(rays), n, _ptr(hits), kind, mask, _stream_handle(stream)),
                        "hiprtB200TraceScene")
        else:
            self._check(lib().hiprtB200TraceGeometry(self.h, handle, _ptr(rays), n, _ptr(hits), kind, _stream_handle(stream)),
                        "hiprtB200TraceGeometry")
        return hits

    def trace_format(self, geom, rays, fmt, out=None, kind=CLOSEST_HIT, stream=None):
        """hiprtB200TraceGeometryFormat: device tensor of records [n, 12 | 4 | 1] int32 in hiprtB200HitFormat `fmt`."""
        t = self.torch
        n = rays.numel() * rays.element_size() // 32
        if out is None:
            out = t.empty((n, FORMAT_WORDS[fmt]), dtype=t.int32, device=self.device)
        self._check(lib().hiprtB200TraceGeometryFormat(self.h, geom.h if hasattr(geom, "h") else geom, _ptr(rays), n, _ptr(out), fmt, kind,
                                                       _stream_handle(stream)), "hiprtB200TraceGeometryFormat")
        return out

    def trace_format_host(self, geom, rays_np, fmt, out_np=None, kind=CLOSEST_HIT, stream=None):
        """hiprtB200TraceGeometryFormatHost with HOST buffers: records as HIT_DTYPE / COMPACT_DTYPE / uint32."""
        n = len(rays_np)
        if out_np is None:
            out_np = np.empty(n, {0: HIT_DTYPE, 1: COMPACT_DTYPE, 2: np.dtype("<u4")}[fmt])
        self._check(lib().hiprtB200TraceGeometryFormatHost(self.h, geom.h if hasattr(geom, "h") else geom, rays_np.ctypes.data_as(C.c_void_p), n,
                                                           out_np.ctypes.data_as(C.c_void_p), fmt, kind, _stream_handle(stream)),
                    "hiprtB200TraceGeometryFormatHost")
        return out_np

    def trace_stats(self, geom, rays, kind=CLOSEST_HIT, stream=None):
        """(box-node visits, leaf visits) of tracing `rays` (instrumented kernel; hits are discarded)."""
        t = self.torch
        n = rays.numel() * rays.element_size() // 32
        hits = t.empty((n, 12), dtype=t.int32, device=self.device)
        counts = t.zeros(2, dtype=t.int64, device=self.device)
        self._check(lib().hiprtB200TraceGeometryStats(self.h, geom.h if hasattr(geom, "h") else geom, _ptr(rays), n, _ptr(hits), kind, _ptr(counts),
                                                      _stream_handle(stream)), "hiprtB200TraceGeometryStats")
        c = counts.cpu().numpy()
        return int(c[0]), int(c[1])

    # ---- multi-GPU plumbing ---------------------------------------------------------------------
    def geometry_size(self, geom):
        sz = C.c_size_t()
        self._check(lib().hiprtB200GetGeometrySize(self.h, geom.h if hasattr(geom, "h") else geom, C.byref(sz)), "hiprtB200GetGeometrySize")
        return sz.value

    def allocate_geometry_blob(self, size):
        g = C.c_void_p()
        self._check(lib().hiprtB200AllocateGeometryBlob(self.h, size, C.byref(g)), "hiprtB200AllocateGeometryBlob")
        return g

    def relocate_geometry(self, handle, stream=None):
        self._check(lib().hiprtB200RelocateGeometry(self.h, handle, _stream_handle(stream)), "hiprtB200RelocateGeometry")

    def blob_tensor(self, handle, size):
        """uint8 device tensor aliasing a geometry blob (for torch.distributed.broadcast); keeps no ownership."""
        t = self.torch
        iface = {"shape": (int(size),), "typestr": "|u1", "data": (int(handle.value if hasattr(handle, "value") else handle), False), "version": 2}
        holder = type("DeviceBlob", (), {"__cuda_array_interface__": iface})()
        return t.as_tensor(holder, device=self.device)

    def shared_alloc(self, nbytes):
        """(device pointer, 64-byte IPC handle) of memory other ranks of the box can map."""
        p = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        self._check(lib().hiprtB200SharedAlloc(self.h, int(nbytes), C.byref(p), handle), "hiprtB200SharedAlloc")
        return p.value, bytes(handle)

    def shared_open(self, handle_bytes):
        p = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle_bytes)
        self._check(lib().hiprtB200SharedOpen(self.h, buf, C.byref(p)), "hiprtB200SharedOpen")
        return p.value

    def shared_close(self, ptr):
        self._check(lib().hiprtB200SharedClose(self.h, C.c_void_p(ptr)), "hiprtB200SharedClose")

    def shared_free(self, ptr):
        self._check(lib().hiprtB200SharedFree(self.h, C.c_void_p(ptr)), "hiprtB200SharedFree")

    def resolve_ao_frame(self, primary, primary_fmt, ao_prim_ids, first_ao_ray, pixel_index, spp, framebuffer_ptr, done_flag_ptr, done_value,
                         block_counter, stream=None):
        """hiprtB200ResolveAoFrame; framebuffer_ptr / done_flag_ptr are raw device addresses (local or peer-mapped)."""
        self._check(lib().hiprtB200ResolveAoFrame(self.h, _ptr(primary), primary_fmt, _ptr(ao_prim_ids), _ptr(first_ao_ray), _ptr(pixel_index),
                                                  int(pixel_index.numel()), spp, C.c_void_p(framebuffer_ptr),
                                                  C.c_void_p(done_flag_ptr) if done_flag_ptr else None, done_value, _ptr(block_counter),
                                                  _stream_handle(stream)), "hiprtB200ResolveAoFrame")

    def wait_flags(self, flags_ptr, count, stride_words, value, timeout=None, stream=None):
        self._check(lib().hiprtB200WaitFlags(self.h, C.c_void_p(flags_ptr), count, stride_words, value, _ptr(timeout), _stream_handle(stream)),
                    "hiprtB200WaitFlags")

    def trace_scene_at_time(self, scene, rays, times=None, time=0.0, hits=None, kind=CLOSEST_HIT, mask=0xFFFFFFFF, stream=None):
        """Scene trace with motion blur: `times` = float32 device tensor (one time per ray) or None (every ray at `time`)."""
        t = self.torch
        n = rays.numel() * rays.element_size() // 32
        if hits is None:
            hits = t.empty((n, 12), dtype=t.int32, device=self.device)
        self._check(lib().hiprtB200TraceSceneAtTime(self.h, scene.h, _ptr(rays), _ptr(times) if times is not None else None, C.c_float(time), n,
                                                    _ptr(hits), kind, mask, _stream_handle(stream)), "hiprtB200TraceSceneAtTime")
        return hits

    def trace_host(self, geom_or_scene, rays_np, hits_np=None, kind=CLOSEST_HIT, mask=0xFFFFFFFF, stream=None):
        """End-to-end call with HOST buffers (numpy structured arrays)."""
        n = len(rays_np)
        if hits_np is None:
            hits_np = np.empty(n, HIT_DTYPE)
        rp, hp = rays_np.ctypes.data_as(C.c_void_p), hits_np.ctypes.data_as(C.c_void_p)
        if isinstance(geom_or_scene, Scene):
            self._check(lib().hiprtB200TraceSceneHost(self.h, geom_or_scene.h, rp, n, hp, kind, mask, _stream_handle(stream)),
                        "hiprtB200TraceSceneHost")
        else:
            self._check(lib().hiprtB200TraceGeometryHost(self.h, geom_or_scene.h, rp, n, hp, kind, _stream_handle(stream)),
                        "hiprtB200TraceGeometryHost")
        return hits_np

    def rays_to_device(self, rays_np):
        t = self.torch
        return t.from_numpy(np.ascontiguousarray(rays_np).view(np.uint8).reshape(-1, 32)).to(self.device)

    def hits_to_host(self, hits):
        return hits.cpu().numpy().view(np.uint8).reshape(-1).view(HIT_DTYPE)

    def read_device(self, ptr, nbytes):
        """Copy raw device memory at an absolute address to a numpy uint8 array (test introspection)."""
        out = np.empty(max(int(nbytes), 1), np.uint8)
        self._check(lib().hiprtB200CopyToHost(self.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(int(ptr)), int(nbytes)),
                    "hiprtB200CopyToHost")
        return out[: int(nbytes)]


class Geometry:
    def __init__(self, ctx, handle, inp, temp, flags):
        self.ctx, self.h, self.inp, self.temp, self.flags = ctx, handle, inp, temp, flags

    def header(self):
        self.ctx.torch.cuda.synchronize()
        return self.ctx.read_device(self.h.value, 48).view(GEOM_HEADER_DTYPE)[0]

    def build_error_flag(self):
        """Non-zero when the last build that used this temporary buffer hit its watchdog (BuildControl.pad[0])."""
        self.ctx.torch.cuda.synchronize()
        return int(self.temp[44:48].cpu().numpy().view(np.uint32)[0])

    def nodes(self):
        """(box nodes, triangle-pair leaves) copied to the host."""
        h = self.header()
        boxes = self.ctx.read_device(h["boxNodes"], int(h["boxNodeCount"]) * 128).view(BOX4_DTYPE)
        if int(h["geomType"]) & 1:
            leaves = self.ctx.read_device(h["primNodes"], int(h["primNodeCount"]) * 64).view(TRIPAIR_DTYPE)
        else:
            leaves = self.ctx.read_device(h["primNodes"], int(h["primNodeCount"]) * 4).view(np.uint32)
        return boxes, leaves

    def temp_array(self, name, dtype, count):
        """Intermediate array of the last build, read from the temporary buffer."""
        n_tri = self.inp.primitive.triangleMesh.triangleCount
        prim_max = self.ctx_prim_max()
        pairing = (self.inp.primitive.triangleMesh.triangleIndices is not None and n_tri > 2
                   and self.inp.primitive.triangleMesh.trianglePairCount == 0 and not (self.flags & DISABLE_PAIRING))
        L = temp_layout(prim_max, prim_max if pairing else 0)
        self.ctx.torch.cuda.synchronize()
        dt = np.dtype(dtype)
        raw = self.temp[L[name]: L[name] + count * dt.itemsize].cpu().numpy()
        return raw.view(dt)

    def ctx_prim_max(self):
        m = self.inp.primitive.triangleMesh
        if m.trianglePairCount > 0:
            return m.trianglePairCount
        if m.triangleCount == 0 or not m.triangleIndices:
            return m.vertexCount // 3
        return m.triangleCount

    def destroy(self):
        if self.h is not None:
            self.ctx.destroy_geometry(self.h)
            self.h = None


class Scene:
    def __init__(self, ctx, handle, keep, inp=None, temp=None, flags=FAST_BUILD):
        self.ctx, self.h, self.keep, self.inp, self.temp, self.flags = ctx, handle, keep, inp, temp, flags

    def header(self):
        self.ctx.torch.cuda.synchronize()
        return self.ctx.read_device(self.h.value, 64).view(SCENE_HEADER_DTYPE)[0]

    def nodes(self):
        """(box nodes, instance leaves) copied to the host."""
        h = self.header()
        boxes = self.ctx.read_device(h["boxNodes"], int(h["boxNodeCount"]) * 128).view(BOX4_DTYPE)
        leaves = self.ctx.read_device(h["primNodes"], int(h["primNodeCount"]) * 64).view(INSTANCE_NODE_DTYPE)
        return boxes, leaves

    def destroy(self):
        if self.h is not None:
            self.ctx._check(lib().hiprtDestroyScene(self.ctx.h, self.h), "hiprtDestroyScene")
            self.h = None
