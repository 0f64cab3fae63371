"""ctypes binding of oracle/_ref/libref_harness.so: the UNMODIFIED reference (GPUOpen HIPRT
v3.1.0) built by oracle/ref/Makefile, driven on the GPU box.

TEST / BENCHMARK INFRASTRUCTURE ONLY (parity pinning and the HIPRT-on-Orochi timing baseline);
never imported by the product package.
"""
import ctypes as C
import os

import numpy as np

from .binding import BOX4_DTYPE, HIT_DTYPE, RAY_DTYPE, TRIPAIR_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
LIB = os.path.join(REF_DIR, "libref_harness.so")

GEOM_HEADER_DTYPE = np.dtype(
    [("boxNodes", "<u8"), ("primNodes", "<u8"), ("size", "<u8"), ("referenceCount", "<u4"), ("boxNodeCount", "<u4"),
     ("primNodeCount", "<u4"), ("geomType", "<u4"), ("rtip", "<u4"), ("_pad", "<u4")]
)


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in
               ("libref_harness.so", "libhiprt_ref.so", "hiprt03001_nv.fatbin", "oro_compiled_kernels.fatbin", "ref_trace.fatbin"))


_lib = None


def lib(device=0):
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built (make -C oracle/ref, needs /root/reference)")
        L = C.CDLL(LIB)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
        L.ref_last_error.restype = C.c_char_p
        L.ref_init.argtypes = [i32]
        L.ref_build_geometry.argtypes = [vp, u32, vp, u32, u32, i32, C.POINTER(vp), C.POINTER(C.c_double)]
        L.ref_geometry_header.argtypes = [vp, vp]
        L.ref_geometry_nodes.argtypes = [vp, vp, u32, vp, u32]
        L.ref_geometry_device_handle.argtypes = [vp]
        L.ref_geometry_device_handle.restype = vp
        L.ref_destroy_geometry.argtypes = [vp]
        L.ref_trace_geometry.argtypes = [vp, vp, u32, vp, i32, i32, C.POINTER(C.c_double)]
        L.ref_trace_scene.argtypes = [vp, vp, u32, vp, i32, i32, C.POINTER(C.c_double)]
        L.ref_build_scene.argtypes = [vp, vp, vp, u32, u32, i32, C.POINTER(vp), C.POINTER(C.c_double)]
        L.ref_scene_device_handle.argtypes = [vp]
        L.ref_scene_device_handle.restype = vp
        L.ref_scene_dump.argtypes = [vp, vp, vp, u32, vp, u32]
        L.ref_destroy_scene.argtypes = [vp]
        _check(L, L.ref_init(device), "ref_init")
        _lib = L
    return _lib


def _check(L, rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {L.ref_last_error().decode()}")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefGeometry:
    """Triangle geometry built by the reference's own builders on the GPU."""

    def __init__(self, vertices, indices=None, build_flags=0, reps=1):
        L = lib()
        self.vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        self.indices = None if indices is None else np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        h, ms = C.c_void_p(), C.c_double()
        _check(L, L.ref_build_geometry(_p(self.vertices), len(self.vertices), _p(self.indices),
                                       0 if self.indices is None else len(self.indices), build_flags, reps, C.byref(h), C.byref(ms)),
               "ref_build_geometry")
        self.h, self.build_ms = h, ms.value

    @property
    def device_handle(self):
        return lib().ref_geometry_device_handle(self.h)

    def header(self):
        out = np.zeros(1, GEOM_HEADER_DTYPE)
        _check(lib(), lib().ref_geometry_header(self.h, _p(out)), "ref_geometry_header")
        return out[0]

    def nodes(self):
        h = self.header()
        boxes = np.zeros(int(h["boxNodeCount"]), BOX4_DTYPE)
        leaves = np.zeros(int(h["primNodeCount"]), TRIPAIR_DTYPE)
        _check(lib(), lib().ref_geometry_nodes(self.h, _p(boxes), len(boxes), _p(leaves), len(leaves)), "ref_geometry_nodes")
        return boxes, leaves

    def trace(self, rays, kind=2, reps=1):
        return trace_geometry(self.device_handle, rays, kind, reps)

    def destroy(self):
        if self.h:
            lib().ref_destroy_geometry(self.h)
            self.h = None


def trace_geometry(device_handle, rays, kind=2, reps=1):
    """Trace with the reference's traversal kernel through ANY device geometry handle (the
    reference's own, or one built by libhiprt_b200: the formats are byte-compatible)."""
    rays = np.ascontiguousarray(rays, RAY_DTYPE)
    hits = np.zeros(len(rays), HIT_DTYPE)
    ms = C.c_double()
    handle = device_handle.value if hasattr(device_handle, "value") else device_handle
    _check(lib(), lib().ref_trace_geometry(C.c_void_p(handle), _p(rays), len(rays), _p(hits), kind, reps, C.byref(ms)), "ref_trace_geometry")
    return hits, ms.value


def trace_scene(device_handle, rays, kind=2, reps=1):
    rays = np.ascontiguousarray(rays, RAY_DTYPE)
    hits = np.zeros(len(rays), HIT_DTYPE)
    ms = C.c_double()
    handle = device_handle.value if hasattr(device_handle, "value") else device_handle
    _check(lib(), lib().ref_trace_scene(C.c_void_p(handle), _p(rays), len(rays), _p(hits), kind, reps, C.byref(ms)), "ref_trace_scene")
    return hits, ms.value


class RefScene:
    def __init__(self, geometry_device_handles, frames, masks=None, build_flags=0, reps=1):
        L = lib()
        n = len(geometry_device_handles)
        handles = (C.c_void_p * n)(*[int(h.value if hasattr(h, "value") else h) for h in geometry_device_handles])
        self.frames = np.ascontiguousarray(frames, np.float32).reshape(n, 16)
        self.masks = None if masks is None else np.ascontiguousarray(masks, np.uint32)
        h, ms = C.c_void_p(), C.c_double()
        _check(L, L.ref_build_scene(handles, _p(self.frames), _p(self.masks), n, build_flags, reps, C.byref(h), C.byref(ms)), "ref_build_scene")
        self.h, self.build_ms, self.count = h, ms.value, n

    @property
    def device_handle(self):
        return lib().ref_scene_device_handle(self.h)

    def trace(self, rays, kind=2, reps=1):
        return trace_scene(self.device_handle, rays, kind, reps)

    def destroy(self):
        if self.h:
            lib().ref_destroy_scene(self.h)
            self.h = None
