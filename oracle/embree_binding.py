"""ctypes binding of oracle/_ref/embree/libembree_bench.so: Embree 4.3 (the binary bundled with
the reference) as the CPU baseline. TEST / BENCHMARK INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from .binding import HIT_DTYPE, RAY_DTYPE

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "embree")
LIB = os.path.join(_DIR, "libembree_bench.so")


def available():
    return os.path.exists(LIB)


_lib = None


def lib(threads=0):
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
        L.emb_init.argtypes = [i32]
        L.emb_build.argtypes = [vp, u32, vp, u32, i32, C.POINTER(C.c_double)]
        L.emb_build.restype = vp
        L.emb_release.argtypes = [vp]
        L.emb_trace.argtypes = [vp, vp, u32, vp, i32, i32]
        L.emb_trace.restype = C.c_double
        if L.emb_init(threads) != 0:
            raise RuntimeError("rtcNewDevice failed")
        _lib = L
    return _lib


class EmbreeScene:
    def __init__(self, vertices, indices=None, quality=0, threads=0):
        self.v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        self.i = None if indices is None else np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        ntri = len(self.v) // 3 if self.i is None else len(self.i)
        ms = C.c_double()
        L = lib(threads)
        self.h = L.emb_build(self.v.ctypes.data_as(C.c_void_p), len(self.v),
                             None if self.i is None else self.i.ctypes.data_as(C.c_void_p), ntri, quality, C.byref(ms))
        self.build_ms, self.tri_count, self.threads = ms.value, ntri, threads

    def trace(self, rays, kind=2, want_hits=True):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE) if want_hits else None
        ms = lib().emb_trace(self.h, rays.ctypes.data_as(C.c_void_p), len(rays),
                             None if hits is None else hits.ctypes.data_as(C.c_void_p), kind, self.threads)
        return hits, ms

    def release(self):
        if self.h:
            lib().emb_release(self.h)
            self.h = None
