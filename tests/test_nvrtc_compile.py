"""CPU-only: the public device header (embedded in the shared library) and every kernel that is compiled at run time through
hiprtBuildTraceKernels compile under NVRTC for sm_100a exactly as the library assembles them (no GPU needed for compilation)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

nvrtc_check = pytest.importorskip("nvrtc_check")


def test_user_test_kernels_compile_for_sm_100a():
    src = open(os.path.join(ROOT, "tests", "kernels", "user_kernels.cu")).read()
    err, log = nvrtc_check.compile_user_source(src, nvrtc_check.DISPATCH)
    assert err == 0, log[-3000:]


def test_pathtrace_bounce_kernel_compiles_for_sm_100a():
    src = open(os.path.join(ROOT, "hiprt_b200", "pathtrace_kernels.cu")).read()
    default_table = ("__device__ bool intersectFunc( uint32_t, uint32_t, const hiprtFuncTableHeader&, const hiprtRay&, void*, hiprtHit& ) { return false; }\n"
                     "__device__ bool filterFunc( uint32_t, uint32_t, const hiprtFuncTableHeader&, const hiprtRay&, void*, const hiprtHit& ) { return false; }\n")
    err, log = nvrtc_check.compile_user_source(src, default_table)
    assert err == 0, log[-3000:]


def test_embedded_headers_include_the_frame_header():
    import ctypes as C

    from hiprt_b200 import api

    names, srcs = C.POINTER(C.c_char_p)(), C.POINTER(C.c_char_p)()
    n = api.lib().hiprtB200GetDeviceHeaders(C.byref(names), C.byref(srcs))
    have = {names[i].decode() for i in range(n)}
    assert {"hiprt/hiprt_device.h", "hiprt/hiprt_b200_frame.h", "hiprt/hiprt_b200_layout.h", "hiprt/hiprt_types.h"} <= have
