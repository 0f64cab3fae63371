"""hiprt_b200 -- B200-native implementation of the HIPRT ray-tracing hot path.

The product is the shared library libhiprt_b200.so (C-ABI: include/hiprt/hiprt.h,
include/hiprt/hiprt_b200.h; CUDA sources in hiprt_b200/csrc). This package only adds the
ctypes binding used by the tests and the benchmark.
"""
from . import api  # noqa: F401
