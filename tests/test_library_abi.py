"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and
exports every symbol include/hiprt/hiprt.h and include/hiprt/hiprt_b200.h declare."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "hiprt_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL)
    from hiprt_b200 import api

    return api.lib()


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", "hiprt", header)).read()
    txt = re.sub(r"//.*", "", txt)
    return sorted(set(re.findall(r"\b(hiprt[A-Z]\w+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions("hiprt.h") + declared_functions("hiprt_b200.h")
    assert len(declared_functions("hiprt.h")) == 37  # the reference's API surface (hiprt/hiprt.h.in:76-601)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared but not exported"


def test_binding_covers_the_whole_api(lib):
    from hiprt_b200 import api

    assert sorted(api.HIPRT_API_FUNCTIONS) == declared_functions("hiprt.h")


def test_version_gate_and_argument_validation(lib):
    from hiprt_b200 import api

    inp = api.ContextCreationInput(None, 0, api.DEVICE_NVIDIA)
    h = C.c_void_p()
    # reference: hiprt/impl/hiprt.cpp:37
    assert lib.hiprtCreateContext(3000, C.byref(inp), C.byref(h)) == api.ERR_API_VERSION
    assert lib.hiprtDestroyContext(None) == api.ERR_INVALID_PARAM
    assert lib.hiprtDestroyGeometry(None, None) == api.ERR_INVALID_PARAM
    assert lib.hiprtSetLogLevel(None, 0) == api.ERR_INVALID_PARAM


def test_struct_layouts_match_the_reference_abi():
    # SURVEY.md appendix B (offsets verified against the reference headers)
    from hiprt_b200 import api

    assert api.GeometryBuildInput.primitive.offset == 8 and api.GeometryBuildInput.nodeList.offset == 56
    assert api.SceneBuildInput.nodeList.offset == 32 and api.SceneBuildInput.instanceCount.offset == 56
    assert api.SceneBuildInput.frameType.offset == 64
    m = api.TriangleMeshPrimitive
    assert (m.vertexCount.offset, m.vertexStride.offset, m.triangleIndices.offset, m.triangleCount.offset,
            m.triangleStride.offset, m.trianglePairIndices.offset, m.trianglePairCount.offset) == (8, 12, 16, 24, 28, 32, 40)
    assert api.HIT_DTYPE.fields["primID"][1] == 16 and api.HIT_DTYPE.fields["uv"][1] == 24
    assert api.HIT_DTYPE.fields["normal"][1] == 32 and api.HIT_DTYPE.fields["t"][1] == 44
    assert api.BOX4_DTYPE.fields["box"][1] == 16 and api.BOX4_DTYPE.fields["parentAddr"][1] == 112


def test_device_headers_are_embedded(lib):
    names = C.POINTER(C.c_char_p)()
    srcs = C.POINTER(C.c_char_p)()
    n = lib.hiprtB200GetDeviceHeaders(C.byref(names), C.byref(srcs))
    got = {names[i].decode(): srcs[i].decode() for i in range(n)}
    assert "hiprt/hiprt_device.h" in got and "hiprtGeomTraversalClosest" in got["hiprt/hiprt_device.h"]


def test_no_cpu_fallback_without_gpu():
    import torch

    from hiprt_b200 import api

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        api.Context(0)
