"""CPU-only: the device-independent half of hiprtBuildTraceKernelsFromBitcode (reference: Compiler::buildKernelsFromBitcode,
hiprt/impl/Compiler.cpp:338-470; its tests build every kernel this way when the bitcode option is on). The test kernels are
compiled offline with nvcc as relocatable device code for sm_100a -- intersectFunc / filterFunc left undefined, as an application
shipping precompiled kernels would -- and hiprtB200LinkTraceKernels generates the dispatchers for a function-name table, compiles
them with NVRTC and links both with nvJitLink. No GPU is needed up to the linked cubin; loading it (cuModuleLoadData +
cuModuleGetFunction) is the path hiprtBuildTraceKernels already takes."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

from hiprt_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")
KERNELS = ("customKernel", "cutoutKernel", "allHitsKernel", "sceneAllHitsKernel", "aoKernel", "motionSamplesKernel")


@pytest.fixture(scope="module")
def images(tmp_path_factory):
    """tests/kernels/user_kernels.cu as relocatable device code: a cubin (SASS) and PTX."""
    d = tmp_path_factory.mktemp("bitcode")
    out = {}
    for kind, code in (("cubin", "sm_100a"), ("ptx", "compute_100a")):
        path = str(d / f"user_kernels.{kind}")
        subprocess.check_call([NVCC, "-gencode", f"arch=compute_100a,code={code}", "-std=c++17", "-dc", f"-{kind}",
                               "-I", os.path.join(ROOT, "include"), "-include", "hiprt/hiprt_device.h", "-x", "cu",
                               os.path.join(ROOT, "tests", "kernels", "user_kernels.cu"), "-o", path])
        out[kind] = open(path, "rb").read()
    return out


def link(image, name_sets, num_geom_types):
    sets = (api.FuncNameSet * len(name_sets))(*[api.FuncNameSet(a.encode() if a else None, b.encode() if b else None) for a, b in name_sets])
    size = C.c_size_t(0)
    log = C.create_string_buffer(1 << 16)
    args = (10, 0, b"user_kernels", image, len(image), num_geom_types, 1, sets)
    rc = api.lib().hiprtB200LinkTraceKernels(*args, None, C.byref(size), log, len(log))
    if rc != 0:
        return rc, b"", log.value.decode(errors="replace")
    cubin = C.create_string_buffer(size.value)
    rc = api.lib().hiprtB200LinkTraceKernels(*args, cubin, C.byref(size), log, len(log))
    return rc, cubin.raw[: size.value], log.value.decode(errors="replace")


def elf_dump(cubin, tmp_path):
    path = str(tmp_path / "linked.cubin")
    with open(path, "wb") as f:
        f.write(cubin)
    return subprocess.check_output([os.path.join(os.path.dirname(NVCC), "cuobjdump"), "-elf", path]).decode()


@pytest.mark.parametrize("kind", ["cubin", "ptx"])
def test_offline_kernels_link_with_generated_dispatchers(images, tmp_path, kind):
    rc, cubin, log = link(images[kind], [("intersectSphere", "cutoutFilter")], 1)
    assert rc == 0, log
    assert cubin[:4] == b"\x7fELF"
    dump = elf_dump(cubin, tmp_path)
    assert "sm_100" in dump
    for kernel in KERNELS:
        assert f".text.{kernel}" in dump, kernel


def test_two_geometry_types_and_a_missing_function_is_reported(images):
    rc, cubin, log = link(images["cubin"], [("intersectCircle", None), ("intersectSphere", None)], 2)
    assert rc == 0 and cubin[:4] == b"\x7fELF", log
    rc, _, log = link(images["cubin"], [("noSuchIntersector", None)], 1)
    assert rc == api.ERR_INTERNAL
    assert "noSuchIntersector" in log


def test_garbage_image_and_arguments_are_rejected():
    rc, _, log = link(b"this is not device code" * 4, [(None, None)], 1)
    assert rc == api.ERR_INTERNAL and log
    size = C.c_size_t(0)
    assert api.lib().hiprtB200LinkTraceKernels(10, 0, None, None, 0, 0, 1, None, None, C.byref(size), None, 0) == api.ERR_INVALID_PARAM
    assert api.lib().hiprtBuildTraceKernelsFromBitcode(None, 1, None, None, b"x", 1, 0, 1, None, None, False) == api.ERR_INVALID_PARAM
