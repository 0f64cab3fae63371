"""CPU-only: the PRODUCT's device code for animated instances -- include/hiprt/hiprt_b200_frame.h, the header the CUDA traversal
objects, the transform queries and the scene builder compile -- built for the host (tests/host/product_frame_host.cpp: device
qualifiers defined away, nothing else changed) and held BIT FOR BIT to the reference's own Transform / Quaternion code run on the
CPU (tests/golden/ref_math.npz, see tests/test_oracle_vs_reference_math.py): key-frame selection and interpolation, the ray
carried into an instance, points and normal-like vectors carried both ways. This is a check of product source, not of the oracle;
what the GPU adds on top is the compiler's own contraction of a * b + c, which the GPU tests bound."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.golden import make_ref_math as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INCLUDE = "/usr/local/cuda/include"
pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INCLUDE, "vector_types.h")),
                                reason="needs g++ and the CUDA toolkit's vector type headers")
GOLDEN = np.load(os.path.join(M.HERE, "ref_math.npz"))
INPUTS = M.ref_math_inputs()


@pytest.fixture(scope="module")
def product(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("host") / "libproduct_frame_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I", CUDA_INCLUDE,
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "product_frame_host.cpp"), "-o", so])
    return C.CDLL(so)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_ray_carried_into_an_animated_instance_is_the_references(product):
    rays, times = INPUTS["motion_rays"], INPUTS["motion_times"]
    for j in range(int(INPUTS["motion_count"])):
        stored = np.ascontiguousarray(GOLDEN[f"motion_stored_{j}"])
        out = np.zeros_like(rays)
        product.prod_motion_ray(len(stored), M.p(stored), len(rays), M.p(times), M.p(rays), M.p(out))
        assert np.array_equal(out, GOLDEN[f"motion_out_{j}"]), j


def test_points_and_vectors_under_the_interpolated_pose_are_the_references(product):
    pts = np.ascontiguousarray(INPUTS["motion_rays"].view(M.RAY).reshape(-1)["direction"][:100])
    for j in range(int(INPUTS["motion_count"])):
        stored = np.ascontiguousarray(GOLDEN[f"motion_stored_{j}"])
        for k, time in enumerate(M.POINT_TIMES):
            out = np.zeros((len(pts), 12), np.float32)
            product.prod_motion_points(len(stored), M.p(stored), C.c_float(time), len(pts), M.p(pts), M.p(out))
            assert np.array_equal(bits(out), bits(GOLDEN[f"motion_points_{j}"][k])), (j, time)
