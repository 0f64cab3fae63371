"""CPU-only: the oracle's arithmetic against the REFERENCE'S OWN CODE run on the CPU. oracle/_ref/libref_host_math.so is the
reference's HIPRT_HOST_DEVICE headers (MortonCode.h, Aabb.h, Triangle.h, Transform.h, Quaternion.h, QrDecomposition.h) compiled
for the host where they lie; its outputs on seeded inputs are committed as tests/golden/ref_math.npz (made by
tests/golden/make_ref_math.py). The oracle's restatement of the same functions must reproduce them BIT FOR BIT; where the
compiled reference is present (this container) it is also run live, which checks that the committed vectors are current."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import binding as orc
from tests.golden import make_ref_math as M

GOLDEN = np.load(os.path.join(M.HERE, "ref_math.npz"))
INPUTS = M.ref_math_inputs()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


@pytest.fixture(scope="module")
def oracle_out():
    return M.run(orc.lib(), "orc_math_", INPUTS)


def test_morton_codes_are_the_references(oracle_out):
    for e in range(len(INPUTS["morton_ext"])):
        assert np.array_equal(oracle_out[f"morton_{e}"], GOLDEN[f"morton_{e}"]), e


def test_triangle_test_is_the_references(oracle_out):
    assert 0.2 < GOLDEN["tri_hit"].mean() < 0.8
    assert np.array_equal(oracle_out["tri_hit"], GOLDEN["tri_hit"])
    assert np.array_equal(bits(oracle_out["tri_t"]), bits(GOLDEN["tri_t"]))
    assert np.array_equal(bits(oracle_out["tri_uv"]), bits(GOLDEN["tri_uv"]))


def test_slab_test_is_the_references(oracle_out):
    a, b = oracle_out["slab_t01"], GOLDEN["slab_t01"]
    both_nan = np.isnan(a) & np.isnan(b)  # 0 * inf on a box plane: NaN in both, payloads aside
    assert np.array_equal(bits(a)[~both_nan], bits(b)[~both_nan])
    assert ((b[:, 0] <= b[:, 1]).mean() > 0.02) and np.isinf(INPUTS["slab_inv"]).any()


def test_triangle_pair_leaves_are_the_references(oracle_out):
    """TriangleMesh::fetchTrianglePairNode: vertex rotation, primitive indices and the flags of all 54 ways to share an edge."""
    from oracle.binding import TRIPAIR_DTYPE

    got, want = oracle_out["pair_nodes"].view(TRIPAIR_DTYPE).reshape(-1), GOLDEN["pair_nodes"].view(TRIPAIR_DTYPE).reshape(-1)
    assert len(np.unique(want["flags"])) >= 6
    for field in ("v", "primIndex0", "primIndex1", "flags"):
        assert np.array_equal(bits(got[field]), bits(want[field])), field
    assert np.array_equal(bits(oracle_out["pair_boxes"]), bits(GOLDEN["pair_boxes"]))


@pytest.mark.parametrize("name", ["frames_srt", "frames_matrix"])
def test_frame_decomposition_and_inverse_are_the_references(oracle_out, name):
    assert np.array_equal(bits(oracle_out[name + "_decomposed"]), bits(GOLDEN[name + "_decomposed"]))
    assert np.array_equal(bits(oracle_out[name + "_inverse"]), bits(GOLDEN[name + "_inverse"]))


def test_motion_ray_transform_is_the_references():
    rays = INPUTS["motion_rays"].view(M.RAY).reshape(-1)
    for j in range(int(INPUTS["motion_count"])):
        keys, kind = INPUTS[f"motion_keys_{j}"], int(INPUTS[f"motion_kind_{j}"])
        want = GOLDEN[f"motion_out_{j}"].view(M.RAY).reshape(-1)
        for i, time in enumerate(INPUTS["motion_times"]):
            _, got = orc.motion_pose(keys, float(time), rays[i], frame_type=kind)
            for field in ("origin", "direction", "minT", "maxT"):
                assert np.array_equal(bits(np.atleast_1d(got[field])), bits(np.atleast_1d(want[i][field]))), (j, i, field)


def test_motion_bounds_are_the_references():
    """Transform::motionBounds: what the scene builder bounds an (animated) instance with."""
    got = M.run_motion_bounds(orc.lib(), "orc_math_", INPUTS)
    for k, v in got.items():
        assert np.array_equal(bits(v), bits(GOLDEN[k])), k


@pytest.mark.skipif(not os.path.exists(M.REF_LIB), reason="the compiled reference (oracle/_ref) is not on this machine")
def test_committed_vectors_are_what_the_compiled_reference_returns():
    L = C.CDLL(M.REF_LIB)
    live = M.run(L, "ref_", INPUTS)
    live.update(M.run_reference_motion(L, INPUTS))
    live.update(M.run_motion_bounds(L, "ref_", INPUTS))
    live.update(M.run_reference_points(L, INPUTS))
    assert set(live) == set(GOLDEN.files)
    for k, v in live.items():
        assert np.array_equal(bits(v), bits(GOLDEN[k]), equal_nan=False) or np.array_equal(v, GOLDEN[k], equal_nan=True), k
