"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs."""
import numpy as np
import pytest

from hiprt_b200 import api, scenes
from oracle import binding as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def to_dev(ctx, a, dtype=None):
    t = ctx.torch
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return t.from_numpy(a).to(ctx.device)


def compare_hits(got, ref, rel=1e-5, uv_tol=None):
    """primID exact; t within `rel` relative, uv within `uv_tol` (default `rel`) absolute. Returns the number of
    exact-tie disagreements."""
    uv_tol = rel if uv_tol is None else uv_tol
    same = got["primID"] == ref["primID"]
    hit = ref["primID"] != orc.INVALID
    ties = 0
    if not same.all():
        bad = ~same
        # tie-aware: different primitives are tolerated only at (nearly) the same distance
        both_hit = (got["primID"][bad] != orc.INVALID) & (ref["primID"][bad] != orc.INVALID)
        assert both_hit.all(), f"{(~both_hit).sum()} rays hit in one implementation only"
        dt = np.abs(got["t"][bad] - ref["t"][bad]) / np.maximum(np.abs(ref["t"][bad]), 1e-20)
        assert (dt <= rel).all(), f"{(dt > rel).sum()} primID mismatches away from ties (max rel dt {dt.max():.3g})"
        ties = int(bad.sum())
    ok = same & hit
    dt = np.abs(got["t"][ok] - ref["t"][ok]) / np.maximum(np.abs(ref["t"][ok]), 1e-20)
    assert dt.size == 0 or dt.max() <= rel, f"t differs: max rel {dt.max():.3g}"
    duv = np.abs(got["uv"][ok] - ref["uv"][ok])
    assert duv.size == 0 or duv.max() <= uv_tol, f"uv differs: max abs {duv.max():.3g}"
    return ties


def check_build_against_oracle(ctx, geom, og, expect_same_keys=False):
    """Binary topology and the canonicalised wide structure must be bit-exact for identical sorted keys."""
    n = og.prim_count
    hdr = geom.header()
    assert int(hdr["primNodeCount"]) == n
    if n > 1:
        assert geom.build_error_flag() == 0, "the collapse watchdog fired"
    if n == 1:  # singleton: no sort, no binary tree (reference: SingletonConstruction)
        boxes, leaves = geom.nodes()
        assert boxes.tobytes() == og.box_nodes.tobytes() and leaves.tobytes() == og.leaves.tobytes()
        return 0.0
    skeys = geom.temp_array("keys0", np.uint32, n)
    svals = geom.temp_array("vals0", np.uint32, n)
    # sorted, stable, a permutation
    assert np.all(skeys[1:] >= skeys[:-1])
    eq = skeys[1:] == skeys[:-1]
    assert np.all(svals[1:][eq] > svals[:-1][eq]), "sort is not stable"
    assert np.array_equal(np.sort(svals), np.arange(n, dtype=np.uint32))
    # Morton codes: report mismatches against the IEEE oracle (the GPU uses fast division like the reference)
    keys_by_prim = np.empty(n, np.uint32)
    keys_by_prim[svals] = skeys
    differs = np.nonzero(keys_by_prim != og.keys)[0]
    mismatch = len(differs) / n
    if expect_same_keys:
        assert mismatch == 0.0
    # every mismatch must be a quantisation-boundary case: the GPU divides with div.approx (2 ulp), as
    # the reference's --use_fast_math build does, so a position within 2 ulp of a cell boundary may land
    # in the neighbouring cell. Perturbing the oracle's normalised position by <= 2 ulp must reproduce it.
    pb, sb = og.prim_boxes, og.scene_box
    ext = (sb[1] - sb[0]).astype(np.float32)
    for p in differs[:64]:
        c = ((pb[p, 1] + pb[p, 0]) * np.float32(0.5)).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            pos = ((c - sb[0]).astype(np.float32) / ext).astype(np.float32)
        found = False
        for dx in range(-2, 3):
            for dy in range(-2, 3):
                for dz in range(-2, 3):
                    q = pos.copy()
                    for ax, k in enumerate((dx, dy, dz)):
                        for _ in range(abs(k)):
                            q[ax] = np.nextafter(q[ax], np.float32(np.inf if k > 0 else -np.inf))
                    if orc.morton(q, ext) == int(keys_by_prim[p]):
                        found = True
        assert found, f"prim {p}: GPU key {keys_by_prim[p]:#x} vs oracle {og.keys[p]:#x} is not a boundary case"
    # topology on identical keys
    og.rebuild_from_sorted(skeys, svals)
    if n > 1:
        gscratch = geom.temp_array("scratch", api.SCRATCH_DTYPE, n - 1)
        oscratch = og.scratch
        assert np.array_equal(gscratch["child0"], oscratch["child0"])
        assert np.array_equal(gscratch["child1"], oscratch["child1"])
        assert np.array_equal(gscratch["box"].view(np.uint32), oscratch["box"].view(np.uint32))
    boxes, leaves = geom.nodes()
    cb, cl = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaves.view(orc.TRIPAIR_DTYPE))
    ob, ol = og.box_nodes, og.leaves
    assert len(cb) == len(ob) == int(hdr["boxNodeCount"])
    assert cb.tobytes() == ob.tobytes(), "wide nodes differ from the oracle"
    assert cl.tobytes() == ol.tobytes(), "leaf nodes differ from the oracle"
    return mismatch


def test_c1_cornell_box_hits_match_oracle(ctx):
    verts, _, _ = H.cornell_box()
    idx = np.arange(96, dtype=np.uint32).reshape(-1, 3)
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    og = orc.Geometry(verts, idx)
    check_build_against_oracle(ctx, geom, og)
    rays = scenes.cornell_rays(512, 512)
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
    ref = og.trace(rays, tie_mode=1)
    assert np.array_equal(got["primID"], ref["primID"])
    assert np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))
    hit = ref["primID"] != orc.INVALID  # uv / normal of a miss are unspecified (hiprtHit leaves them uninitialised)
    assert np.array_equal(got["uv"][hit].view(np.uint32), ref["uv"][hit].view(np.uint32))
    assert np.array_equal(got["normal"][hit].view(np.uint32), ref["normal"][hit].view(np.uint32))
    # and the end-to-end host-buffer call returns the same thing
    got2 = ctx.trace_host(geom, rays)
    assert np.array_equal(got2["primID"], got["primID"]) and np.array_equal(got2["t"].view(np.uint32), got["t"].view(np.uint32))
    geom.destroy()


def test_anyhit_agrees_with_closest_on_occlusion(ctx):
    verts, _, _ = H.cornell_box()
    idx = np.arange(96, dtype=np.uint32).reshape(-1, 3)
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    rays = scenes.cornell_rays(256, 256)
    d = ctx.rays_to_device(rays)
    closest = ctx.hits_to_host(ctx.trace(geom, d))
    anyhit = ctx.hits_to_host(ctx.trace(geom, d, kind=api.ANY_HIT))
    assert np.array_equal(closest["primID"] != api.INVALID, anyhit["primID"] != api.INVALID)
    geom.destroy()


@pytest.mark.parametrize("n_tris", [1, 2, 3, 5, 33, 1000, 100_000])
def test_soup_build_and_hits(ctx, n_tris):
    verts = scenes.random_soup(n_tris, seed=n_tris)
    geom = ctx.build(to_dev(ctx, verts))
    og = orc.Geometry(verts)
    check_build_against_oracle(ctx, geom, og)
    rays = scenes.random_rays(20000, verts.min(0), verts.max(0), seed=3)
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
    ref = og.trace(rays, tie_mode=1)
    compare_hits(got, ref)
    geom.destroy()


@pytest.mark.parametrize("shape", [(3, 1), (16, 16), (300, 300)])
def test_indexed_mesh_pairing_topology_hits(ctx, shape):
    verts, idx = scenes.sphere_mesh(*shape)
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    og = orc.Geometry(verts, idx)
    n = og.prim_count
    # deterministic pairing: identical pair list, in group order
    L = api.temp_layout(len(idx), len(idx))
    ctx.torch.cuda.synchronize()
    pairs = geom.temp[L["pairs"]: L["pairs"] + n * 8].cpu().numpy().view(np.uint32).reshape(-1, 2)
    assert int(geom.header()["primNodeCount"]) == n
    assert np.array_equal(pairs, og.pairs)
    check_build_against_oracle(ctx, geom, og)
    rays = scenes.pinhole_rays(256, 256, (0, 0, 3), (0, 0, 0))
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
    ref = og.trace(rays, tie_mode=1)
    compare_hits(got, ref)
    geom.destroy()


def test_planar_grid_uses_2d_morton(ctx):
    verts, idx = scenes.grid_mesh(64, 48)
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    og = orc.Geometry(verts, idx)
    check_build_against_oracle(ctx, geom, og)
    rays = H.ortho_rays(128, 128)
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
    compare_hits(got, og.trace(rays, tie_mode=1))
    geom.destroy()


def test_identical_triangles_all_equal_morton_codes(ctx):
    # reference test PlocFallback / degenerate input: every primitive has the same code
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    verts = np.tile(tri, (5000, 1))
    geom = ctx.build(to_dev(ctx, verts))
    og = orc.Geometry(verts)
    check_build_against_oracle(ctx, geom, og)
    rays = scenes.make_rays([[0.2, 0.2, -1.0]], [[0, 0, 1.0]])
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
    assert got["primID"][0] == 0 and got["t"][0] == 1.0  # min (t, primID)
    geom.destroy()


def test_disable_pairing_flag(ctx):
    verts, idx = scenes.sphere_mesh(32, 16)
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx), flags=api.FAST_BUILD | api.DISABLE_PAIRING)
    og = orc.Geometry(verts, idx, allow_pairing=False)
    assert int(geom.header()["primNodeCount"]) == len(idx) == og.prim_count
    check_build_against_oracle(ctx, geom, og)
    geom.destroy()


# ---- balanced builder (H-PLOC): topology is not pinned by the reference ("balanced-build topology is not
# parity-checked, only hits are"), so: structural validity, hits equal to the oracle's, and better SAH than the fast build.
BALANCED_CASES = {
    "soup2": lambda: (scenes.random_soup(2, seed=1), None),
    "soup17": lambda: (scenes.random_soup(17, seed=2), None),
    "soup33": lambda: (scenes.random_soup(33, seed=3), None),
    "soup5000": lambda: (scenes.random_soup(5000, seed=4), None),
    "soup200k": lambda: (scenes.random_soup(200_000, seed=5), None),
    "sphere": lambda: scenes.sphere_mesh(300, 300),
    "grid": lambda: scenes.grid_mesh(64, 48),
    "identical": lambda: (np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (3000, 1)), None),
}


@pytest.mark.parametrize("case", sorted(BALANCED_CASES))
def test_balanced_build_is_valid_and_hits_match(ctx, case):
    verts, idx = BALANCED_CASES[case]()
    d_idx = None if idx is None else to_dev(ctx, idx)
    n_tris = len(verts) // 3 if idx is None else len(idx)
    geom = ctx.build(to_dev(ctx, verts), d_idx, flags=api.BALANCED_BUILD)
    boxes, leaves = geom.nodes()
    cost_balanced = H.validate_wide_bvh(boxes, leaves, n_tris)
    fast = ctx.build(to_dev(ctx, verts), d_idx, flags=api.FAST_BUILD)
    cost_fast = H.validate_wide_bvh(*fast.nodes(), n_tris)
    if case.startswith("soup") and n_tris >= 5000:  # irregular input: clustering must beat the Morton-order tree
        assert cost_balanced < cost_fast, f"H-PLOC SAH {cost_balanced:.1f} is not better than LBVH {cost_fast:.1f}"
    elif n_tris >= 3000 and case != "identical":  # regular grids: LBVH is already near-optimal
        assert cost_balanced < 1.05 * cost_fast, f"H-PLOC SAH {cost_balanced:.1f} is much worse than LBVH {cost_fast:.1f}"
    og = orc.Geometry(verts, idx)
    lo, hi = verts.min(0), verts.max(0)
    rays = scenes.random_rays(20000, lo - 0.05, hi + 0.05, seed=9)
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
    compare_hits(got, og.trace(rays, tie_mode=1))
    # building twice gives the same tree shape (node addresses may differ): same canonical structure
    geom2 = ctx.build(to_dev(ctx, verts), d_idx, flags=api.BALANCED_BUILD)
    b2, l2 = geom2.nodes()
    c1 = orc.canonicalize(boxes.view(orc.BOX4_DTYPE), leaves.view(orc.TRIPAIR_DTYPE))
    c2 = orc.canonicalize(b2.view(orc.BOX4_DTYPE), l2.view(orc.TRIPAIR_DTYPE))
    assert c1[0].tobytes() == c2[0].tobytes() and c1[1].tobytes() == c2[1].tobytes(), "H-PLOC tree shape is not deterministic"
    for g in (geom, fast, geom2):
        g.destroy()


def test_fast_build_structure_is_valid(ctx):
    verts, idx = scenes.sphere_mesh(100, 50)
    geom = ctx.build(to_dev(ctx, verts), to_dev(ctx, idx))
    H.validate_wide_bvh(*geom.nodes(), len(idx))
    geom.destroy()


def _octant_rays(lo, hi, n, seed):
    """Warps of 32 rays sharing a direction octant (the fast kernel's specialised slab code), then warps whose rays
    have exactly-zero / negative-zero direction components or mixed octants (its general code)."""
    rng = np.random.default_rng(seed)
    centre, ext = (lo + hi) / 2, (hi - lo)
    origins, dirs = [], []
    for octant in range(8):
        sign = np.array([-1.0 if octant & (1 << k) else 1.0 for k in range(3)])
        o = centre - sign * ext * rng.uniform(0.6, 1.4, (n, 3))
        target = centre + ext * rng.uniform(-0.5, 0.5, (n, 3))
        d = np.abs(target - o) * sign + sign * 1e-3
        origins.append(o), dirs.append(d)
    # axis-aligned and planar rays: zero components (both signs of zero)
    for axis in range(3):
        for s in (1.0, -1.0):
            o = centre + ext * rng.uniform(-0.5, 0.5, (n, 3))
            o[:, axis] = centre[axis] - s * ext[axis] * 1.2
            d = np.zeros((n, 3))
            d[:, axis] = s
            d[::2, (axis + 1) % 3] = -0.0
            d[::3, (axis + 2) % 3] = 0.25 * s
            origins.append(o), dirs.append(d)
    o = centre + ext * rng.uniform(-1.3, 1.3, (n, 3))
    d = rng.normal(size=(n, 3))
    origins.append(o), dirs.append(d)
    return orc.make_rays(np.concatenate(origins).astype(np.float32), np.concatenate(dirs).astype(np.float32))


@pytest.mark.parametrize("mesh", ["cornell", "sphere", "soup"])
def test_fast_kernel_octants_and_degenerate_directions(ctx, mesh):
    """The library's triangle kernel takes the near/far planes from the warp's direction octant when it can; hits must be
    bit-identical to the oracle (which restates the reference's min/max slab test) in every octant, for rays with
    zero / negative-zero direction components and for mixed warps."""
    if mesh == "cornell":
        verts, _, _ = H.cornell_box()
        idx = np.arange(96, dtype=np.uint32).reshape(-1, 3)
    elif mesh == "sphere":
        verts, idx = scenes.sphere_mesh(96, 48)
    else:
        verts, idx = scenes.random_soup(3000, seed=11, extent=1.0, size=0.08), None
    geom = ctx.build(to_dev(ctx, verts), None if idx is None else to_dev(ctx, idx))
    og = orc.Geometry(verts, idx)
    rays = _octant_rays(verts.min(0), verts.max(0), 1024, seed=5)
    got = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays)))
    ref = og.trace(rays, tie_mode=1)
    assert (ref["primID"] != orc.INVALID).mean() > 0.2
    assert np.array_equal(got["primID"], ref["primID"])
    hit = ref["primID"] != orc.INVALID
    assert np.array_equal(got["t"][hit], ref["t"][hit]) and np.array_equal(got["uv"][hit], ref["uv"][hit])
    occluded = ctx.hits_to_host(ctx.trace(geom, ctx.rays_to_device(rays), kind=api.ANY_HIT))
    assert np.array_equal(occluded["primID"] != api.INVALID, ref["primID"] != orc.INVALID)
    geom.destroy()
