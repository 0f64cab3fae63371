"""The CUDA path against the reference's own golden images (the same image formation as tests/test_oracle_golden.py, hits from
the GPU through the C-ABI instead of the oracle): cornellBox.obj as a two-level scene -- PrimIds, HitDist, Uvs, Normals,
PrimaryRay colours -- and the same scene with every instance translated / scaled / rotated (hiprtFrameSRT); MinimumCornellBox
through the bare geometry. Tolerance = the reference's (a pixel is bad when a channel differs by more than 10; < 0.3 % bad)."""
import numpy as np
import pytest

from hiprt_b200 import api
from oracle import binding as orc
from tests import helpers as H
from tests import test_oracle_golden as G
from tests.test_gpu_parity import to_dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def gpu_obj_hits(ctx, srt_frame=None):
    d = np.load(H.os.path.join(H.GOLDEN, "cornell_obj.npz"))
    n = int(d["shape_count"])
    d_verts = to_dev(ctx, d["vertices"])
    geoms = [ctx.build(d_verts, to_dev(ctx, d[f"tris_{k}"])) for k in range(n)]
    identity = [0, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0]  # hiprtFrameSRT: rotation (axis, angle 0), scale, translation, time
    frames = np.tile(np.asarray(identity if srt_frame is None else srt_frame, np.float32).reshape(1, 12), (n, 1))
    scene = ctx.build_scene(geoms, frames, frame_type=api.FRAME_SRT)
    hits = ctx.hits_to_host(ctx.trace(scene, ctx.rays_to_device(H.obj_camera_rays()))).copy()
    scene.destroy()
    for g in geoms:
        g.destroy()
    return hits


def test_cornell_obj_images_from_the_gpu(ctx, monkeypatch):
    hits = gpu_obj_hits(ctx)
    hit = hits["primID"] != api.INVALID
    assert hit.all()
    # reuse the oracle tests' image formation with the GPU's hits
    monkeypatch.setattr(G, "_obj_hits", lambda: (hits, hit))
    G.test_prim_ids_cornell_box_matches_reference_image()
    G.test_hit_dist_cornell_box_matches_reference_image()
    G.test_uvs_cornell_box_matches_reference_image()
    G.test_normals_cornell_box_matches_reference_image()
    G._check(G._colour_image(hits), "PrimaryRayCornellBox")


@pytest.mark.parametrize("name, frame", [
    ("TranslateCornellBox", [0, 0, 1, 0, 1, 1, 1, 0, 0, -5, 0, 0]),
    ("ScaleCornellBox", [0, 0, 1, 0, 2, 2, 2, 0, 0, 0, 0, 0]),
    ("RotateCornellBox", [0, 1, 0, 0.5, 1, 1, 1, 0, 0, -3, 0, 0]),
])
def test_transformed_cornell_obj_images_from_the_gpu(ctx, name, frame):
    G._check(G._colour_image(gpu_obj_hits(ctx, frame)), name)
