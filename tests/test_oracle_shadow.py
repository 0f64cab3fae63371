"""CPU-only: the oracle against the reference's ShadowRayCornellBox.png (test/main.cpp:414-424, test/kernels/ShadowRayKernel.h):
a closest-hit primary ray through the pixel centre, then per area-light triangle 32 any-hit shadow rays towards points sampled
on the light with the pixel's own `tea<16>` / `lcg` sequence (test/shared.h:95-125), `maxT = 0.99 |light point - surface|` --
the primary + shadow-ray shape of BASELINE.json's configs[1]. The random sequence is reproduced sample by sample, so the
comparison is not statistical: 6 of 262 144 pixels differ by more than the reference's tolerance, mean |difference| 0.003."""
import numpy as np

from oracle import binding as orc
from tests import helpers as H
from tests.test_oracle_ao import normalize, randf

f32 = np.float32


def dot(a, b):
    return (a * b).sum(-1, dtype=f32)


def test_shadow_ray_cornell_box_matches_reference_image():
    d = np.load(H.os.path.join(H.GOLDEN, "cornell_obj.npz"))
    n = int(d["shape_count"])
    scene, _ = H.cornell_obj_scene()
    first = np.cumsum([0] + [len(d[f"tris_{k}"]) for k in range(n)])
    diffuse = np.concatenate([d[f"diffuse_{k}"] for k in range(n)])
    emission = np.concatenate([d[f"emission_{k}"] for k in range(n)])
    is_light = emission.sum(1) > 0  # Material::light(), test/shared.h:75
    # one Light per emissive face, radiance = emission + 40 (hiprtTest.cpp:1182-1199)
    lights = [(d["vertices"][tri].astype(f32), (emission[first[k] + f] + f32(40)).astype(f32))
              for k in range(n) for f, tri in enumerate(d[f"tris_{k}"]) if is_light[first[k] + f]]
    assert len(lights) == 2

    res, spp = 512, 32
    rays = H.obj_camera_rays()
    hits = scene.trace(rays, tie_mode=0)
    hit = hits["primID"] != orc.INVALID
    face = first[np.where(hit, hits["instanceIDs"][:, 0], 0)] + np.where(hit, hits["primID"], 0)
    direction, origin = rays["direction"], rays["origin"]
    ng = hits["normal"].astype(f32)
    ng = np.where((dot(direction, ng) > 0)[:, None], -ng, ng)
    ng = normalize(np.where(hit[:, None], ng, f32(1)))
    surface = (origin + hits["t"][:, None] * direction).astype(f32)
    xs = np.tile(np.arange(res, dtype=np.uint32), res)
    ys = np.repeat(np.arange(res, dtype=np.uint32), res)
    seed = H.tea16(xs, ys)[0].astype(np.uint32)
    shade = hit & ~is_light[face]
    est = np.zeros((res * res, 3), f32)
    for lv, le in lights:
        e1, e2 = lv[1] - lv[0], lv[2] - lv[0]
        ln = np.cross(e1, e2).astype(f32)
        area = f32(np.sqrt(dot(ln, ln)) / f32(2))
        ln = normalize(ln)
        for _ in range(spp):
            # sampleLightVertex (ShadowRayKernel.h:35-64): uniform point on the triangle, pdf turned into solid angle
            x1 = randf(seed)
            x2 = randf(seed)
            sq = np.sqrt(x1, dtype=f32)
            light_point = (lv[0] + (f32(1) - sq)[:, None] * e1 + (x2 * sq)[:, None] * e2).astype(f32)
            r = light_point - surface
            dist = np.sqrt(dot(r, r), dtype=f32)
            cos = np.abs(dot(ln, -(r / np.maximum(dist, f32(1e-30))[:, None])))
            pdf = dot(r, r) * (f32(1) / area) / np.maximum(cos, f32(1e-30))
            pdf = np.where((dist < f32(1e-4)) | (cos < f32(1e-4)) | (dot(r, ln) > 0), f32(0), pdf)
            light_dir = normalize(np.where(shade[:, None], r, f32(1)))
            light_vec = light_point + f32(1e-3) * ln - surface
            shadow = orc.make_rays(surface + f32(1e-3) * ng, light_dir)
            shadow["maxT"] = f32(0.99) * np.sqrt(dot(light_vec, light_vec), dtype=f32)
            lit = shade & (pdf != 0)
            blocked = np.zeros(res * res, bool)
            blocked[lit] = scene.trace(shadow[lit], tie_mode=0)["primID"] != orc.INVALID  # any hit in (0, maxT)
            weight = np.where(lit & ~blocked, np.maximum(f32(0), dot(ng, light_dir)) / np.where(pdf != 0, pdf, f32(1)), f32(0))
            est += weight[:, None] * le[None, :]
    final = np.where(shade[:, None], f32(1.0 / spp) * est * diffuse[face] / f32(np.pi), np.where(hit[:, None], emission[face], f32(0)))
    img = np.zeros((res * res, 4), np.uint8)
    img[:, :3] = ((final * f32(255)).astype(np.int64) & 0xFF).astype(np.uint8)  # float -> uint32 -> uint8 store, as the kernel does
    img[:, 3] = 255
    ok, bad, mx = H.image_matches(img.reshape(res, res, 4), H.load_golden_image("ShadowRayCornellBox"))
    assert ok and bad < 1e-4, (bad, mx)
