#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own test data.

Run in the build container (needs /root/reference); the resulting small fixtures are
committed so that nothing reads /root/reference at test time.

Sources (reference file:line):
  * test/CornellBox.h:28-164      -> cornell_box.npz  (96 vertices, material ids, colours)
  * test/references/*.png          -> golden_images.npz (RGBA uint8, 512x512) -- the images the
    reference's unit tests compare against (test/hiprtTest.cpp:943-1002)
  * test/common/meshes/cornellbox/cornellBox.obj -> cornell_obj.npz (per-shape vertices/indices)
"""
import os
import re
import sys

import numpy as np

REF = os.environ.get("HIPRT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_cornell_header():
    txt = open(os.path.join(REF, "test/CornellBox.h")).read()
    body = txt[txt.index("cornellBoxVertices") : txt.index("cornellBoxMatIndices")]
    nums = re.findall(r"\{\s*(-?[\d.]+)f\s*,\s*(-?[\d.]+)f\s*,\s*(-?[\d.]+)f\s*\}", body)
    verts = np.array(nums, np.float32)
    assert verts.shape == (96, 3), verts.shape
    mat_body = txt[txt.index("cornellBoxMatIndices") : txt.index("cornellBoxEmissionColors")]
    mat_body = re.sub(r"//.*", "", mat_body)
    mats = np.array(re.findall(r"\b(\d+)\b", mat_body.split("{ {")[1].split("}")[0]), np.uint32)
    assert mats.shape == (32,), mats.shape
    dif_body = txt[txt.index("cornellBoxDiffuseColors") :]
    dif = np.array(re.findall(r"\{\s*([\d.]+)f\s*,\s*([\d.]+)f\s*,\s*([\d.]+)f\s*\}", dif_body), np.float32)
    assert dif.shape == (4, 3), dif.shape
    return verts, mats, dif


def parse_obj(path):
    """Minimal OBJ reader: per-'g'/'o' shape triangle lists, fan-triangulated, indices into a
    global vertex array (what tinyobjloader gives the reference tests)."""
    verts, shapes, cur, material = [], [], None, None
    for line in open(path):
        p = line.split()
        if not p:
            continue
        if p[0] == "usemtl":
            material = p[1]
            continue
        if p[0] == "v":
            verts.append([float(x) for x in p[1:4]])
        elif p[0] in ("g", "o"):
            cur = {"name": " ".join(p[1:]), "tris": [], "mats": []}
            shapes.append(cur)
        elif p[0] == "f":
            if cur is None:
                cur = {"name": "default", "tris": [], "mats": []}
                shapes.append(cur)
            ids = []
            for tok in p[1:]:
                i = int(tok.split("/")[0])
                ids.append(i - 1 if i > 0 else len(verts) + i)
            for k in range(1, len(ids) - 1):
                cur["tris"].append([ids[0], ids[k], ids[k + 1]])
                cur["mats"].append(material)
    shapes = [s for s in shapes if s["tris"]]
    return np.array(verts, np.float32), shapes


def main():
    from PIL import Image

    verts, mats, dif = parse_cornell_header()
    np.savez_compressed(os.path.join(HERE, "cornell_box.npz"), vertices=verts, mat_indices=mats, diffuse=dif)

    names = [
        "MinimumCornellBox", "MeshIntersection", "PairTriangles", "PrimIdsCornellBox", "HitDistCornellBox",
        "UvsCornellBox", "PrimaryRayCornellBox", "NormalsCornellBox", "SceneIntersection", "SceneIntersectionSingleton",
        "TranslateCornellBox", "ScaleCornellBox", "RotateCornellBox", "Cutout", "CustomIntersection", "MotionBlur", "Shear",
        "BvhUpdateCornellBox", "RayMaskCornellBox", "Rebuild0", "Rebuild1", "Update0", "Update1", "MotionBlurSlerp",
        "SceneIntersectionMlas", "AoRayCornellBox", "ShadowRayCornellBox",
        # byte-identical copies of images above (checked below): the reference renders them through other builders / the
        # import and save-load paths and validates against the same picture
        "BvhFastCornellBox", "BvhBalancedCornellBox", "BvhHighQCornellBox", "CustomBvhImport", "BvhIoApi",
    ]
    same_as = {"BvhFastCornellBox": "PrimaryRayCornellBox", "BvhBalancedCornellBox": "PrimaryRayCornellBox",
               "BvhHighQCornellBox": "PrimaryRayCornellBox", "CustomBvhImport": "MinimumCornellBox", "BvhIoApi": "MinimumCornellBox"}
    imgs = {}
    for n in names:
        im = np.array(Image.open(os.path.join(REF, "test/references", n + ".png")).convert("RGBA"))
        if n in same_as:
            assert np.array_equal(im, imgs[same_as[n]]), n  # recorded as an alias, not stored twice
            print(n, "==", same_as[n])
            continue
        imgs[n] = im
        print(n, im.shape)
    np.savez_compressed(os.path.join(HERE, "golden_images.npz"), **imgs)

    v, shapes = parse_obj(os.path.join(REF, "test/common/meshes/cornellbox/cornellBox.obj"))
    out = {"vertices": v, "shape_count": np.array(len(shapes))}
    # diffuse colour per face (cornellBox.mtl, Kd): what PrimaryRayKernel<VisualizeColor> paints (PrimaryRayKernel.h:43-56)
    kd, ke, name = {}, {}, None
    for line in open(os.path.join(REF, "test/common/meshes/cornellbox/cornellBox.mtl")):
        q = line.split()
        if q and q[0] == "newmtl":
            name = q[1]
        elif q and q[0] == "Kd":
            kd[name] = [float(x) for x in q[1:4]]
        elif q and q[0] == "Ke":
            ke[name] = [float(x) for x in q[1:4]]  # emission: Material::light() (test/shared.h:75) = "is an area light"
    for i, s in enumerate(shapes):
        out[f"tris_{i}"] = np.array(s["tris"], np.uint32)
        out[f"diffuse_{i}"] = np.array([kd[m] for m in s["mats"]], np.float32)
        out[f"emission_{i}"] = np.array([ke[m] for m in s["mats"]], np.float32)
        print("shape", i, s["name"], len(s["tris"]), sorted(set(s["mats"])))
    np.savez_compressed(os.path.join(HERE, "cornell_obj.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
