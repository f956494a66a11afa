"""Generates tests/golden/cornell_scene.json by running the REFERENCE's own vendored tinyobjloader
(oracle/_ref/libref_loader.so, compiled from /root/reference/external/tinyobjloader where it lies) on
the reference's asset with the loader semantics of main.cpp:28-58. Run in the build container only:

    python tests/golden/make_golden.py

Floats are stored as C99 hex literals so the fixture is bit-exact.

Also generates the PIXEL goldens from the reference's own shader text compiled as C++ (oracle/_ref/libref_shade.so,
see oracle/ref_shade_glue.cpp): tests/golden/ref_shade_*.npz (float32 images, bit-exact) and the scalar constants of
the shipped SPIR-V binaries (tests/golden/spv_constants.json, decoded by oracle/spv_constants.py).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import oracle_lib  # noqa: E402

REF = "/root/reference"


def main():
    oracle_lib.build()
    verts, idx, faces, shape_tris = oracle_lib.ref_load_obj(os.path.join(REF, "assets/CornellBox-Original.obj"),
                                                           os.path.join(REF, "assets"))
    out = {
        "source": "reference tinyobjloader v2.0.0-rc.9 on assets/CornellBox-Original.obj, loader semantics main.cpp:28-58",
        "nverts": int(len(verts)), "ntris": int(len(faces)), "shape_tris": shape_tris,
        "indices": [int(i) for i in idx],
        "verts_hex": [float(x).hex() for x in verts.reshape(-1)],
        "faces_hex": [float(x).hex() for x in faces.reshape(-1)],
        "bbox_min": [float(x) for x in verts.min(0)], "bbox_max": [float(x) for x in verts.max(0)],
    }
    pos, corner, fmat, mats = oracle_lib.ref_load_obj_raw(os.path.join(REF, "assets/CornellBox-Original.obj"),
                                                         os.path.join(REF, "assets"))
    out.update({"raw_positions_hex": [float(x).hex() for x in pos.reshape(-1)], "raw_corner_vertex": [int(i) for i in corner],
                "raw_face_material": [int(i) for i in fmat], "raw_materials_hex": [float(x).hex() for x in mats.reshape(-1)]})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cornell_scene.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, out["nverts"], "verts", out["ntris"], "tris", shape_tris, out["bbox_min"], out["bbox_max"])
    pixel_goldens(verts, idx, faces)


# name -> arguments of oracle_lib.ref_shade_render; every case names the BASELINE.json config it stands for
PIXEL_CASES = {
    # configs[0]: 256x256, 1 spp, depth 2 (the two loop bounds of raygen.rgen:43,62 overridden, rule R7 of glsl_to_cpp.py)
    "ref_shade_cfg1_256x256_1spp_depth2": dict(width=256, height=256, frames=1, spp=1, depth=2),
    # the shader text with NOTHING overridden (32 spp, depth 8), two frames of the running mean, float image
    "ref_shade_text_64x64_2frames": dict(width=64, height=64, frames=2),
    # the same with the reference's own rgba8 storage image (raygen.rgen:7, main.cpp:481-484), three frames
    "ref_shade_text_64x64_3frames_rgba8": dict(width=64, height=64, frames=3, rgba8=True),
    # configs[1] (1024x1024, 256 spp = 8 frames x 32, depth 8): rows 508..515 of the full-size launch
    "ref_shade_cfg2_1024x1024_rows508_516_8frames": dict(width=1024, height=1024, frames=8, rows=(508, 516)),
}


def pixel_goldens(verts, idx, faces):
    here = os.path.dirname(os.path.abspath(__file__))
    for name, kw in PIXEL_CASES.items():
        img, rays = oracle_lib.ref_shade_render(verts, idx, faces, **kw)
        r0, r1 = kw.get("rows", (0, 0))
        if r1:
            img = img[r0:r1]
        np.savez_compressed(os.path.join(here, name + ".npz"), image=img, rays=np.uint64(rays))
        print("wrote", name, img.shape, rays, "rays")
    sys.path.insert(0, os.path.join(os.path.dirname(here), "..", "oracle"))
    import spv_constants
    spv = {n: spv_constants.constants(os.path.join(REF, "shaders", n)) for n in
           ("raygen.rgen.spv", "closesthit.rchit.spv", "miss.rmiss.spv")}
    with open(os.path.join(here, "spv_constants.json"), "w") as f:
        json.dump(spv, f, indent=1)
    print("wrote spv_constants.json")


if __name__ == "__main__":
    main()
