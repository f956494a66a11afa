"""Generates tests/golden/cornell_scene.json by running the REFERENCE's own vendored tinyobjloader
(oracle/_ref/libref_loader.so, compiled from /root/reference/external/tinyobjloader where it lies) on
the reference's asset with the loader semantics of main.cpp:28-58. Run in the build container only:

    python tests/golden/make_golden.py

Floats are stored as C99 hex literals so the fixture is bit-exact.
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
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cornell_scene.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, out["nverts"], "verts", out["ntris"], "tris", shape_tris, out["bbox_min"], out["bbox_max"])


if __name__ == "__main__":
    main()
