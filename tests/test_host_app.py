"""The C++20 host application (single-file-vulkan-pathtracing_b200/host/main.cpp): built by __graft_entry__.build(),
fails loudly without a GPU, and on a GPU reproduces the oracle's image through the same C ABI the reference would bind."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "single-file-vulkan-pathtracing_b200", "lib", "bpt_host")


def write_scene_bin(path, verts, idx, faces):
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(idx, np.uint32).reshape(-1)
    faces = np.ascontiguousarray(faces, np.float32).reshape(-1, 6)
    with open(path, "wb") as f:
        f.write(b"BPTSCN1\0" + struct.pack("<3I", len(verts), len(idx), len(faces)))
        f.write(verts.tobytes()); f.write(idx.tobytes()); f.write(faces.tobytes())


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = map(int, f.readline().split())
        assert float(f.readline()) < 0
        return np.frombuffer(f.read(), "<f4").reshape(h, w, 3)[::-1]


def test_host_app_is_built_and_fails_loudly_without_gpu(tmp_path, cornell):
    import torch
    assert os.path.exists(HOST), "run __graft_entry__.build()"
    assert "usage" in subprocess.run([HOST, "--help"], capture_output=True, text=True).stdout
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    scene = tmp_path / "cornell.bin"
    write_scene_bin(scene, *cornell)
    r = subprocess.run([HOST, "--scene", str(scene), "--frames", "1", "--width", "8", "--height", "8", "--spp", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    assert "108 vertices, 36 triangles" in r.stdout


@pytest.mark.gpu
def test_host_app_renders_the_oracle_image(tmp_path, cornell, cornell_oracle):
    scene = tmp_path / "cornell.bin"
    write_scene_bin(scene, *cornell)
    out = tmp_path / "img"
    r = subprocess.run([HOST, "--scene", str(scene), "--frames", "3", "--width", "96", "--height", "96", "--spp", "4",
                        "--depth", "5", "--out", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    img = read_pfm(str(out) + ".pfm")
    ref = np.zeros((96, 96, 4), np.float32)
    for frame in range(3):  # the oracle applies raygen.rgen:86-90 to `ref` frame by frame
        cornell_oracle.render(O.default_params(96, 96, 4, 5, frame), 32, image=ref)
    assert O.rel_l2(img, ref[..., :3]) <= 1e-3
    ppm = open(str(out) + ".ppm", "rb").read()
    assert ppm.startswith(b"P6\n96 96\n255\n") and len(ppm) == len(b"P6\n96 96\n255\n") + 96 * 96 * 3
