"""Pins oracle/oracle.cpp against the reference's OWN shader text.

oracle/_ref/libref_shade.so is shaders/common.glsl, raygen.rgen, closesthit.rchit and miss.rmiss of the reference,
read where they lie and compiled as C++ (oracle/glsl_to_cpp.py + oracle/glsl_shim.h + oracle/ref_shade_glue.cpp).
 * `live` tests run wherever that library exists (the build container; it also travels to the GPU box) and compare
   the oracle with it BIT FOR BIT: RNG, sampling, whole images in float and rgba8 mode, small and big scenes.
 * `golden` tests run everywhere: the oracle against tests/golden/ref_shade_*.npz, images that library produced
   (generator: tests/golden/make_golden.py), again bit for bit.
The only thing neither side takes from the reference is the driver's traversal behind traceRayEXT (closed source);
both use the same closest-hit contract (raygen.rgen:63-75, main.cpp:525). CPU only.
"""
import json
import os
import struct

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
live = pytest.mark.skipif(not O.ref_shade_available(), reason="oracle/_ref/libref_shade.so not built (needs /root/reference)")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def oracle_render(scene, w, h, frames, spp=32, depth=8, mode=0, brute=True, rows=None):
    img = np.zeros((h, w, 4), np.float32)
    rays = 0
    kw = dict(tile_y0=rows[0], tile_rows=rows[1] - rows[0]) if rows else {}
    for f in range(frames):
        _, r = scene.render(O.default_params(w, h, spp, depth, f, accum_mode=mode, **kw), 32, brute=brute, image=img)
        rays += r
    return img, rays


# ------------------------------------------------------------------------------------------ golden (everywhere)
GOLDEN = {
    "ref_shade_cfg1_256x256_1spp_depth2": dict(w=256, h=256, frames=1, spp=1, depth=2),
    "ref_shade_text_64x64_2frames": dict(w=64, h=64, frames=2),
    "ref_shade_text_64x64_3frames_rgba8": dict(w=64, h=64, frames=3, mode=1),
    "ref_shade_cfg2_1024x1024_rows508_516_8frames": dict(w=1024, h=1024, frames=8, rows=(508, 516)),
}


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_equals_the_shader_text_golden(name, cornell_oracle):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    kw = GOLDEN[name]
    img, rays = oracle_render(cornell_oracle, **kw)
    if "rows" in kw:
        img = img[kw["rows"][0]:kw["rows"][1]]
    assert rays == int(g["rays"])
    assert np.array_equal(bits(img), bits(g["image"])), np.abs(img - g["image"]).max()


def test_spirv_constants_are_the_constants_of_the_text():
    """The reference executes the .spv files (main.cpp:541-543): their scalar constants (decoded by
    oracle/spv_constants.py, committed as tests/golden/spv_constants.json) are the ones the text and the oracle use."""
    spv = json.load(open(os.path.join(HERE, "golden", "spv_constants.json")))
    f = lambda x: struct.unpack("<I", struct.pack("<f", x))[0]
    rg = spv["raygen.rgen.spv"]
    two_pi = np.float32(2) * np.float32(3.14159265358979323846)
    inv_two_pi = np.float32(1) / two_pi
    for v in (0.001, 10000.0, 5.0, -1.0, 2.0, 1.0, float(two_pi), float(inv_two_pi), 2.0 ** -32, 3.14159265358979323846):
        assert f(v) in rg["f32"], v
    assert f(6.2831855) == f(float(two_pi)) and f(0.15915494) == f(float(inv_two_pi))   # the oracle's literals
    for v in (32, 8, 255, 1664525, 1013904223, 747796405, 2891336453, 277803737, 16, 22, 28, 4):
        assert v in rg["u32"], v
    assert f(3.1415927) in spv["closesthit.rchit.spv"]["f32"]
    for v in (0.7, 0.6, 0.5):
        assert f(v) in spv["miss.rmiss.spv"]["f32"]
    if os.path.isdir("/root/reference/shaders"):
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
        import spv_constants
        for name, want in spv.items():
            assert spv_constants.constants(os.path.join("/root/reference/shaders", name)) == want


# ------------------------------------------------------------------------------------------ live (library present)
@live
def test_live_rng_and_sampling_bit_exact():
    import ctypes as C
    R, L = O.ref_shade_lib(), O.lib()
    rng = np.random.default_rng(5)
    for s0 in [0, 1, 0xFFFFFFFF, 0x80000000] + [int(x) for x in rng.integers(0, 2 ** 32, 200)]:
        a, b = C.c_uint32(s0), C.c_uint32(s0)
        assert R.ref_pcg(C.byref(a)) == L.orc_pcg(C.byref(b)) and a.value == b.value
        a, b = C.c_uint32(s0), C.c_uint32(s0)
        ra, rb = R.ref_rand(C.byref(a)), L.orc_rand(C.byref(b))
        assert struct.pack("<f", ra) == struct.pack("<f", rb) and 0.0 <= ra <= 1.0
        x1, y1, x2, y2 = C.c_uint32(s0), C.c_uint32(s0 ^ 0x1234567), C.c_uint32(s0), C.c_uint32(s0 ^ 0x1234567)
        R.ref_pcg2d(C.byref(x1), C.byref(y1)); L.orc_pcg2d(C.byref(x2), C.byref(y2))
        assert (x1.value, y1.value) == (x2.value, y2.value)
    a = C.c_uint32(0)  # T5: rand() can return exactly 1.0
    vals = set()
    for _ in range(8):
        vals.add(R.ref_rand(C.byref(a)))
    assert all(0.0 <= v <= 1.0 for v in vals)


@live
def test_live_images_bit_exact_cornell(cornell, cornell_oracle):
    verts, idx, faces = cornell
    for (w, h, frames, spp, depth, rgba8) in ((64, 64, 2, 0, 0, False), (48, 80, 3, 0, 0, True), (256, 256, 1, 1, 2, False),
                                              (96, 32, 2, 5, 3, False), (33, 17, 1, 2, 8, True)):
        img, rays = O.ref_shade_render(verts, idx, faces, w, h, frames, spp, depth, rgba8)
        ref, orays = oracle_render(cornell_oracle, w, h, frames, spp or 32, depth or 8, int(rgba8))
        assert rays == orays, (w, h)
        assert np.array_equal(bits(img), bits(ref)), (w, h, np.abs(img - ref).max())
    # traceRayEXT answered by the oracle's intersector instead of the library's own brute-force loop: same image
    img2, _ = O.ref_shade_render(verts, idx, faces, 64, 64, 2, scene=cornell_oracle, brute=True)
    img1, _ = O.ref_shade_render(verts, idx, faces, 64, 64, 2)
    assert np.array_equal(bits(img1), bits(img2))
    one, _ = O.ref_shade_render(verts, idx, faces, 64, 64, 1, spp=1)
    assert np.array_equal(one[0, 0], np.array([0.7, 0.6, 0.5, 1.0], np.float32))   # KAT-2: pixel (0,0) sees the sky


@live
def test_live_images_bit_exact_soup_through_the_oracle_bvh():
    """A scene the brute-force loop cannot render in seconds: the shader text traces through the oracle's BVH (callback),
    so the two renders differ only where the restatement of the shader text differs from the text. Unmodified text:
    32 spp, depth 8; 20 k-triangle soup (every 128th triangle emissive)."""
    verts, idx, faces = O.soup(20_000, 0x5EED0001)
    scene = O.Scene(verts, idx, faces)
    img, rays = O.ref_shade_render(verts, idx, faces, 40, 40, 2, scene=scene, brute=False)
    ref, orays = oracle_render(scene, 40, 40, 2, brute=False)
    assert rays == orays and np.array_equal(bits(img), bits(ref))
    assert (img[..., :3] != np.array([0.7, 0.6, 0.5], np.float32)).any(-1).mean() > 0.3


@live
def test_live_row_subset_is_a_subset_of_the_launch(cornell):
    verts, idx, faces = cornell
    full, _ = O.ref_shade_render(verts, idx, faces, 64, 64, 1)
    part, _ = O.ref_shade_render(verts, idx, faces, 64, 64, 1, rows=(20, 28))
    assert np.array_equal(part[20:28], full[20:28]) and np.all(part[:20] == 0) and np.all(part[28:] == 0)
