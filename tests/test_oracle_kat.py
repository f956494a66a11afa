"""Pins the oracle (oracle/oracle.cpp) against every known-answer vector that exists for this path.

The reference has no tests or golden images of its own; the vectors below are SURVEY.md 8c's
KAT-1/2/3, which were derived from the shader text, plus an independent pure-Python restatement of
shaders/common.glsl:13-37 in this file. CPU only.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

M32 = 0xFFFFFFFF


# -- independent pure-python restatement of common.glsl:13-37 (small cases only) -------------
def py_pcg(state):
    prev = (state * 747796405 + 2891336453) & M32
    word = (((prev >> ((prev >> 28) + 4)) ^ prev) * 277803737) & M32
    return prev, ((word >> 22) ^ word) & M32


def py_pcg2d(x, y):
    x = (x * 1664525 + 1013904223) & M32
    y = (y * 1664525 + 1013904223) & M32
    x = (x + y * 1664525) & M32
    y = (y + x * 1664525) & M32
    x ^= x >> 16
    y ^= y >> 16
    x = (x + y * 1664525) & M32
    y = (y + x * 1664525) & M32
    x ^= x >> 16
    y ^= y >> 16
    return x, y


def py_rand(u):
    return np.float32(u) * np.float32(2.0 ** -32)


# (pixel, k) -> (pcg2d.x, pcg2d.y, seed, [(pcg out, rand)...])   SURVEY.md 8c KAT-1
KAT1 = [
    ((1, 1), 1, 0x4EC53945, 0xF18B9C58, 0x4050D59D,
     [(0x31A081A4, 0.19385539), (0xECC5A0AF, 0.92489058), (0xFACBD399, 0.97967267), (0xD6799B21, 0.83779305)]),
    ((128, 128), 1, 0xDFEE247B, 0x4820EFCB, 0x280F1446,
     [(0x50547AB3, 0.31378904), (0xBB9FB1A8, 0.73290551), (0x358B3D17, 0.20915586), (0x2D1A630F, 0.17618388)]),
    ((255, 255), 1, 0x4FA4AA0E, 0xD27BE578, 0x22208F86,
     [(0x4C135DAC, 0.29717049), (0xEB538E7A, 0.91924369), (0x50DD3F0B, 0.31587595), (0x05AF2D62, 0.022204243)]),
    ((512, 384), 1, 0x4D22C801, 0x026720EA, 0x4F89E8EB,
     [(0x5A82A72A, 0.35355610), (0xE7507ACF, 0.90357178), (0x5A462E54, 0.35263339), (0x0684D4EE, 0.025464352)]),
    ((512, 384), 2, 0x2F235AC7, 0xE39BCC97, 0x12BF275E,
     [(0xB80E4F57, 0.71896833), (0x2F7690BA, 0.18540291), (0xAE608CEB, 0.68116075), (0x62B9DA86, 0.38564840)]),
    ((1023, 1023), 256, 0x72279616, 0x65E681D9, 0xD80E17EF,
     [(0x7285F2CC, 0.44735640), (0x2BB72C89, 0.17076376), (0xCB00A48B, 0.79297858), (0x9577AABE, 0.58385724)]),
    ((4095, 4095), 128, 0x903C722B, 0x982AFD28, 0x28676F53,
     [(0xE393CBBF, 0.88897395), (0xA50DA312, 0.64473933), (0x2812EA12, 0.15653861), (0x20CF2DD8, 0.12816130)]),
]


def test_kat1_rng_vectors():
    L = O.lib()
    for (px, py), k, ex, ey, eseed, outs in KAT1:
        x, y = C.c_uint32((px * k) & M32), C.c_uint32((py * k) & M32)
        L.orc_pcg2d(C.byref(x), C.byref(y))
        assert (x.value, y.value) == (ex, ey)
        assert py_pcg2d((px * k) & M32, (py * k) & M32) == (ex, ey)
        assert L.orc_seed(px, py, k) == eseed
        st = C.c_uint32(eseed)
        pst = eseed
        for word, r in outs:
            st2 = C.c_uint32(st.value)
            assert L.orc_pcg(C.byref(st2)) == word
            got = L.orc_rand(C.byref(st))
            pst, pword = py_pcg(pst)
            assert pword == word and pst == st.value
            assert got == py_rand(word)
            assert abs(got - r) < 5e-8


def test_kat1_pixel00_seed_is_constant():
    L = O.lib()
    x, y = C.c_uint32(0), C.c_uint32(0)
    L.orc_pcg2d(C.byref(x), C.byref(y))
    assert (x.value, y.value) == (0x18E431A7, 0x055DF4D1)
    for k in (1, 2, 33, 257, 4097):
        assert L.orc_seed(0, 0, k) == 0x1E422678  # T4: 0*k == 0


def test_rand_can_return_one():
    # T5: float(0xffffffff) rounds to 2^32 -> rand == 1.0 exactly; scale is exactly 2^-32
    assert py_rand(0xFFFFFFFF) == np.float32(1.0)
    assert py_rand(0xFFFFFF7F) == np.float32(0.99999994)
    L = O.lib()
    # random sweep: oracle rand == numpy restatement, bit for bit
    rng = np.random.default_rng(1)
    for s in rng.integers(0, 2 ** 32, 2000, dtype=np.uint64):
        st = C.c_uint32(int(s))
        got = L.orc_rand(C.byref(st))
        ps, w = py_pcg(int(s))
        assert st.value == ps and got == py_rand(w)


def test_kat3_geometry(cornell):
    verts, idx, faces = cornell
    _, _, _, meta = O.load_cornell_golden()
    assert verts.shape == (108, 3) and faces.shape == (36, 6)
    assert meta["shape_tris"] == [2, 2, 2, 2, 14, 12, 2]        # shifted names, T9
    assert np.array_equal(idx, np.arange(108, dtype=np.uint32))   # de-indexed, T8
    np.testing.assert_allclose(verts.min(0), [-1.02, -1.99, -1.04], atol=1e-6)
    np.testing.assert_allclose(verts.max(0), [1.00, 0.00, 0.99], atol=1e-6)
    np.testing.assert_allclose(verts[:3], [[-1.01, 0, 0.99], [1, 0, 0.99], [-0.99, 0, -1.04]], atol=1e-6)
    tri = verts.reshape(36, 3, 3)
    for a, b in ((20, 16), (21, 17), (32, 30), (33, 31)):        # exact duplicates, T9
        assert np.array_equal(tri[a], tri[b])
    # material table: tris 0-1 floor .. 34-35 light
    kd = faces[:, :3]
    np.testing.assert_allclose(kd[8], [0.63, 0.065, 0.05], atol=1e-6)
    np.testing.assert_allclose(kd[6], [0.14, 0.45, 0.091], atol=1e-6)
    np.testing.assert_allclose(kd[0], [0.725, 0.71, 0.68], atol=1e-6)
    assert np.count_nonzero(faces[:, 3:].any(axis=1)) == 2
    np.testing.assert_allclose(faces[34], [0.78, 0.78, 0.78, 17, 12, 4], atol=1e-6)
    np.testing.assert_allclose(faces[35], [0.78, 0.78, 0.78, 17, 12, 4], atol=1e-6)
    # shader normals (closesthit.rchit:43-48) in flipped space
    n = -np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    np.testing.assert_allclose(n[0], [0, -1, 0], atol=1e-6)
    np.testing.assert_allclose(n[2], [0, 1, 0], atol=1e-6)
    np.testing.assert_allclose(n[4], [0, 0, 1], atol=1e-6)
    np.testing.assert_allclose(n[6], [-1, 0, 0], atol=1e-6)
    np.testing.assert_allclose(n[8], [0.9998, -0.0151, 0.0099], atol=1e-4)
    np.testing.assert_allclose(n[9], [1.0, -0.005, 0.0], atol=1e-4)
    np.testing.assert_allclose(n[34], [0, 1, 0], atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_golden_fixture_matches_reference_loader(cornell):
    """The committed fixture is what the reference's vendored tinyobjloader produces right now."""
    verts, idx, faces = cornell
    rv, ri, rf, shapes = O.ref_load_obj("/root/reference/assets/CornellBox-Original.obj", "/root/reference/assets")
    assert np.array_equal(rv, verts) and np.array_equal(ri, idx) and np.array_equal(rf, faces)


def test_kat2_primary_rays(cornell_oracle):
    """jitter -> direction -> closest hit for named pixels (k = 1)."""
    p = O.default_params(256, 256, 1, 2)
    rays, seeds = O.generate_rays(p, 0)
    hits32 = cornell_oracle.intersect(rays, precision=32, brute=True)
    hits64 = cornell_oracle.intersect(rays, precision=64, brute=True)

    def at(x, y):
        return y * 256 + x

    r = rays[at(128, 128)]
    np.testing.assert_allclose(r[:4], [0, -1, 5, 0.001])
    np.testing.assert_allclose(r[4:7], [8.1714e-4, 1.9086e-3, -0.99999785], rtol=2e-4)
    h = hits64[at(128, 128)]
    assert h["prim"] == 30                                        # lowest id of the duplicate pair 30/32
    np.testing.assert_allclose([h["t"], h["u"], h["v"]], [5.07869, 0.82526, 0.06289], atol=2e-5)
    assert hits32[at(128, 128)]["prim"] == 30
    h = hits64[at(200, 200)]
    np.testing.assert_allclose(rays[at(200, 200)][4:7], [0.182575, 0.181322, -0.966327], atol=2e-6)
    assert h["prim"] == 6
    np.testing.assert_allclose([h["t"], h["u"], h["v"]], [5.47720, 0.36810, 0.00345], atol=2e-5)
    for (x, y) in ((0, 0), (128, 10), (128, 250), (10, 128), (245, 128)):
        assert hits64[at(x, y)]["prim"] == O.MISS and hits32[at(x, y)]["prim"] == O.MISS
    np.testing.assert_allclose(rays[at(0, 0)][4:7], [-0.301532, -0.300905, -0.904729], atol=2e-6)
    # f32 and f64 intersectors agree on the primitive for (almost) every primary ray
    assert np.mean(hits32["prim"] != hits64["prim"]) < 1e-3

    p = O.default_params(1024, 1024, 1, 2)
    rays, _ = O.generate_rays(p, 0)
    i = 384 * 1024 + 512
    np.testing.assert_allclose(rays[i][4:7], [2.29e-4, -0.082463, -0.996594], atol=2e-6)
    h = cornell_oracle.intersect(rays[i:i + 1], precision=64, brute=True)[0]
    assert h["prim"] == 5
    np.testing.assert_allclose([h["t"], h["u"], h["v"]], [6.06064, 0.50198, 0.25168], atol=2e-5)


def test_sky_pixels_exact(cornell_oracle):
    """Border pixels see only sky: radiance is exactly (0.7,0.6,0.5), alpha exactly 1 (SURVEY 8c framing fact)."""
    p = O.default_params(256, 256, 4, 8)
    img, rays = cornell_oracle.render(p, 32)
    sky = np.array([0.7, 0.6, 0.5, 1.0], np.float32)
    for (x, y) in ((0, 0), (128, 10), (128, 250), (10, 128), (245, 128)):
        np.testing.assert_allclose(img[y, x], sky, rtol=3e-7)
    assert 256 * 256 * 4 <= rays <= 256 * 256 * 4 * 8
    assert np.all(img[..., 3] == 1.0)
    assert np.isfinite(img).all()


def test_light_pixels_carry_emission(cornell_oracle):
    p = O.default_params(256, 256, 1, 1)
    rays, _ = O.generate_rays(p, 0)
    hits = cornell_oracle.intersect(rays, brute=True)
    img, _ = cornell_oracle.render(p, 32)
    lit = np.isin(hits["prim"], (34, 35)).reshape(256, 256)
    assert lit.sum() > 100
    np.testing.assert_allclose(img[lit][:, :3], np.tile([17, 12, 4], (lit.sum(), 1)), rtol=1e-6)
    # image is upright: the light (ceiling, flipped y = -1.98) is in the upper half
    ys = np.nonzero(lit)[0]
    assert ys.max() < 128


def test_f32_vs_f64_noise_floor(cornell_oracle):
    """Sets the floor of the parity metric: same algorithm, float vs double."""
    p = O.default_params(128, 128, 16, 8)
    a, _ = cornell_oracle.render(p, 32)
    b, _ = cornell_oracle.render(p, 64)
    e = O.rel_l2(a, b)
    assert e < 1e-3, e


def test_frame_split_invariance(cornell_oracle):
    """T3: 4 frames x 8 spp == 1 frame x 32 spp up to float summation order."""
    p = O.default_params(96, 96, 32, 4)
    one, _ = cornell_oracle.render(p, 32)
    img = np.zeros((96, 96, 4), np.float32)
    for f in range(4):
        q = O.default_params(96, 96, 8, 4, frame=f)
        cornell_oracle.render(q, 32, image=img)
    assert O.rel_l2(img, one) < 1e-5


def test_tile_invariance(cornell_oracle):
    """T3: rendering row tiles separately reproduces the full image bit for bit."""
    p = O.default_params(64, 64, 2, 4)
    full, _ = cornell_oracle.render(p, 32)
    img = np.zeros((64, 64, 4), np.float32)
    for y0 in range(0, 64, 16):
        q = O.default_params(64, 64, 2, 4, tile_y0=y0, tile_rows=16)
        cornell_oracle.render(q, 32, image=img)
    assert np.array_equal(img, full)


def test_bvh_equals_brute_force_on_soup():
    verts, idx, faces = O.soup(3000, 0x5EED0001)
    s = O.Scene(verts, idx, faces, brute_threshold=64)
    rng = np.random.default_rng(7)
    n = 4000
    o = rng.uniform([-1, -2, -1], [1, 0, 1], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, np.full((n, 1), 1e-3, np.float32), d, np.full((n, 1), 1e4, np.float32)], 1)
    a = s.intersect(rays, 32, brute=True)
    b = s.intersect(rays, 32, brute=False)
    assert np.array_equal(a, b)
    assert 0.05 < np.mean(a["prim"] != O.MISS) <= 1.0


def test_soup_definition():
    n = 1000
    verts, idx, faces = O.soup(n, 0x5EED0001)
    assert np.array_equal(idx, np.arange(3 * n, dtype=np.uint32))
    tri = verts.reshape(n, 3, 3)
    s = O.soup_scale(n)
    c = tri.mean(1)
    assert (c.min(0) > np.array([-1, -2, -1]) - s).all() and (c.max(0) < np.array([1, 0, 1]) + s).all()
    assert (np.abs(tri - c[:, None]).max() <= 2 * s + 1e-6)
    em = faces[:, 3:].any(1)
    assert np.array_equal(np.nonzero(em)[0], np.arange(0, n, 128))
    assert faces[:, :3].min() >= 0.1 and faces[:, :3].max() <= 0.9 + 1e-6
    # counter-based: value j of triangle i from the python restatement
    i, j = 17, 4
    st = (0x5EED0001 + (16 * i + j) * 0x9E3779B9) & M32
    _, w = py_pcg(st)
    f = [py_rand(py_pcg((0x5EED0001 + (16 * i + jj) * 0x9E3779B9) & M32)[1]) for jj in range(15)]
    cx = np.float32(f[0] * np.float32(2) - np.float32(1))
    off = np.float32((f[4] * np.float32(2) - np.float32(1)) * np.float32(s))
    cy = np.float32(f[1] * np.float32(2) - np.float32(2))
    assert tri[i, 0, 1] == np.float32(cy + off)
    off0 = np.float32((f[3] * np.float32(2) - np.float32(1)) * np.float32(s))
    assert tri[i, 0, 0] == np.float32(cx + off0)


def test_rgba8_mode_matches_quantised_float(cornell_oracle):
    """T2: the reference's image is unorm8; the emulation clamps and rounds every frame."""
    p = O.default_params(64, 64, 4, 4, accum_mode=1)
    img = np.zeros((64, 64, 4), np.float32)
    cornell_oracle.render(p, 32, image=img)
    q = O.default_params(64, 64, 4, 4)
    ref, _ = cornell_oracle.render(q, 32)
    want = np.rint(np.clip(ref, 0, 1) * np.float32(255)) / np.float32(255)
    np.testing.assert_array_equal(img, want.astype(np.float32))


def test_non_reference_estimators_estimate_the_same_integral(cornell_oracle):
    """The oracle's restatements of the opt-in estimators (include/bpt.h: cosine sampling, Russian roulette, next-event
    estimation with the balance heuristic) are what the CUDA path is compared with at equal seeds, so they are checked
    here against the reference's estimator: with the sky switched off (all light from the emitter) the 8 x 8-pixel block
    means of 2048 samples agree within Monte-Carlo noise, and next-event estimation is the less noisy of the two."""
    sky = (0.0, 0.0, 0.0)

    def frames(n, f0, **kw):
        out = []
        for f in range(f0, f0 + n):
            img = np.zeros((24, 24, 4), np.float32)
            p = O.default_params(24, 24, 256, 8, 0, sky=sky, **kw)
            p.frame = f
            cornell_oracle.render(p, 32, brute=True, image=img)
            out.append((img[..., :3] * (f + 1)).reshape(3, 8, 3, 8, 3).mean(axis=(1, 3)))   # undo the running mean
        return out
    truth = np.mean(frames(12, 0), axis=0)
    noise = {}
    for name, kw in (("reference", {}), ("nee", dict(nee=1)), ("nee+cosine+rr", dict(nee=1, sampler=1, rr_start_depth=3)),
                     ("rr", dict(rr_start_depth=2))):
        est = frames(6, 50, **kw)
        mean = np.mean(est, axis=0)
        noise[name] = float(np.mean([np.linalg.norm(e - mean) for e in est]) / np.linalg.norm(truth))
        assert np.linalg.norm(mean - truth) / np.linalg.norm(truth) < 0.06, name
    assert noise["nee"] < 0.8 * noise["reference"], noise
