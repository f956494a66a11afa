"""GPU parity tests: the CUDA path, called through the C ABI (ctypes over libbpt.so), against the oracle.

Run on the B200 box: python -m pytest tests -m gpu. Tolerances are stated per test; integer work is bit-exact.
"""
import importlib

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
bpt = importlib.import_module("single-file-vulkan-pathtracing_b200")


@pytest.fixture(scope="module")
def pt_cornell(cornell):
    pt = bpt.PathTracer(0)
    pt.upload_mesh(*cornell)
    pt.build_accel()
    yield pt
    pt.close()


def random_rays(n, seed, lo=(-1.2, -2.2, -1.2), hi=(1.2, 0.2, 1.2)):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, np.full((n, 1), 1e-3, np.float32), d, np.full((n, 1), 1e4, np.float32)], 1)


def compare_hits(gpu, ref, tris_world, max_mismatch=2e-4, tol=2e-4):
    """prim ids must agree except for a tiny fraction of edge-grazing rays; where they agree t,u,v agree to `tol`
    (relative for t). Mismatches must at least agree on the hit distance (coincident/adjacent geometry)."""
    same = gpu["prim"] == ref["prim"]
    frac = 1.0 - same.mean()
    assert frac <= max_mismatch, f"{(~same).sum()} of {len(ref)} rays disagree on the primitive"
    hit = same & (ref["prim"] != O.MISS)
    # f32 Moeller-Trumbore against the f64 oracle: a ray that grazes its triangle (|d.n| tiny) amplifies the f32
    # rounding of the determinant, so up to 1e-4 of the hits may exceed `tol`, and none may exceed 10 * tol
    for name, rtol in (("t", tol), ("u", 0.0), ("v", 0.0)):
        g, r = gpu[name][hit].astype(np.float64), ref[name][hit].astype(np.float64)
        err = np.abs(g - r) - rtol * np.abs(r)
        assert (err > tol).mean() <= 1e-4, (name, (err > tol).sum(), len(err))
        assert err.max() <= 10 * tol, (name, err.max())
    return frac


# ------------------------------------------------------------------------------------------ K9
def test_generate_rays_bit_exact(pt_cornell):
    for (w, h, spp, frame, s) in ((256, 256, 1, 0, 0), (64, 48, 32, 3, 17), (1024, 8, 4, 1, 2)):
        p = bpt.default_params(w, h, spp, 8, frame)
        rays, seeds = pt_cornell.generate_rays(p, s)
        orays, oseeds = O.generate_rays(O.default_params(w, h, spp, 8, frame), s)
        assert np.array_equal(seeds, oseeds)
        assert np.array_equal(rays.view(np.uint32), orays.view(np.uint32)), np.abs(rays - orays).max()


def test_generate_rays_tile(pt_cornell):
    p = bpt.default_params(128, 128, 2, 8, 0, tile_y0=32, tile_rows=16)
    rays, seeds = pt_cornell.generate_rays(p, 1)
    orays, oseeds = O.generate_rays(O.default_params(128, 128, 2, 8, 0, tile_y0=32, tile_rows=16), 1)
    assert np.array_equal(seeds, oseeds) and np.array_equal(rays, orays)


# ------------------------------------------------------------------------------------------ K1-K7
def test_build_invariants_cornell(pt_cornell, cornell):
    check_build_invariants(pt_cornell, cornell[0], cornell[1])


def check_build_invariants(pt, verts, idx):
    info = pt.accel_info()
    n = info.num_tris
    tri = verts[idx].reshape(n, 3, 3)
    lo, hi = tri.min(1), tri.max(1)
    keys = pt.download_morton()
    # K2/K3: sorted, unique, low word is a permutation of the primitives
    assert np.all(keys[1:] > keys[:-1])
    prim_sorted = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    assert np.array_equal(np.sort(prim_sorted), np.arange(n, dtype=np.uint32))
    # K2: morton codes match the oracle's definition on the normalised AABB centre
    slo, shi = lo.min(0), hi.max(0)
    cen = (np.float32(0.5) * (lo + hi)).astype(np.float32)
    ext = (shi - slo).astype(np.float32)
    nrm = np.where(ext > 0, (cen - slo) / np.where(ext > 0, ext, 1), 0).astype(np.float32)
    L = O.lib()
    sample = np.random.default_rng(0).choice(n, min(n, 2000), replace=False)
    code_of = {int(k & np.uint64(0xFFFFFFFF)): int(k >> np.uint64(32)) for k in keys}
    for i in sample:
        assert code_of[int(i)] == L.orc_morton30(float(nrm[i, 0]), float(nrm[i, 1]), float(nrm[i, 2]))
    # K4/K5: binary hierarchy — every leaf reachable exactly once, child boxes inside parents
    left, right, aabbs = pt.download_lbvh()
    if n > 1:
        seen = np.zeros(2 * n - 1, np.int32)
        stack = [0]
        while stack:
            b = stack.pop()
            seen[b] += 1
            if b < n - 1:
                for c in (int(left[b]), int(right[b])):
                    assert np.all(aabbs[c, :3] >= aabbs[b, :3]) and np.all(aabbs[c, 3:] <= aabbs[b, 3:])
                    stack.append(c)
        assert np.all(seen == 1)
        leaf_boxes = aabbs[n - 1:]
        # leaf order = the sorted order, permuted only inside the small subtrees the SAH stage rebuilt
        order = pt.download_leaf_order()
        assert np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
        np.testing.assert_array_equal(leaf_boxes[:, :3], lo[order])
        np.testing.assert_array_equal(leaf_boxes[:, 3:], hi[order])
        np.testing.assert_array_equal(aabbs[0, :3], slo)
        np.testing.assert_array_equal(aabbs[0, 3:], shi)
    # K6: BVH8 — one array of 64-byte records; every triangle in exactly one triangle record; quantised child boxes
    # contain their triangles; children records are contiguous (internal children, then triangles in slot order)
    recs, woop, rec_prim, gbias, gstep = pt.download_accel()
    nrec = len(recs)
    is_tri = rec_prim != O.MISS
    assert is_tri.sum() == n and np.array_equal(np.sort(rec_prim[is_tri]), np.arange(n, dtype=np.uint32))
    assert info.num_nodes8 == nrec - n
    visited = np.zeros(nrec, np.int32)
    tri_seen = np.zeros(n, np.int32)
    depth = 0
    level = [(0, None, None)]
    gbias64, gstep64 = gbias.astype(np.float64), gstep.astype(np.float64)
    while level:
        depth += 1
        nxt = []
        for (ni, plo, phi) in level:
            assert not is_tri[ni]
            nd = recs[ni]
            visited[ni] += 1
            org = int(nd["org"])
            cell = np.array([org & 0x1FFFFF, (org >> 21) & 0x1FFFFF, (org >> 42) & 0x1FFFFF], np.float64)
            # the kernel decodes the origin with one f32 FMA per axis: fmaf(float(2^23 + c), grid_step, grid_bias)
            p = ((np.float64(8388608.0) + cell) * gstep64 + gbias64).astype(np.float32).astype(np.float64)
            ev = int(nd["e_valid"])
            scale = np.ldexp(1.0, (ev >> 24) - 127)
            counts, imask = ev & 0xFFFF, (ev >> 16) & 0xFF
            base = int(nd["child_base"])
            n_int = bin(imask).count("1")
            rank = 0
            leaf_rank = 0
            for s in range(8):
                inner = (imask >> s) & 1
                cnt = (counts >> (2 * s)) & 3
                if not inner and not cnt:
                    continue
                assert not (inner and cnt)
                qlo = np.array([nd["qlox"][s], nd["qloy"][s], nd["qloz"][s]], np.float64)
                qhi = np.array([nd["qhix"][s], nd["qhiy"][s], nd["qhiz"][s]], np.float64)
                blo, bhi = p + qlo * scale, p + qhi * scale
                if plo is not None:  # both boxes are conservative supersets of the same exact box, each on its own
                    # node's grid, so they nest only up to one quantisation step of this node (+ the padding)
                    assert np.all(blo >= plo - scale - 1e-4) and np.all(bhi <= phi + scale + 1e-4)
                if inner:
                    nxt.append((base + rank, blo, bhi))
                    rank += 1
                else:
                    for j in range(cnt):
                        rec = base + n_int + leaf_rank + j
                        assert is_tri[rec] and woop["prim"][rec] == rec_prim[rec]
                        prim = rec_prim[rec]
                        visited[rec] += 1
                        tri_seen[prim] += 1
                        # quantised box contains the triangle with the margin the traversal kernel relies on
                        assert np.all(blo <= lo[prim] - scale / 256) and np.all(bhi >= hi[prim] + scale / 256)
                    leaf_rank += cnt
        level = nxt
    assert np.all(visited == 1) and np.all(tri_seen == 1)
    assert depth == info.max_depth8
    # K7: Woop rows map v0,v1,v2 to (0,0,0),(1,0,0),(0,1,0)
    tri_recs = np.nonzero(is_tri)[0]
    t = tri[rec_prim[tri_recs]].astype(np.float64)
    area = np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1)
    ok = area > 1e-12
    rows = woop["rows"][tri_recs]
    M, c = rows[:, :, :3].astype(np.float64), rows[:, :, 3].astype(np.float64)
    # the rows are rounded once to f32, so the residual is bounded by a few f32 ulps of the magnitudes summed
    # (slivers have large rows: an absolute tolerance would measure the triangle's conditioning, not the kernel)
    eps = float(np.finfo(np.float32).eps)
    for k, want in enumerate(([0, 0, 0], [1, 0, 0], [0, 1, 0])):
        got = np.einsum("nij,nj->ni", M, t[:, k]) + c
        mag = np.einsum("nij,nj->ni", np.abs(M), np.abs(t[:, k])) + np.abs(c)
        assert np.all(np.abs(got - np.array(want, np.float64))[ok] <= 4 * eps * mag[ok] + 1e-6)


# ------------------------------------------------------------------------------------------ K10
def test_trace_primary_rays_cornell(pt_cornell, cornell_oracle, cornell):
    rays, _ = O.generate_rays(O.default_params(256, 256, 1, 2), 0)
    gpu = pt_cornell.trace_rays(rays)
    ref = cornell_oracle.intersect(rays, 64, brute=True)
    compare_hits(gpu, ref, None)
    # the Cornell box runs the STAGED instance (whole record array TMA-copied into shared memory); the global
    # instance must return the same hits bit for bit
    assert pt_cornell.accel_info().top_nodes_smem == pt_cornell.accel_info().num_records > 36
    pt_cornell.set_option(bpt.OPT_SMEM_TOP_NODES, 0)
    assert pt_cornell.accel_info().top_nodes_smem == 0
    assert np.array_equal(pt_cornell.trace_rays(rays), gpu)
    pt_cornell.set_option(bpt.OPT_SMEM_TOP_NODES, 1 << 20)
    # the staged instance tests 2 triangles per lane per iteration by default; the hits do not depend on that
    for n in (1, 3):
        pt_cornell.set_option(bpt.OPT_TRACE_STAGED_TRIS_PER_STEP, n)
        assert np.array_equal(pt_cornell.trace_rays(rays), gpu)
    pt_cornell.set_option(bpt.OPT_TRACE_STAGED_TRIS_PER_STEP, 2)
    # KAT-2 named pixels
    assert gpu[128 * 256 + 128]["prim"] == 30 and gpu[200 * 256 + 200]["prim"] == 6
    assert gpu[0]["prim"] == O.MISS


def test_trace_random_rays_cornell(pt_cornell, cornell_oracle):
    rays = random_rays(200_000, 3)
    gpu = pt_cornell.trace_rays(rays)
    ref = cornell_oracle.intersect(rays, 64, brute=True)
    compare_hits(gpu, ref, None)


def test_trace_ragged_counts(pt_cornell, cornell_oracle):
    """Ray counts that are not multiples of the warp / CTA size, including 1."""
    for n in (1, 31, 33, 1023, 1025, 4097):
        rays = random_rays(n, 100 + n)
        gpu = pt_cornell.trace_rays(rays)
        ref = cornell_oracle.intersect(rays, 64, brute=True)
        assert (gpu["prim"] != ref["prim"]).sum() <= 1
    assert len(pt_cornell.trace_rays(np.zeros((0, 8), np.float32))) == 0


def test_trace_axis_aligned_and_degenerate_rays(pt_cornell, cornell_oracle):
    """Directions with exact zeros (1/d = inf guarded) and rays starting on surfaces."""
    o = np.array([[0, -1, 5], [0, -1, 0.5], [0.3, -0.5, 0.2], [0, -1, 5]], np.float32)
    d = np.array([[0, 0, -1], [0, 1, 0], [1, 0, 0], [0, -0.0, -1]], np.float32)
    rays = np.concatenate([o, np.full((4, 1), 1e-3, np.float32), d, np.full((4, 1), 1e4, np.float32)], 1)
    gpu = pt_cornell.trace_rays(rays)
    ref = cornell_oracle.intersect(rays, 64, brute=True)
    assert np.array_equal(gpu["prim"], ref["prim"])
    np.testing.assert_allclose(gpu["t"], ref["t"], rtol=1e-5)


@pytest.fixture(scope="module")
def soup20k():
    verts, idx, faces = O.soup(20_000, 0x5EED0001)
    return verts, idx, faces, O.Scene(verts, idx, faces)


def test_soup_generator_bit_exact(soup20k):
    verts, idx, faces, _ = soup20k
    with bpt.PathTracer(0) as pt:
        pt.upload_soup(20_000, 0x5EED0001)
        gv, gi, gf = pt.download_mesh(20_000)
    assert np.array_equal(gi, idx)
    assert np.array_equal(gv.view(np.uint32), verts.view(np.uint32))
    assert np.array_equal(gf.view(np.uint32), faces.view(np.uint32))


def test_soup_build_and_trace(soup20k):
    verts, idx, faces, scene = soup20k
    with bpt.PathTracer(0) as pt:
        pt.upload_soup(20_000, 0x5EED0001)
        pt.build_accel()
        check_build_invariants(pt, verts, idx)
        rays = random_rays(100_000, 5)
        gpu = pt.trace_rays(rays)
        ref = scene.intersect(rays, 64)
        compare_hits(gpu, ref, None, max_mismatch=5e-4)
        # a 20 k-triangle BVH does not fit in shared memory: the global instance runs whatever the staging option says
        assert pt.accel_info().top_nodes_smem == 0
        pt.set_option(bpt.OPT_SMEM_TOP_NODES, 0)
        assert np.array_equal(pt.trace_rays(rays), gpu)
        pt.set_option(bpt.OPT_SMEM_TOP_NODES, 1 << 20)
        # instrumented kernel: same hits, plausible counters
        pt.set_option(bpt.OPT_COUNT_TRAVERSAL, 1)
        pt.reset_stats()
        gpu2 = pt.trace_rays(rays)
        st = pt.stats()
        assert np.array_equal(gpu2, gpu)
        assert st.rays_traced == len(rays) and st.nodes_visited >= len(rays) and st.tris_tested > 0


def test_tiny_meshes():
    """1, 2, 3, 4, 9 triangles (root-only trees, single leaf) incl. a degenerate triangle."""
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 4, 9):
        verts = rng.uniform(-1, 1, (3 * n, 3)).astype(np.float32)
        if n >= 3:
            verts[3:6] = verts[3]  # zero-area triangle
        idx = np.arange(3 * n, dtype=np.uint32)
        faces = np.tile(np.array([0.5, 0.5, 0.5, 0, 0, 0], np.float32), (n, 1))
        scene = O.Scene(verts, idx, faces)
        with bpt.PathTracer(0) as pt:
            pt.upload_mesh(verts, idx, faces)
            pt.build_accel()
            rays = random_rays(20_000, n, lo=(-2, -2, -2), hi=(2, 2, 2))
            gpu = pt.trace_rays(rays)
            ref = scene.intersect(rays, 64, brute=True)
            compare_hits(gpu, ref, None, max_mismatch=5e-4)


# ------------------------------------------------------------------------------------------ whole path
def test_cfg1_image_parity(pt_cornell, cornell_oracle):
    """BASELINE config 1: Cornell 256x256, 1 spp, depth 2. rel-L2 <= 1e-3 (north_star tolerance)."""
    pt_cornell.clear_image()
    pt_cornell.reset_stats()
    img = pt_cornell.render(bpt.default_params(256, 256, 1, 2))
    ref, rays = cornell_oracle.render(O.default_params(256, 256, 1, 2), 32)
    assert pt_cornell.stats().rays_traced == rays
    err = O.rel_l2(img, ref)
    assert err <= 1e-3, err
    assert np.array_equal(img[0, 0], np.array([0.7, 0.6, 0.5, 1.0], np.float32))
    # strict per-pixel view of the same thing: all but a handful of pixels agree to 1e-5
    bad = np.abs(img - ref).max(-1) > 1e-5 * (1 + np.abs(ref).max(-1))
    assert bad.mean() < 1e-3, bad.sum()


def test_reference_defaults_crop_parity(pt_cornell, cornell_oracle):
    """The reference's own constants (1024x1024, 32 spp, depth 8) on a 24-row tile; 2 frames."""
    pt_cornell.clear_image()
    ref = np.zeros((1024, 1024, 4), np.float32)
    for f in range(2):
        pt_cornell.trace(bpt.default_params(1024, 1024, 32, 8, f, tile_y0=500, tile_rows=24))
        cornell_oracle.render(O.default_params(1024, 1024, 32, 8, f, tile_y0=500, tile_rows=24), 32, image=ref)
    img = pt_cornell.read_image(1024, 1024)
    assert np.all(img[:500] == 0) and np.all(img[524:] == 0)
    err = O.rel_l2(img[500:524], ref[500:524])
    assert err <= 1e-3, err


def test_cfg2_downscaled_parity(pt_cornell, cornell_oracle):
    """Config 2's depth/spp structure at 128x128: 64 spp as 2 frames x 32, depth 8."""
    pt_cornell.clear_image()
    ref = np.zeros((128, 128, 4), np.float32)
    for f in range(2):
        pt_cornell.trace(bpt.default_params(128, 128, 32, 8, f))
        cornell_oracle.render(O.default_params(128, 128, 32, 8, f), 32, image=ref)
    img = pt_cornell.read_image(128, 128)
    err = O.rel_l2(img, ref)
    assert err <= 1e-3, err


def test_frame_split_and_tile_invariance_gpu(pt_cornell):
    """T3 on the device: tiles reproduce the full image bit for bit; 4x8 spp == 1x32 spp to float order."""
    pt_cornell.clear_image()
    full = pt_cornell.render(bpt.default_params(96, 96, 8, 6))
    pt_cornell.clear_image()
    for y0 in range(0, 96, 32):
        pt_cornell.trace(bpt.default_params(96, 96, 8, 6, tile_y0=y0, tile_rows=32))
    tiled = pt_cornell.read_image(96, 96)
    assert np.array_equal(full, tiled)
    pt_cornell.clear_image()
    one = pt_cornell.render(bpt.default_params(96, 96, 32, 6))
    pt_cornell.clear_image()
    four = pt_cornell.render(bpt.default_params(96, 96, 8, 6), frames=4)
    assert O.rel_l2(four, one) < 1e-5


def test_interleaved_tiling_gpu(pt_cornell, cornell_oracle):
    """Multi-GPU tiling on one GPU: the 4 interleaved tiles (8-row blocks dealt round-robin) are disjoint, each equals
    the same rows of the full image bit for bit, and the BGRA8 view is de-interleaved the same way."""
    pt_cornell.clear_image()
    full = pt_cornell.render(bpt.default_params(80, 96, 4, 6), frames=2)
    acc = np.zeros_like(full)
    for rank in range(4):
        tile = dict(tile_block=8, tile_nranks=4, tile_rank=rank)
        p = bpt.default_params(80, 96, 4, 6, **tile)
        part = pt_cornell.render(p, frames=2)
        rows = [y for y in range(96) if (y // 8) % 4 == rank]
        other = [y for y in range(96) if (y // 8) % 4 != rank]
        assert np.array_equal(part[rows], full[rows]) and np.all(part[other] == 0)
        rays, seeds = pt_cornell.generate_rays(p, 1)
        orays, oseeds = O.generate_rays(O.default_params(80, 96, 4, 6, **tile), 1)
        assert np.array_equal(seeds, oseeds) and np.array_equal(rays, orays)
        bgra = pt_cornell.read_image_bgra8(80, 96)
        assert np.all(bgra[other] == 0) and np.all(bgra[rows][..., 3] == 255)
        acc += part
    assert np.array_equal(acc, full)
    ref = np.zeros((96, 80, 4), np.float32)
    for f in range(2):
        cornell_oracle.render(O.default_params(80, 96, 4, 6, f, tile_block=8, tile_nranks=4, tile_rank=2), 32, image=ref)
    rows = [y for y in range(96) if (y // 8) % 4 == 2]
    pt_cornell.render(bpt.default_params(80, 96, 4, 6, tile_block=8, tile_nranks=4, tile_rank=2), frames=2)
    assert O.rel_l2(pt_cornell.read_image(80, 96)[rows], ref[rows]) <= 1e-3
    with pytest.raises(bpt.BptError):
        pt_cornell.trace(bpt.default_params(80, 96, 4, 6, tile_block=8, tile_nranks=5, tile_rank=0))   # 96 % 40 != 0
    with pytest.raises(bpt.BptError):
        pt_cornell.trace(bpt.default_params(80, 96, 4, 6, tile_block=8, tile_nranks=4, tile_rank=4))
    pt_cornell.clear_image()


def test_determinism(pt_cornell):
    pt_cornell.clear_image()
    a = pt_cornell.render(bpt.default_params(200, 120, 4, 8))
    pt_cornell.clear_image()
    b = pt_cornell.render(bpt.default_params(200, 120, 4, 8))
    assert np.array_equal(a, b)


def test_long_frame_loop_folds_its_events(cornell):
    """A caller that never asks for statistics (the reference's frame loop, main.cpp:647-685) must not pile up timing
    events: bpt_trace folds finished event pairs into the running sums once 256 are pending; the sums stay right and
    the image equals the one of a loop that reads the statistics every frame."""
    import time
    imgs = []
    for poll in (False, True):
        with bpt.PathTracer(0) as pt:
            pt.upload_mesh(*cornell)
            pt.build_accel()
            pt.set_option(bpt.OPT_PROFILE, 1)
            pt.reset_stats()
            t0 = time.perf_counter()
            for f in range(700):
                pt.trace(bpt.default_params(32, 32, 1, 3, f))
                if poll:
                    pt.stats()
            st = pt.stats()
            wall_ms = 1e3 * (time.perf_counter() - t0)
            assert st.paths == 700 * 32 * 32 and st.trace_launches >= 700
            assert 0.0 < st.trace_kernel_ms < st.frame_ms <= wall_ms
            imgs.append(pt.read_image(32, 32))
    assert np.array_equal(imgs[0], imgs[1])


def test_rgba8_mode(pt_cornell, cornell_oracle):
    """T2: unorm8 running mean, two frames, against the oracle's emulation; BGRA8 readback order."""
    pt_cornell.clear_image()
    ref = np.zeros((64, 64, 4), np.float32)
    for f in range(2):
        pt_cornell.trace(bpt.default_params(64, 64, 4, 4, f, accum_mode=bpt.ACCUM_RGBA8))
        cornell_oracle.render(O.default_params(64, 64, 4, 4, f, accum_mode=1), 32, image=ref)
    img = pt_cornell.read_image(64, 64)
    q = np.rint(img * 255).astype(np.int32)
    qr = np.rint(ref * 255).astype(np.int32)
    assert np.abs(q - qr).max() <= 1 and (q != qr).mean() < 2e-3
    bgra = pt_cornell.read_image_bgra8(64, 64)
    assert np.array_equal(bgra[..., 2], q[..., 0]) and np.array_equal(bgra[..., 0], q[..., 2])
    assert np.all(bgra[..., 3] == 255)


def test_soup_image_parity(soup20k):
    verts, idx, faces, scene = soup20k
    with bpt.PathTracer(0) as pt:
        pt.upload_mesh(verts, idx, faces)
        pt.build_accel()
        img = pt.render(bpt.default_params(128, 128, 4, 8))
        rays_gpu = pt.stats().rays_traced
    ref, rays = scene.render(O.default_params(128, 128, 4, 8), 32)
    err = O.rel_l2(img, ref)
    assert err <= 1e-3, err
    assert abs(rays_gpu - rays) <= 1e-4 * rays


# ------------------------------------------------------------------------------------------ K8: instanced scenes
def instance_grid(n_side=3, spacing=2.5, jitter=True, seed=3):
    """n_side^3 instance transforms (3x4 row-major): translations on a grid, plus a rotation about z and a uniform
    scale per instance when `jitter` (BASELINE config 5 is the plain grid)."""
    rng = np.random.default_rng(seed)
    xf = []
    c = 0.5 * (n_side - 1) * spacing
    for ix in range(n_side):
        for iy in range(n_side):
            for iz in range(n_side):
                a = rng.uniform(0, 2 * np.pi) if jitter else 0.0
                sc = rng.uniform(0.6, 1.3) if jitter else 1.0
                R = sc * np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
                t = np.array([ix * spacing - c, iy * spacing - c - 1.0, iz * spacing - c])
                xf.append(np.concatenate([R, t[:, None]], 1).reshape(12))
    return np.asarray(xf, np.float32)


@pytest.fixture(scope="module")
def pt_instanced(cornell):
    xf = instance_grid()
    pt = bpt.PathTracer(0)
    pt.upload_mesh(*cornell)
    pt.set_instances(xf)
    info = pt.build_accel()
    assert info.num_instances == len(xf) and info.num_tlas_nodes8 >= 1
    yield pt, xf, O.Scene(*cornell, xforms=xf)
    pt.close()


def test_instanced_trace_rays(pt_instanced, cornell):
    pt, xf, scene = pt_instanced
    ntris = len(cornell[1]) // 3
    rays = random_rays(200_000, 11, lo=(-5, -6, -5), hi=(5, 4, 5))
    gpu = pt.trace_rays(rays)
    ref = scene.intersect(rays, 64)
    assert (ref["prim"] != O.MISS).mean() > 0.3 and ref["prim"][ref["prim"] != O.MISS].max() >= 26 * ntris
    # the Cornell asset holds exact duplicate triangles: compare modulo the duplicates (T9)
    dup = {20: 16, 21: 17, 32: 30, 33: 31}
    def canon(p):
        q = p.copy()
        hit = q != O.MISS
        inst, tri = q[hit] // ntris, q[hit] % ntris
        for a, b in dup.items():
            tri[tri == a] = b
        q[hit] = inst * ntris + tri
        return q
    g2, r2 = gpu.copy(), ref.copy()
    g2["prim"], r2["prim"] = canon(gpu["prim"]), canon(ref["prim"])
    compare_hits(g2, r2, None, max_mismatch=5e-4)
    # staged (whole BVH in shared memory) and global instances of the kernel agree bit for bit
    pt.set_option(bpt.OPT_SMEM_TOP_NODES, 0)
    assert pt.accel_info().top_nodes_smem == 0
    assert np.array_equal(pt.trace_rays(rays), gpu)
    pt.set_option(bpt.OPT_SMEM_TOP_NODES, 1 << 20)
    assert pt.accel_info().top_nodes_smem > 0
    for n in (1, 4):  # triangle / instance-record tests per iteration of the staged two-level instance
        pt.set_option(bpt.OPT_TRACE_STAGED_TRIS_PER_STEP, n)
        assert np.array_equal(pt.trace_rays(rays), gpu)
    pt.set_option(bpt.OPT_TRACE_STAGED_TRIS_PER_STEP, 2)


def test_instanced_image_parity(pt_instanced):
    pt, xf, scene = pt_instanced
    kw = dict(cam_origin=(0.0, -1.0, 14.0), cam_target=(0.0, -1.0, 11.0))
    pt.clear_image()
    img = pt.render(bpt.default_params(160, 160, 4, 6, **kw))
    ref, rays = scene.render(O.default_params(160, 160, 4, 6, **kw), 32)
    assert O.rel_l2(img, ref) <= 1e-3
    assert (ref[..., :3] != np.array([0.7, 0.6, 0.5], np.float32)).any(axis=-1).mean() > 0.2  # the grid is in frame


def test_single_transformed_instance_matches_transformed_mesh(cornell):
    """One instance with a non-identity transform == the mesh with that transform baked into its vertices."""
    verts, idx, faces = cornell
    M = np.array([[0.8, 0, 0.6, 0.3], [0, 1, 0, -0.2], [-0.6, 0, 0.8, 0.1]], np.float32)
    baked = np.stack([M[r, 0] * verts[:, 0] + M[r, 1] * verts[:, 1] + M[r, 2] * verts[:, 2] + M[r, 3] for r in range(3)], 1).astype(np.float32)
    rays = random_rays(50_000, 12)
    with bpt.PathTracer(0) as a, bpt.PathTracer(0) as b:
        a.upload_mesh(verts, idx, faces); a.set_instances(M.reshape(1, 12)); a.build_accel()
        b.upload_mesh(baked, idx, faces); b.build_accel()
        ha, hb = a.trace_rays(rays), b.trace_rays(rays)
        same = ha["prim"] == hb["prim"]
        assert 1.0 - same.mean() <= 5e-4
        np.testing.assert_allclose(ha["t"][same], hb["t"][same], rtol=2e-4, atol=2e-4)
        pa, pb = bpt.default_params(96, 96, 2, 4), bpt.default_params(96, 96, 2, 4)
        assert O.rel_l2(a.render(pa), b.render(pb)) <= 1e-3


def test_moving_instances_rebuilds_only_the_instance_level(cornell):
    """new transforms + bpt_build_accel keep the mesh-level BVH (the reference's BLAS) and rebuild
    the instance level (its TLAS); the result equals a context built from scratch with the new transforms."""
    xf0, xf1 = instance_grid(seed=3), instance_grid(seed=4)
    rays = random_rays(100_000, 13, lo=(-5, -6, -5), hi=(5, 4, 5))
    with bpt.PathTracer(0) as a, bpt.PathTracer(0) as b:
        a.upload_mesh(*cornell); a.set_instances(xf0); a.build_accel()
        recs0 = a.download_accel()[0].copy()
        h0 = a.trace_rays(rays)
        a.set_instances(xf1); a.build_accel()                      # instances moved
        assert np.array_equal(a.download_accel()[0], recs0)        # mesh-level records untouched
        b.upload_mesh(*cornell); b.set_instances(xf1); b.build_accel()
        h1 = a.trace_rays(rays)
        assert np.array_equal(h1, b.trace_rays(rays)) and not np.array_equal(h1, h0)
        a.upload_mesh(*cornell)                                     # a new mesh resets to one identity instance
        a.build_accel()
        assert a.accel_info().num_instances == 1 and a.accel_info().num_tlas_nodes8 == 0


def test_instance_errors(cornell):
    with bpt.PathTracer(0) as pt:
        with pytest.raises(bpt.BptError):
            pt.set_instances(np.eye(3, 4, dtype=np.float32).reshape(1, 12))   # before upload
        pt.upload_mesh(*cornell)
        with pytest.raises(bpt.BptError):
            pt.set_instances(np.zeros((1, 12), np.float32))                      # singular
        pt.set_instances(np.eye(3, 4, dtype=np.float32).reshape(1, 12))
        with pytest.raises(bpt.BptError):
            pt.trace(bpt.default_params(8, 8, 1, 1))                             # instances invalidate the build
        pt.build_accel()
        pt.trace(bpt.default_params(8, 8, 1, 1))


def test_cosine_sampler_converges_to_the_parity_estimator(pt_cornell):
    """the opt-in cosine-weighted sampler is a different estimator of the same integral — not
    same-seed comparable with the reference, so it is validated by convergence: at 2048 spp both means agree within
    Monte-Carlo noise on a 48 x 48 image (averaged over 8 x 8 pixel blocks to cut the noise further)."""
    imgs = {}
    for name, sampler in (("uniform", bpt.SAMPLER_UNIFORM), ("cosine", bpt.SAMPLER_COSINE)):
        pt_cornell.clear_image()
        img = pt_cornell.render(bpt.default_params(48, 48, 256, 8, sampler=sampler), frames=8)[..., :3]
        imgs[name] = img.reshape(6, 8, 6, 8, 3).mean(axis=(1, 3))
    pt_cornell.clear_image()
    u, c = imgs["uniform"], imgs["cosine"]
    assert np.isfinite(c).all() and np.linalg.norm(u - c) / np.linalg.norm(u) < 0.08
    # same seeds, different estimator: the images must NOT be identical (the parity mode is the uniform sampler)
    assert not np.array_equal(u, c)


def test_graph_replay_is_bit_identical(pt_cornell):
    """BPT_OPT_USE_GRAPH: a frame's launch list captured once as a CUDA graph and replayed with only the frame index
    changing gives the same images as launching every kernel, across frames, parameter changes (re-capture) and
    option changes; Cfg1 (256 x 256, 1 spp, depth 2) is the launch-bound case it exists for."""
    def run(graph):
        pt_cornell.set_option(bpt.OPT_USE_GRAPH, graph)
        out = []
        for (w, h, spp, depth, frames) in ((256, 256, 1, 2, 5), (64, 48, 3, 6, 3), (256, 256, 1, 2, 2)):
            pt_cornell.clear_image()
            out.append(pt_cornell.render(bpt.default_params(w, h, spp, depth), frames=frames).copy())
        pt_cornell.set_option(bpt.OPT_USE_GRAPH, 0)
        return out
    plain, replay = run(0), run(1)
    for a, b in zip(plain, replay):
        assert np.array_equal(a, b)
    # the wavefront frame loop: generate, 2 x (traverse, shade), gather, accumulate = 7 kernels in the graph + the
    # frame setter; the fused path kernel (opt-in): path kernel, gather, accumulate + the frame setter
    for fused, traces, kernels in ((0, 2, 7 + 1), (1, 1, 3 + 1)):
        pt_cornell.set_option(bpt.OPT_FUSED_PATHS, fused)
        pt_cornell.reset_stats()
        pt_cornell.set_option(bpt.OPT_USE_GRAPH, 1)
        pt_cornell.clear_image()
        img = pt_cornell.render(bpt.default_params(256, 256, 1, 2), frames=4).copy()
        st = pt_cornell.stats()
        pt_cornell.set_option(bpt.OPT_USE_GRAPH, 0)
        pt_cornell.clear_image()
        assert st.trace_launches == 4 * traces and st.kernel_launches == 4 * kernels
        assert np.array_equal(img, pt_cornell.render(bpt.default_params(256, 256, 1, 2), frames=4))
        pt_cornell.clear_image()
    pt_cornell.set_option(bpt.OPT_FUSED_PATHS, 0)


def test_error_paths(cornell):
    verts, idx, faces = cornell
    with bpt.PathTracer(0) as pt:
        with pytest.raises(bpt.BptError):
            pt.trace(bpt.default_params(8, 8, 1, 1))            # trace before build
        with pytest.raises(bpt.BptError):
            pt.build_accel()                                     # build before upload
        with pytest.raises(bpt.BptError):
            pt.upload_mesh(verts, idx[:-1], faces)               # ragged index buffer
        bad = idx.copy(); bad[5] = 10_000
        with pytest.raises(bpt.BptError):
            pt.upload_mesh(verts, bad, faces)                    # index out of range
        pt.upload_mesh(verts, idx, faces)
        pt.build_accel()
        for kw in (dict(spp=0), dict(depth=0), dict(width=0), dict(tile_y0=9), dict(frame=-1)):
            args = dict(width=8, height=8, spp=1, depth=1)
            args.update(kw)
            with pytest.raises(bpt.BptError):
                pt.trace(bpt.default_params(**args))
    with pytest.raises(bpt.BptError):
        bpt.PathTracer(99)


# ------------------------------------------------------------------------------------------ K11 alone (stage level)
def test_shade_step_matches_oracle(pt_cornell, cornell_oracle, soup20k):
    """bpt_shade_step = closesthit.rchit:50-65 / miss.rmiss:8-12 + raygen.rgen:76-83 for a batch of paths, nothing else:
    an error that cancels in an image mean (rand order, weight update, emission before the break) shows here per path.
    Integer work (seeds, alive flags) and everything made of + - * / sqrt is bit-exact (shade.cu is built without FMA
    contraction, the oracle too); only sinf/cosf may differ from the host libm in the last bits -> direction 1e-6
    absolute; the weight multiplies dot(direction, normal), so it inherits that absolute error (2e-5 on weights of
    order 1) next to 2e-6 relative."""
    def check(pt, scene, rays, sampler):
        hits = pt.trace_rays(rays)                       # (t, u, v) as shading re-derives them from the vertices
        rng = np.random.default_rng(17)
        w = rng.uniform(0.05, 1.5, (len(rays), 3)).astype(np.float32)
        seed = rng.integers(0, 2 ** 32, len(rays), dtype=np.uint32)
        seed[:4] = (0, 1, 0xFFFFFFFF, 0x80000000)
        g = pt.shade_step(bpt.default_params(64, 64, 1, 8, sampler=sampler), rays, hits, w, seed)
        r = scene.shade(O.default_params(64, 64, 1, 8, sampler=sampler), hits, w, seed)
        assert np.array_equal(g["alive"], r["alive"]) and 0.2 < g["alive"].mean() < 1.0
        assert np.array_equal(g["contrib"].view(np.uint32), r["contrib"].view(np.uint32))
        assert (g["contrib"] != 0).any(axis=1).sum() > 100            # sky on misses, Ke on the light
        a = g["alive"].astype(bool)
        assert np.array_equal(g["seed"][a], r["seed"][a])             # two rands consumed, in order (r1 then r2)
        assert np.array_equal(g["ray"][a][:, :4].view(np.uint32), r["ray"][a][:, :4].view(np.uint32))   # position, tmin
        assert np.array_equal(g["ray"][a][:, 7], r["ray"][a][:, 7])                                       # tmax
        assert np.abs(g["ray"][a][:, 4:7] - r["ray"][a][:, 4:7]).max() <= 1e-6
        np.testing.assert_allclose(g["weight"][a], r["weight"][a], rtol=2e-6, atol=2e-5)
        assert (g["ray"][~a] == 0).all() and (g["weight"][~a] == 0).all()
        return a.sum()

    for sampler in (bpt.SAMPLER_UNIFORM, bpt.SAMPLER_COSINE):
        assert check(pt_cornell, cornell_oracle, random_rays(50_000, 21), sampler) > 20_000
    verts, idx, faces, scene = soup20k
    with bpt.PathTracer(0) as pt:
        pt.upload_mesh(verts, idx, faces)
        pt.build_accel()
        check(pt, scene, random_rays(50_000, 22), bpt.SAMPLER_UNIFORM)
        # ragged sizes incl. 1 path
        for n in (1, 31, 257):
            check_small = pt.shade_step(bpt.default_params(8, 8, 1, 8), random_rays(n, n), pt.trace_rays(random_rays(n, n)),
                                        np.ones((n, 3), np.float32), np.arange(n, dtype=np.uint32))
            assert len(check_small["alive"]) == n


# ------------------------------------------------------------------------------------------ the reference's shader text
def golden(name):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    return g["image"], int(g["rays"])


def test_cfg1_against_the_reference_shader_text(pt_cornell):
    """BASELINE config 1 (256 x 256, 1 spp, depth 2) against the image the reference's OWN shader text produces
    (tests/golden/ref_shade_cfg1_*.npz: shaders/*.glsl|rgen|rchit|rmiss compiled as C++, oracle/ref_shade_glue.cpp).
    rel-L2 <= 1e-3 (north_star); same number of traceRayEXT calls."""
    ref, rays = golden("ref_shade_cfg1_256x256_1spp_depth2")
    pt_cornell.clear_image(); pt_cornell.reset_stats()
    img = pt_cornell.render(bpt.default_params(256, 256, 1, 2))
    assert pt_cornell.stats().rays_traced == rays
    assert O.rel_l2(img, ref) <= 1e-3
    bad = np.abs(img - ref).max(-1) > 1e-5 * (1 + np.abs(ref).max(-1))
    assert bad.mean() < 1e-3, bad.sum()


def test_cfg2_full_size_rows_against_the_reference_shader_text(pt_cornell):
    """BASELINE config 2 at its stated size: 1024 x 1024, 256 spp (8 frames of the text's 32), depth 8 — rows 508..515 of
    the full launch against the reference's shader text (golden), rel-L2 <= 1e-3."""
    ref, rays = golden("ref_shade_cfg2_1024x1024_rows508_516_8frames")
    pt_cornell.clear_image(); pt_cornell.reset_stats()
    for f in range(8):
        pt_cornell.trace(bpt.default_params(1024, 1024, 32, 8, f, tile_y0=508, tile_rows=8))
    img = pt_cornell.read_image(1024, 1024)[508:516]
    st = pt_cornell.stats()
    assert st.paths == 8 * 1024 * 256 and abs(st.rays_traced - rays) <= 1e-4 * rays
    err = O.rel_l2(img, ref)
    assert err <= 1e-3, err
    pt_cornell.clear_image()


def test_text_defaults_rgba8_against_the_reference_shader_text(pt_cornell):
    """The reference as it ships — 32 spp, depth 8, rgba8 storage image, three frames — at 64 x 64: quantised levels
    equal the shader text's except where a float sits on a rounding boundary."""
    ref, _ = golden("ref_shade_text_64x64_3frames_rgba8")
    pt_cornell.clear_image()
    for f in range(3):
        pt_cornell.trace(bpt.default_params(64, 64, 32, 8, f, accum_mode=bpt.ACCUM_RGBA8))
    q = np.rint(pt_cornell.read_image(64, 64) * 255).astype(np.int32)
    qr = np.rint(ref * 255).astype(np.int32)
    assert np.abs(q - qr).max() <= 1 and (q != qr).mean() < 2e-3
    pt_cornell.clear_image()


# ------------------------------------------------------------------------------------------ lanes, async read-back
def test_sample_lanes_do_not_change_the_image(pt_cornell):
    """BPT_OPT_STREAMS: the samples of a pass run as 1..4 independent wavefronts on their own streams; the image is
    bit-identical (per-path colours are folded in sample order), also under graph replay and with odd sample counts."""
    imgs = []
    for lanes in (1, 2, 3, 4):
        pt_cornell.set_option(bpt.OPT_STREAMS, lanes)
        pt_cornell.clear_image(); pt_cornell.reset_stats()
        imgs.append(pt_cornell.render(bpt.default_params(160, 96, 7, 6), frames=2).copy())
        st = pt_cornell.stats()
        assert st.paths == 2 * 160 * 96 * 7
        assert st.trace_launches == 2 * 6 * min(lanes, 7)
    for im in imgs[1:]:
        assert np.array_equal(im, imgs[0])
    pt_cornell.set_option(bpt.OPT_STREAMS, 3)
    pt_cornell.set_option(bpt.OPT_USE_GRAPH, 1)
    pt_cornell.clear_image()
    assert np.array_equal(pt_cornell.render(bpt.default_params(160, 96, 7, 6), frames=2), imgs[0])
    pt_cornell.set_option(bpt.OPT_USE_GRAPH, 0)
    pt_cornell.set_option(bpt.OPT_STREAMS, 1)
    with pytest.raises(bpt.BptError):
        pt_cornell.set_option(bpt.OPT_STREAMS, 5)
    pt_cornell.clear_image()


def test_async_read_back_overlaps_the_next_frame(pt_cornell):
    """bpt_read_image_async: frame f is copied out while frame f+1 is traced; every copy equals the synchronous read of
    the same frame, in float and BGRA8, with and without interleaved tiling."""
    import torch
    for tile in ({}, dict(tile_block=8, tile_nranks=2, tile_rank=1)):
        p = lambda f: bpt.default_params(128, 96, 4, 6, f, **tile)
        pt_cornell.clear_image()
        want, want8 = [], []
        for f in range(4):
            pt_cornell.trace(p(f))
            want.append(pt_cornell.read_image(128, 96).copy())
            want8.append(pt_cornell.read_image_bgra8(128, 96).copy())
        pt_cornell.clear_image()
        bufs = [torch.empty((96, 128, 4), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
        bufs8 = [torch.empty((96, 128, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
        pt_cornell.trace(p(0))
        for f in range(4):
            pt_cornell.read_image_async(bufs[f & 1])
            if f + 1 < 4:
                pt_cornell.trace(p(f + 1))          # enqueued while the copy of frame f is in flight
            pt_cornell.read_wait()
            assert np.array_equal(bufs[f & 1], want[f]), f
        pt_cornell.clear_image()
        pt_cornell.trace(p(0))
        for f in range(4):
            pt_cornell.read_image_bgra8_async(bufs8[f & 1])
            if f + 1 < 4:
                pt_cornell.trace(p(f + 1))
            pt_cornell.read_wait()
            assert np.array_equal(bufs8[f & 1], want8[f]), f
    pt_cornell.read_wait()                           # nothing pending: no-op
    pt_cornell.clear_image()


def test_russian_roulette_is_the_same_estimator(pt_cornell, cornell_oracle):
    """Russian roulette (bpt_params.rr_start_depth) is opt-in and not the reference's estimator. With
    the same seeds the CUDA path and the oracle's restatement of the SAME rule agree to 1e-3 (it is ordinary shading
    arithmetic + one more rand per bounce); against the reference's estimator it is validated by convergence: both
    means agree within Monte-Carlo noise at 2048 spp, and it traces fewer rays."""
    kw = dict(rr_start_depth=2)
    pt_cornell.clear_image(); pt_cornell.reset_stats()
    img = pt_cornell.render(bpt.default_params(96, 96, 8, 8, **kw))
    rays_rr = pt_cornell.stats().rays_traced
    ref, rays = cornell_oracle.render(O.default_params(96, 96, 8, 8, **kw), 32)
    assert O.rel_l2(img, ref) <= 1e-3 and abs(rays_rr - rays) <= 1e-4 * rays
    pt_cornell.clear_image(); pt_cornell.reset_stats()
    pt_cornell.render(bpt.default_params(96, 96, 8, 8))
    assert rays_rr < 0.9 * pt_cornell.stats().rays_traced
    means = {}
    for name, extra in (("reference", {}), ("rr", dict(rr_start_depth=3))):
        pt_cornell.clear_image()
        im = pt_cornell.render(bpt.default_params(48, 48, 256, 8, **extra), frames=8)[..., :3]
        means[name] = im.reshape(6, 8, 6, 8, 3).mean(axis=(1, 3))
    pt_cornell.clear_image()
    a, b = means["reference"], means["rr"]
    assert np.isfinite(b).all() and np.linalg.norm(a - b) / np.linalg.norm(a) < 0.08 and not np.array_equal(a, b)
    with pytest.raises(bpt.BptError):
        pt_cornell.trace(bpt.default_params(8, 8, 1, 2, rr_start_depth=1000))


def test_next_event_estimation(pt_cornell, cornell_oracle, soup20k):
    """next-event estimation with the balance heuristic (bpt_params.nee) is opt-in and not the
    reference's estimator. (1) With the same seeds the CUDA wavefront (k_nee -> shadow-ray traversal -> k_nee_resolve ->
    k_shade) and the oracle's restatement of the SAME rule agree to 1e-3, on the Cornell box and on a soup with 157
    emitters, for both samplers and with roulette. (2) Against the reference's estimator it is validated by convergence,
    and with the sky switched off (all light comes from the small emitter) its noise is well below the reference's."""
    for kw in (dict(nee=1), dict(nee=1, sampler=bpt.SAMPLER_COSINE, rr_start_depth=3)):
        pt_cornell.clear_image(); pt_cornell.reset_stats()
        img = pt_cornell.render(bpt.default_params(96, 96, 8, 8, **kw))
        rays_gpu = pt_cornell.stats().rays_traced
        ref, rays = cornell_oracle.render(O.default_params(96, 96, 8, 8, **kw), 32)
        assert O.rel_l2(img, ref) <= 1e-3, kw
        assert abs(rays_gpu - rays) <= 2e-4 * rays
    verts, idx, faces, scene = soup20k
    with bpt.PathTracer(0) as pt:
        pt.upload_mesh(verts, idx, faces)
        pt.build_accel()
        img = pt.render(bpt.default_params(96, 96, 4, 6, nee=1))
        ref, _ = scene.render(O.default_params(96, 96, 4, 6, nee=1), 32)
        assert O.rel_l2(img, ref) <= 2e-3
    # instanced scene: 27 rotated / scaled / translated boxes, 54 lights
    xf = instance_grid()
    cam = dict(cam_origin=(0.0, -1.0, 14.0), cam_target=(0.0, -1.0, 11.0))
    with bpt.PathTracer(0) as pt:
        pt.upload_mesh(*[cornell_oracle.verts, cornell_oracle.indices, cornell_oracle.faces])
        pt.set_instances(xf)
        pt.build_accel()
        img = pt.render(bpt.default_params(128, 128, 4, 6, nee=1, **cam))
    ref, _ = O.Scene(cornell_oracle.verts, cornell_oracle.indices, cornell_oracle.faces, xforms=xf).render(
        O.default_params(128, 128, 4, 6, nee=1, **cam), 32)
    assert O.rel_l2(img, ref) <= 2e-3
    dark = dict(sky=(0.0, 0.0, 0.0))
    def blocks(frames, f0, **kw):
        pt_cornell.clear_image()
        acc = np.zeros((48, 48, 3))
        for f in range(f0, f0 + frames):                              # independent frames (no running mean): sum by hand
            pt_cornell.clear_image()
            p = bpt.default_params(48, 48, 256, 8, 0, **kw)
            p.frame = f
            acc += pt_cornell.render(p)[..., :3] * (f + 1)
        return (acc / frames).reshape(6, 8, 6, 8, 3).mean(axis=(1, 3))
    truth = blocks(16, 0, **dark)                                     # 4096 spp of the reference's estimator
    est = {name: [blocks(1, 100 + k, **dark, **kw) for k in range(6)] for name, kw in (("reference", {}), ("nee", dict(nee=1)))}
    noise = {name: np.mean([np.linalg.norm(e - np.mean(v, axis=0)) for e in v]) for name, v in est.items()}
    assert np.linalg.norm(np.mean(est["nee"], axis=0) - truth) / np.linalg.norm(truth) < 0.05
    assert noise["nee"] < 0.75 * noise["reference"], noise
    pt_cornell.clear_image()


def test_device_side_scene_front_end(cornell, pt_cornell):
    """bpt_upload_obj_arrays does on the device what the body of the reference's loadFromFile does on
    the host (main.cpp:37-57: de-index, negate Y, per-face Kd/Ke) from the RAW arrays the reference's own tinyobjloader
    returns for the shipped asset (fixture fields raw_*, generated by oracle/_ref/libref_loader.so). The resulting
    buffers equal the reference loader's output bit for bit, and so does the image."""
    pos, corner, fmat, mats = O.load_cornell_raw_golden()
    verts, idx, faces = cornell
    with bpt.PathTracer(0) as pt:
        pt.upload_obj_arrays(pos, corner, fmat, mats)
        gv, gi, gf = pt.download_mesh(len(fmat))
        assert np.array_equal(gi, idx)
        assert np.array_equal(gv.view(np.uint32), verts.view(np.uint32))
        assert np.array_equal(gf.view(np.uint32), faces.view(np.uint32))
        pt.build_accel()
        img = pt.render(bpt.default_params(64, 64, 2, 4))
        pt_cornell.clear_image()
        assert np.array_equal(img, pt_cornell.render(bpt.default_params(64, 64, 2, 4)))
        pt_cornell.clear_image()
        bad = corner.copy(); bad[7] = 1000
        with pytest.raises(bpt.BptError):
            pt.upload_obj_arrays(pos, bad, fmat, mats)                  # a corner that names no position
        badm = fmat.copy(); badm[3] = -1
        with pytest.raises(bpt.BptError):
            pt.upload_obj_arrays(pos, corner, badm, mats)               # a face without a material (main.cpp:49-51)
        with pytest.raises(bpt.BptError):
            pt.trace(bpt.default_params(8, 8, 1, 1))                    # a failed upload leaves no scene behind


def test_sah_subtree_stage(soup20k):
    """the build's quality stage (BPT_OPT_BVH_SAH_SUBTREE, sah.cu) rebuilds every LBVH subtree of at
    most 32 triangles with the surface-area heuristic. It may only permute leaves inside such a subtree (a window of
    at most 32 positions of the Morton order), every structural invariant of the hierarchy and of the BVH8 still holds,
    the closest hits are the same, and the traversal visits fewer nodes."""
    verts, idx, faces, scene = soup20k
    rays = random_rays(100_000, 31)
    stats = {}
    hits = {}
    for size in (0, 32, 8):
        with bpt.PathTracer(0) as pt:
            pt.set_option(bpt.OPT_BVH_SAH_SUBTREE, size)
            pt.upload_mesh(verts, idx, faces)
            pt.build_accel()
            check_build_invariants(pt, verts, idx)
            keys = pt.download_morton()
            sorted_prims = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            order = pt.download_leaf_order()
            if size == 0:
                assert np.array_equal(order, sorted_prims)
            else:
                pos_sorted = np.empty(len(order), np.int64); pos_sorted[sorted_prims] = np.arange(len(order))
                moved = np.abs(pos_sorted[order] - np.arange(len(order)))
                assert moved.max() < size and (moved > 0).mean() > 0.3
            pt.set_option(bpt.OPT_COUNT_TRAVERSAL, 1)
            pt.reset_stats()
            hits[size] = pt.trace_rays(rays)
            st = pt.stats()
            stats[size] = (st.nodes_visited / len(rays), st.tris_tested / len(rays))
            with pytest.raises(bpt.BptError):
                pt.set_option(bpt.OPT_BVH_SAH_SUBTREE, 64)
    ref = scene.intersect(rays, 64)
    for size in hits:
        compare_hits(hits[size], ref, None, max_mismatch=5e-4)
    print("nodes, triangles per ray by SAH subtree size:", stats)
    assert stats[32][0] < stats[0][0] and stats[32][0] + stats[32][1] < stats[0][0] + stats[0][1]


def test_hostile_meshes_do_not_break_the_build(cornell):
    """Inputs the reference would hand to the driver without a check: triangles with NaN / infinite vertices, all
    triangles identical, all triangles degenerate. The build (Morton keys, SAH ranks, collapse) must terminate with a
    structure that still answers rays for the sane triangles, and must never hang or fault."""
    verts, idx, faces = cornell
    rays = random_rays(20_000, 41)
    v = verts.copy()
    v[3 * 5:3 * 5 + 3] = np.nan                      # triangle 5: NaN
    v[3 * 9] = np.inf                                # triangle 9: one infinite vertex
    with bpt.PathTracer(0) as pt:
        pt.upload_mesh(v, idx, faces)
        pt.build_accel()
        h = pt.trace_rays(rays)
        sane = O.Scene(np.delete(v.reshape(-1, 3, 3), [5, 9], axis=0).reshape(-1, 3), np.arange(3 * 34, dtype=np.uint32),
                       np.delete(faces, [5, 9], axis=0))
        ref = sane.intersect(rays, 64, brute=True)
        hit = ref["prim"] != O.MISS
        # every ray the sane triangles stop is stopped at the same distance (the broken ones can only add or hide nothing finite)
        close = np.abs(h["t"][hit] - ref["t"][hit]) <= 2e-4 * (1 + ref["t"][hit])
        assert close.mean() > 0.98
        img = pt.render(bpt.default_params(32, 32, 2, 4))
        assert img.shape == (32, 32, 4)
    n = 40
    same = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (n, 1))
    f = np.tile(np.array([0.5, 0.5, 0.5, 0, 0, 0], np.float32), (n, 1))
    for vv in (same, np.zeros_like(same)):           # 40 identical triangles; 40 degenerate ones
        with bpt.PathTracer(0) as pt:
            pt.upload_mesh(vv, np.arange(3 * n, dtype=np.uint32), f)
            pt.build_accel()
            if vv is same:   # (a scene that is a single point has no extent to put a quantisation margin around)
                check_build_invariants(pt, vv, np.arange(3 * n, dtype=np.uint32))
            r = np.array([[0.2, 0.2, 1, 1e-3, 0, 0, -1, 1e4]], np.float32)
            hh = pt.trace_rays(r)
            if vv is same:
                assert hh["prim"][0] == 0 and abs(hh["t"][0] - 1.0) < 1e-5   # exact duplicates: the lowest id wins
            else:
                assert hh["prim"][0] == O.MISS


def test_flat_and_rescaled_scenes(cornell):
    """Scenes whose extent is degenerate along an axis (a single quad: the scene grid and every node are flat there) and
    the Cornell box at 1000x and 1/1000 of its size: closest hits against the oracle's brute force."""
    quad = np.array([[-1, 0, -1], [1, 0, -1], [1, 0, 1], [-1, 0, -1], [1, 0, 1], [-1, 0, 1]], np.float32)
    f = np.tile(np.array([0.5, 0.5, 0.5, 0, 0, 0], np.float32), (2, 1))
    rng = np.random.default_rng(51)
    n = 50_000
    o = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(0.2, 2.0, n) * rng.choice([-1, 1], n), rng.uniform(-1.5, 1.5, n)], 1)
    tgt = np.stack([rng.uniform(-1.2, 1.2, n), np.zeros(n), rng.uniform(-1.2, 1.2, n)], 1)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, np.full((n, 1), 1e-3), d, np.full((n, 1), 1e4)], 1).astype(np.float32)
    for axis in range(3):                              # the quad flat in y, then the same quad flat in z and in x
        perm = np.roll(np.arange(3), axis)
        v = quad[:, perm]
        r = rays.copy(); r[:, 0:3] = rays[:, 0:3][:, perm]; r[:, 4:7] = rays[:, 4:7][:, perm]
        with bpt.PathTracer(0) as pt:
            pt.upload_mesh(v, np.arange(6, dtype=np.uint32), f)
            pt.build_accel()
            check_build_invariants(pt, v, np.arange(6, dtype=np.uint32))
            gpu = pt.trace_rays(r)
        ref = O.Scene(v, np.arange(6, dtype=np.uint32), f).intersect(r, 64, brute=True)
        assert (ref["prim"] != O.MISS).mean() > 0.6
        compare_hits(gpu, ref, None, max_mismatch=5e-4)
    verts, idx, faces = cornell
    for scale in (1000.0, 0.001):
        v = (verts * np.float32(scale)).astype(np.float32)
        r = random_rays(50_000, 52)
        r[:, 0:3] *= scale
        r[:, 3] = 1e-3 * scale; r[:, 7] = 1e4 * scale
        with bpt.PathTracer(0) as pt:
            pt.upload_mesh(v, idx, faces)
            pt.build_accel()
            gpu = pt.trace_rays(r)
        ref = O.Scene(v, idx, faces).intersect(r, 64, brute=True)
        same = gpu["prim"] == ref["prim"]
        assert 1.0 - same.mean() <= 5e-4, scale
        hit = same & (ref["prim"] != O.MISS)
        np.testing.assert_allclose(gpu["t"][hit], ref["t"][hit], rtol=3e-4)


# ------------------------------------------------------------------------------------------ fused path kernel
def _both_frame_loops(pt, params, frames=1, **opts):
    """The same frames through the per-bounce wavefront and through the fused path kernel; returns both images and
    both statistics."""
    out = []
    for fused in (0, 1):
        pt.set_option(bpt.OPT_FUSED_PATHS, fused)
        pt.clear_image(); pt.reset_stats()
        img = pt.render(params, frames=frames).copy()
        out.append((img, pt.stats()))
    pt.set_option(bpt.OPT_FUSED_PATHS, 0)
    pt.clear_image()
    return out


def test_fused_path_kernel_is_bit_identical_to_the_wavefront(pt_cornell, soup20k):
    """BPT_OPT_FUSED_PATHS: one path kernel per sample pass (primary rays, traversal, closest-hit / miss and the bounce
    inside the SMs, path state in shared memory) runs the same paths with the same arithmetic as generate + per-bounce
    traverse + shade: images are bit-identical, ray counts equal, and a pass is ONE traversal launch. Cornell box =
    the instance with the records staged in shared memory; the soup and SMEM_TOP_NODES = 0 = the global-memory instance."""
    cases = [dict(width=256, height=256, spp=1, depth=2),                       # Cfg1
             dict(width=160, height=96, spp=7, depth=6),
             dict(width=61, height=37, spp=32, depth=8),                        # ragged: fewer paths than slots in places
             dict(width=128, height=128, spp=4, depth=8, tile_y0=40, tile_rows=24),
             dict(width=128, height=128, spp=4, depth=8, tile_block=8, tile_nranks=4, tile_rank=1),
             dict(width=96, height=96, spp=8, depth=8, accum_mode=bpt.ACCUM_RGBA8),
             dict(width=96, height=96, spp=8, depth=5, sampler=bpt.SAMPLER_COSINE),
             dict(width=64, height=64, spp=3, depth=1),                         # primary rays only
             dict(width=8, height=4, spp=1, depth=3)]                           # fewer paths than one warp has slots
    for staged in (1, 0):
        pt_cornell.set_option(bpt.OPT_SMEM_TOP_NODES, (1 << 20) if staged else 0)
        for kw in cases:
            kw = dict(kw)
            p = bpt.default_params(kw.pop("width"), kw.pop("height"), kw.pop("spp"), kw.pop("depth"), **kw)
            (wi, ws), (fi, fs) = _both_frame_loops(pt_cornell, p, frames=2)
            assert np.array_equal(wi.view(np.uint32), fi.view(np.uint32)), (staged, kw, np.abs(wi - fi).max())
            assert ws.rays_traced == fs.rays_traced and ws.paths == fs.paths
            assert fs.trace_launches == 2 and ws.trace_launches == 2 * p.max_depth
    pt_cornell.set_option(bpt.OPT_SMEM_TOP_NODES, 1 << 20)
    # several passes per frame (pass size below the frame's paths), and the instrumented kernel: same node / triangle counts
    pt_cornell.set_option(bpt.OPT_PASS_PATHS, 64 * 64 * 3)
    pt_cornell.set_option(bpt.OPT_COUNT_TRAVERSAL, 1)
    (wi, ws), (fi, fs) = _both_frame_loops(pt_cornell, bpt.default_params(64, 64, 8, 6), frames=2)
    pt_cornell.set_option(bpt.OPT_COUNT_TRAVERSAL, 0)
    pt_cornell.set_option(bpt.OPT_PASS_PATHS, 1 << 27)
    assert np.array_equal(wi, fi) and fs.trace_launches == 2 * 3
    assert (ws.nodes_visited, ws.tris_tested, ws.rays_traced) == (fs.nodes_visited, fs.tris_tested, fs.rays_traced)
    # the soup: global-memory records, deep stacks, emissive triangles
    verts, idx, faces, scene = soup20k
    with bpt.PathTracer(0) as pt:
        pt.upload_mesh(verts, idx, faces)
        pt.build_accel()
        for (w, h, spp, depth) in ((128, 128, 4, 8), (333, 77, 2, 8)):
            (wi, ws), (fi, fs) = _both_frame_loops(pt, bpt.default_params(w, h, spp, depth), frames=2)
            assert np.array_equal(wi.view(np.uint32), fi.view(np.uint32)), np.abs(wi - fi).max()
            assert ws.rays_traced == fs.rays_traced
        # the estimators the path kernel does not know stay with the wavefront, whatever the option says
        pt.set_option(bpt.OPT_FUSED_PATHS, 1)
        pt.reset_stats()
        pt.render(bpt.default_params(64, 64, 2, 4, nee=1))
        assert pt.stats().trace_launches > 1


def test_fused_path_kernel_on_instanced_scenes(pt_instanced, cornell):
    """Two-level scenes through the path kernel: the instance sentinel reloads the world-space ray from the path slot.
    27 instances are staged in shared memory next to the path slots; with staging off the records come from global memory."""
    pt, xf, scene = pt_instanced
    kw = dict(cam_origin=(0.0, -1.0, 14.0), cam_target=(0.0, -1.0, 11.0))
    for staged in (1, 0):
        pt.set_option(bpt.OPT_SMEM_TOP_NODES, (1 << 20) if staged else 0)
        (wi, ws), (fi, fs) = _both_frame_loops(pt, bpt.default_params(160, 160, 4, 6, **kw), frames=2)
        assert np.array_equal(wi.view(np.uint32), fi.view(np.uint32)), (staged, np.abs(wi - fi).max())
        assert ws.rays_traced == fs.rays_traced and fs.trace_launches == 2
    pt.set_option(bpt.OPT_SMEM_TOP_NODES, 1 << 20)
