"""ctypes binding of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_LOADER_PATH = os.path.join(ORACLE_DIR, "_ref", "libref_loader.so")
REF_SHADE_PATH = os.path.join(ORACLE_DIR, "_ref", "libref_shade.so")
MISS = 0xFFFFFFFF


class OrcParams(C.Structure):
    """Mirror of orc_params / bpt_params (include/bpt.h)."""
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("spp_per_frame", C.c_uint32), ("max_depth", C.c_uint32),
        ("frame", C.c_int32), ("tile_y0", C.c_uint32), ("tile_rows", C.c_uint32),
        ("cam_origin", C.c_float * 3), ("cam_target", C.c_float * 3), ("sky", C.c_float * 3),
        ("tmin", C.c_float), ("tmax", C.c_float), ("accum_mode", C.c_uint32), ("sampler", C.c_uint32),
        ("tile_block", C.c_uint32), ("tile_nranks", C.c_uint32), ("tile_rank", C.c_uint32),
        ("rr_start_depth", C.c_uint32), ("nee", C.c_uint32),
    ]


def default_params(width=1024, height=1024, spp=32, depth=8, frame=0, **kw):
    """The reference's compile-time constants (main.cpp:16-17, raygen.rgen:43,55-56,62,71,73, miss.rmiss:10)."""
    p = OrcParams()
    p.width, p.height, p.spp_per_frame, p.max_depth, p.frame = width, height, spp, depth, frame
    p.tile_y0, p.tile_rows = 0, 0
    p.cam_origin[:] = (0.0, -1.0, 5.0)
    p.cam_target[:] = (0.0, -1.0, 2.0)
    p.sky[:] = (0.7, 0.6, 0.5)
    p.tmin, p.tmax = 0.001, 10000.0
    p.accum_mode, p.sampler = 0, 0
    p.tile_block, p.tile_nranks, p.tile_rank = 0, 0, 0
    p.rr_start_depth, p.nee = 0, 0
    for k, v in kw.items():
        if k in ("cam_origin", "cam_target", "sky"):
            getattr(p, k)[:] = v
        else:
            setattr(p, k, v)
    return p


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "oracle.cpp")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    if os.path.isdir("/root/reference") and not (os.path.exists(REF_LOADER_PATH) and os.path.exists(REF_SHADE_PATH)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "_ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        u32p, f32p, vp = C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.c_void_p
        L.orc_pcg.restype = C.c_uint32; L.orc_pcg.argtypes = [u32p]
        L.orc_pcg2d.restype = None; L.orc_pcg2d.argtypes = [u32p, u32p]
        L.orc_seed.restype = C.c_uint32; L.orc_seed.argtypes = [C.c_uint32] * 3
        L.orc_rand.restype = C.c_float; L.orc_rand.argtypes = [u32p]
        L.orc_scene_create.restype = vp
        L.orc_scene_create.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32]
        L.orc_scene_destroy.restype = None; L.orc_scene_destroy.argtypes = [vp]
        L.orc_scene_ntris.restype = C.c_uint32; L.orc_scene_ntris.argtypes = [vp]
        L.orc_render.restype = C.c_uint64
        L.orc_render.argtypes = [vp, C.POINTER(OrcParams), C.c_int, C.c_int, C.c_int, vp]
        L.orc_generate_rays.restype = None; L.orc_generate_rays.argtypes = [C.POINTER(OrcParams), C.c_uint32, vp, vp]
        L.orc_intersect.restype = None
        L.orc_intersect.argtypes = [vp, vp, C.c_uint32, C.c_int, C.c_int, C.c_int, vp]
        for name in ("orc_intersect_cb", "orc_intersect_cb_brute"):
            getattr(L, name).restype = None
        L.orc_shade.restype = None
        L.orc_shade.argtypes = [vp, C.POINTER(OrcParams), vp, vp, vp, C.c_uint32, vp, vp, vp, vp, vp]
        L.orc_soup.restype = None; L.orc_soup.argtypes = [C.c_uint32, C.c_uint32, C.c_float, vp, vp, vp]
        L.orc_morton30.restype = C.c_uint32; L.orc_morton30.argtypes = [C.c_float] * 3
        L.orc_hardware_threads.restype = C.c_uint
        _lib = L
    return _lib


HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Scene:
    """Oracle scene: world-space triangle list (+ median-split BVH above `brute_threshold` triangles)."""

    def __init__(self, verts, indices, faces, xforms=None, brute_threshold=64):
        self.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        self.indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        self.faces = np.ascontiguousarray(faces, np.float32).reshape(-1, 6)
        self.xforms = None if xforms is None else np.ascontiguousarray(xforms, np.float32).reshape(-1, 12)
        ninst = 0 if self.xforms is None else len(self.xforms)
        self.h = lib().orc_scene_create(_ptr(self.verts), len(self.verts), _ptr(self.indices), len(self.indices),
                                        _ptr(self.faces), len(self.faces), _ptr(self.xforms), ninst, brute_threshold)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h)
            self.h = None

    @property
    def ntris(self):
        return lib().orc_scene_ntris(self.h)

    def render(self, p, precision=32, brute=False, nthreads=0, image=None):
        """Runs one frame of the tile in p; returns (image HxWx4 float32, rays_traced)."""
        if image is None:
            image = np.zeros((p.height, p.width, 4), np.float32)
        rays = lib().orc_render(self.h, C.byref(p), precision, int(brute), nthreads, _ptr(image))
        return image, int(rays)

    def intersect(self, rays, precision=32, brute=False, nthreads=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros(len(rays), HIT_DTYPE)
        lib().orc_intersect(self.h, _ptr(rays), len(rays), precision, int(brute), nthreads, _ptr(hits))
        return hits

    def shade(self, p, hits, weight, seed):
        n = len(hits)
        hits = np.ascontiguousarray(hits)
        weight = np.ascontiguousarray(weight, np.float32).reshape(n, 3)
        seed = np.ascontiguousarray(seed, np.uint32)
        out = dict(contrib=np.zeros((n, 3), np.float32), ray=np.zeros((n, 8), np.float32),
                   weight=np.zeros((n, 3), np.float32), seed=np.zeros(n, np.uint32), alive=np.zeros(n, np.uint8))
        lib().orc_shade(self.h, C.byref(p), _ptr(hits), _ptr(weight), _ptr(seed), n, _ptr(out["contrib"]),
                        _ptr(out["ray"]), _ptr(out["weight"]), _ptr(out["seed"]), _ptr(out["alive"]))
        return out


def generate_rays(p, sample_in_frame=0):
    rows = p.height // p.tile_nranks if p.tile_block else (p.tile_rows if p.tile_rows else p.height - p.tile_y0)
    n = rows * p.width
    rays = np.zeros((n, 8), np.float32)
    seeds = np.zeros(n, np.uint32)
    lib().orc_generate_rays(C.byref(p), sample_in_frame, _ptr(rays), _ptr(seeds))
    return rays, seeds


def soup_scale(ntris):
    return float(np.float32(float(ntris) ** (-1.0 / 3.0)))


def soup(ntris, seed):
    verts = np.zeros((3 * ntris, 3), np.float32)
    idx = np.zeros(3 * ntris, np.uint32)
    faces = np.zeros((ntris, 6), np.float32)
    lib().orc_soup(ntris, seed, soup_scale(ntris), _ptr(verts), _ptr(idx), _ptr(faces))
    return verts, idx, faces


def rel_l2(a, b):
    """The parity metric: ||a-b||2 / ||b||2 over RGB, unclamped."""
    a = np.asarray(a, np.float64)[..., :3]
    b = np.asarray(b, np.float64)[..., :3]
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def load_cornell_golden():
    """Scene fixture produced by the reference's tinyobjloader (tests/golden/make_golden.py)."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "cornell_scene.json")) as f:
        g = json.load(f)
    verts = np.array([float.fromhex(h) for h in g["verts_hex"]], np.float32).reshape(-1, 3)
    faces = np.array([float.fromhex(h) for h in g["faces_hex"]], np.float32).reshape(-1, 6)
    idx = np.array(g["indices"], np.uint32)
    return verts, idx, faces, g


def load_cornell_raw_golden():
    """The raw tinyobj arrays of the same asset (fixture fields raw_*): positions, corner vertex indices, face material
    ids, materials {Kd, Ke} — the input of the reference's loadFromFile body (main.cpp:37-57)."""
    _, _, _, g = load_cornell_golden()
    pos = np.array([float.fromhex(h) for h in g["raw_positions_hex"]], np.float32).reshape(-1, 3)
    mats = np.array([float.fromhex(h) for h in g["raw_materials_hex"]], np.float32).reshape(-1, 6)
    return pos, np.array(g["raw_corner_vertex"], np.int32), np.array(g["raw_face_material"], np.int32), mats


def ref_load_obj(obj_path, mtl_dir):
    """Runs the reference's vendored tinyobjloader (oracle/_ref). Raises if the library is absent."""
    L = C.CDLL(REF_LOADER_PATH)
    L.ref_load_obj.restype = C.c_int
    nv, ni, nf, ns = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    err = C.create_string_buffer(1024)
    shape_tris = (C.c_uint32 * 4096)()
    rc = L.ref_load_obj(obj_path.encode(), mtl_dir.encode(), None, None, None, C.byref(nv), C.byref(ni), C.byref(nf),
                        shape_tris, C.byref(ns), err, 1024)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    verts = np.zeros((nv.value, 3), np.float32)
    idx = np.zeros(ni.value, np.uint32)
    faces = np.zeros((nf.value, 6), np.float32)
    rc = L.ref_load_obj(obj_path.encode(), mtl_dir.encode(), _ptr(verts), _ptr(idx), _ptr(faces), C.byref(nv),
                        C.byref(ni), C.byref(nf), shape_tris, C.byref(ns), err, 1024)
    assert rc == 0
    return verts, idx, faces, list(shape_tris[:ns.value])


# ------------------------------------------------------------------------------------------ the reference's shader text
_ref_shade = None


def ref_shade_available():
    return os.path.exists(REF_SHADE_PATH)


def ref_shade_lib():
    """oracle/_ref/libref_shade.so: the reference's shaders/*.glsl|rgen|rchit|rmiss compiled as C++ (oracle/glsl_shim.h)."""
    global _ref_shade
    if _ref_shade is None:
        build()
        L = C.CDLL(REF_SHADE_PATH)
        vp = C.c_void_p
        L.ref_shade_render.restype = C.c_uint64
        L.ref_shade_render.argtypes = ([vp, vp, C.c_uint32, vp] + [C.c_uint32] * 5 + [C.c_int] * 4 + [vp, vp, C.c_int, vp])
        L.ref_pcg.restype = C.c_uint32; L.ref_pcg.argtypes = [C.POINTER(C.c_uint32)]
        L.ref_pcg2d.restype = None; L.ref_pcg2d.argtypes = [C.POINTER(C.c_uint32)] * 2
        L.ref_rand.restype = C.c_float; L.ref_rand.argtypes = [C.POINTER(C.c_uint32)]
        L.ref_sample_direction.restype = None; L.ref_sample_direction.argtypes = [C.c_float, C.c_float, vp, vp]
        _ref_shade = L
    return _ref_shade


def ref_shade_render(verts, indices, faces, width, height, frames=1, spp=0, depth=0, rgba8=False, rows=(0, 0),
                     scene=None, brute=True, nthreads=0, image=None, first_frame=0, row_step=1):
    """`frames` launches traceRaysKHR(width, height, 1) of the reference's shader text with push constant frame =
    first_frame.. over one storage image. spp / depth = 0 keep the literals of the text (32 / 8). scene: an oracle
    Scene whose intersector answers traceRayEXT (brute force or its BVH); None = the library's own brute force."""
    L = ref_shade_lib()
    verts = np.ascontiguousarray(verts, np.float32)
    indices = np.ascontiguousarray(indices, np.uint32)
    faces = np.ascontiguousarray(faces, np.float32)
    if image is None:
        image = np.zeros((height, width, 4), np.float32)
    fn, user = None, None
    if scene is not None:
        fn = C.cast(lib().orc_intersect_cb_brute if brute else lib().orc_intersect_cb, C.c_void_p)
        user = scene.h
    rays = 0
    for f in range(first_frame, first_frame + frames):
        rays += L.ref_shade_render(_ptr(verts), _ptr(indices), len(indices), _ptr(faces), width, height, rows[0], rows[1], row_step,
                                   f, spp, depth, int(rgba8), fn, user, nthreads, _ptr(image))
    return image, int(rays)


def ref_load_obj_raw(obj_path, mtl_dir):
    """The raw arrays tinyobj::LoadObj gives the reference's loader (oracle/_ref): positions (n,3), the vertex index of
    every face corner, the material id of every face, {Kd, Ke} per material."""
    L = C.CDLL(REF_LOADER_PATH)
    L.ref_load_obj_raw.restype = C.c_int
    n = [C.c_uint32() for _ in range(4)]
    args = [obj_path.encode(), mtl_dir.encode()]
    assert L.ref_load_obj_raw(*args, None, None, None, None, *[C.byref(x) for x in n]) == 0
    pos = np.zeros((n[0].value, 3), np.float32)
    corner = np.zeros(n[1].value, np.int32)
    fmat = np.zeros(n[2].value, np.int32)
    mats = np.zeros((n[3].value, 6), np.float32)
    assert L.ref_load_obj_raw(*args, _ptr(pos), _ptr(corner), _ptr(fmat), _ptr(mats), *[C.byref(x) for x in n]) == 0
    return pos, corner, fmat, mats
