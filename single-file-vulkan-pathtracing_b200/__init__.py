"""single-file-vulkan-pathtracing_b200 — Python mirror of the C ABI in include/bpt.h.

The product is libbpt.so (hand-written sm_100a CUDA behind an extern "C" boundary); this module is a thin
ctypes binding used by the tests, bench.py and Python callers. There is NO CPU fallback: importing works
anywhere (so the symbol table can be checked on a CPU box), but creating a `PathTracer` without the built
library or without a B200 raises.

Reference call sites replaced (paths relative to the reference checkout):
    PathTracer.upload_mesh   -> Buffer(...) x3, main.cpp:492-494
    PathTracer.build_accel   -> Accel(...) -> buildAccelerationStructuresKHR, main.cpp:416-450,512,538
    PathTracer.trace         -> pushConstants(frame) + traceRaysKHR(W,H,1), main.cpp:658-659
    PathTracer.read_image    -> the storage image bound at binding 1, main.cpp:481-484
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BPT_LIB_VARIANT=name loads lib/libbpt_name.so (a `make VARIANT=name` build of csrc/: A/B experiments on the GPU box)
_VARIANT = os.environ.get("BPT_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, "lib", "libbpt%s.so" % ("_" + _VARIANT if _VARIANT else ""))
MISS = 0xFFFFFFFF
ACCUM_FLOAT4, ACCUM_RGBA8 = 0, 1
SAMPLER_UNIFORM, SAMPLER_COSINE = 0, 1
OPT_PROFILE, OPT_COUNT_TRAVERSAL, OPT_SMEM_TOP_NODES, OPT_STREAMS = 1, 2, 3, 4
OPT_TRACE_REFILL_BELOW, OPT_TRACE_STEPS_PER_REFILL, OPT_PASS_PATHS, OPT_TRACE_STAGED_TRIS_PER_STEP = 8, 9, 10, 11
OPT_USE_GRAPH = 6
OPT_BVH_SAH_SUBTREE = 12
OPT_FUSED_PATHS = 13
OPT_TRACE_BLOCK = 5
OPT_BVH_OPTIMAL_COLLAPSE = 7
NCCL_UNIQUE_ID_BYTES = 128


class BptError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bpt error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """bpt_params (include/bpt.h)."""
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("spp_per_frame", C.c_uint32), ("max_depth", C.c_uint32),
        ("frame", C.c_int32), ("tile_y0", C.c_uint32), ("tile_rows", C.c_uint32),
        ("cam_origin", C.c_float * 3), ("cam_target", C.c_float * 3), ("sky", C.c_float * 3),
        ("tmin", C.c_float), ("tmax", C.c_float), ("accum_mode", C.c_uint32), ("sampler", C.c_uint32),
        ("tile_block", C.c_uint32), ("tile_nranks", C.c_uint32), ("tile_rank", C.c_uint32),
        ("rr_start_depth", C.c_uint32), ("nee", C.c_uint32),
    ]


class Stats(C.Structure):
    """bpt_stats."""
    _fields_ = [
        ("rays_traced", C.c_uint64), ("paths", C.c_uint64), ("trace_launches", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("trace_kernel_ms", C.c_double), ("frame_ms", C.c_double),
        ("build_ms", C.c_double), ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64),
        ("warp_iterations", C.c_uint64), ("warp_node_steps", C.c_uint64), ("warp_tri_steps", C.c_uint64),
        ("lane_iterations", C.c_uint64),
    ]


class AccelInfo(C.Structure):
    """bpt_accel_info."""
    _fields_ = [
        ("num_tris", C.c_uint32), ("num_instances", C.c_uint32), ("num_nodes8", C.c_uint32),
        ("num_binary_nodes", C.c_uint32), ("top_nodes_smem", C.c_uint32), ("max_depth8", C.c_uint32),
        ("bytes_nodes", C.c_uint64), ("bytes_tris", C.c_uint64), ("num_tlas_nodes8", C.c_uint32),
        ("num_records", C.c_uint32),
    ]


HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])
# the two faces of a 64-byte record (csrc/common.cuh)
NODE8_DTYPE = np.dtype([
    ("org", "<u8"), ("e_valid", "<u4"), ("child_base", "<u4"),
    ("qlox", "u1", 8), ("qloy", "u1", 8), ("qloz", "u1", 8), ("qhix", "u1", 8), ("qhiy", "u1", 8), ("qhiz", "u1", 8)])
WOOP_DTYPE = np.dtype([("rows", "<f4", (3, 4)), ("prim", "<u4"), ("pad", "<u4", 3)])
assert NODE8_DTYPE.itemsize == 64 and WOOP_DTYPE.itemsize == 64

# every symbol include/bpt.h declares: (restype, argtypes)
_vp, _u32, _i32, _sz = C.c_void_p, C.c_uint32, C.c_int, C.c_size_t
ABI = {
    "bpt_abi_version": (_i32, []),
    "bpt_params_default": (None, [C.POINTER(Params)]),
    "bpt_create": (_i32, [_i32, _vp, C.POINTER(_vp)]),
    "bpt_destroy": (None, [_vp]),
    "bpt_last_error": (C.c_char_p, [_vp]),
    "bpt_set_option": (_i32, [_vp, _i32, C.c_int64]),
    "bpt_upload_mesh": (_i32, [_vp, _vp, _u32, _vp, _u32, _vp, _u32]),
    "bpt_upload_mesh_device": (_i32, [_vp, _vp, _u32, _vp, _u32, _vp, _u32]),
    "bpt_upload_obj_arrays": (_i32, [_vp, _vp, _u32, _vp, _u32, _vp, _vp, _u32]),
    "bpt_set_instances": (_i32, [_vp, _vp, _u32]),
    "bpt_upload_soup": (_i32, [_vp, _u32, _u32]),
    "bpt_build_accel": (_i32, [_vp]),
    "bpt_accel_info_get": (_i32, [_vp, C.POINTER(AccelInfo)]),
    "bpt_trace": (_i32, [_vp, C.POINTER(Params)]),
    "bpt_sync": (_i32, [_vp]),
    "bpt_read_image": (_i32, [_vp, _vp, _sz]),
    "bpt_read_image_bgra8": (_i32, [_vp, _vp, _sz]),
    "bpt_read_image_async": (_i32, [_vp, _vp, _sz]),
    "bpt_read_image_bgra8_async": (_i32, [_vp, _vp, _sz]),
    "bpt_read_wait": (_i32, [_vp]),
    "bpt_image_device_ptr": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_sz)]),
    "bpt_clear_image": (_i32, [_vp]),
    "bpt_get_stats": (_i32, [_vp, C.POINTER(Stats)]),
    "bpt_reset_stats": (_i32, [_vp]),
    "bpt_trace_rays": (_i32, [_vp, _vp, _u32, _vp]),
    "bpt_shade_step": (_i32, [_vp, C.POINTER(Params), _u32] + [_vp] * 9),
    "bpt_generate_rays": (_i32, [_vp, C.POINTER(Params), _u32, _vp, _vp]),
    "bpt_download_accel": (_i32, [_vp, _vp, _vp, _vp]),
    "bpt_download_mesh": (_i32, [_vp, _vp, _vp, _vp]),
    "bpt_download_morton": (_i32, [_vp, _vp, _u32]),
    "bpt_download_leaf_order": (_i32, [_vp, _vp, _u32]),
    "bpt_download_lbvh": (_i32, [_vp, _vp, _vp, _vp]),
    "bpt_nccl_unique_id": (_i32, [_vp]),
    "bpt_nccl_init": (_i32, [_vp, _vp, _i32, _i32]),
    "bpt_allgather_image": (_i32, [_vp, _u32, _u32]),
    "bpt_tile_rows": (None, [_u32, _i32, _i32, C.POINTER(_u32), C.POINTER(_u32)]),
}

_lib = None


def load_library():
    """Loads libbpt.so and binds every ABI symbol. Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)  # AttributeError if a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def default_params(width=1024, height=1024, spp=32, depth=8, frame=0, **kw):
    """bpt_params_default() (the reference's constants) with overrides."""
    p = Params()
    load_library().bpt_params_default(C.byref(p))
    p.width, p.height, p.spp_per_frame, p.max_depth, p.frame = width, height, spp, depth, frame
    for k, v in kw.items():
        if k in ("cam_origin", "cam_target", "sky"):
            getattr(p, k)[:] = v
        else:
            setattr(p, k, v)
    return p


def tile_rows(height, rank, nranks):
    y0, rows = C.c_uint32(), C.c_uint32()
    load_library().bpt_tile_rows(height, rank, nranks, C.byref(y0), C.byref(rows))
    return y0.value, rows.value


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PathTracer:
    """One context = one GPU + one CUDA stream (bpt_context)."""

    def __init__(self, device=0, stream=None):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.bpt_create(device, stream, C.byref(h))
        if rc != 0:
            raise BptError(rc, self._L.bpt_last_error(None).decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.bpt_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise BptError(rc, self._L.bpt_last_error(self._h).decode())

    # -- scene -------------------------------------------------------------------------------
    def upload_mesh(self, verts, indices, faces):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
        indices = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        faces = np.ascontiguousarray(faces, np.float32).reshape(-1, 6)
        self._check(self._L.bpt_upload_mesh(self._h, _ptr(verts), len(verts), _ptr(indices), len(indices),
                                            _ptr(faces), len(faces)))

    def upload_obj_arrays(self, positions, corner_vertex, face_material, materials_kd_ke):
        """The reference's loadFromFile body on the device (bpt_upload_obj_arrays) from the arrays tinyobj::LoadObj returns."""
        positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        corner_vertex = np.ascontiguousarray(corner_vertex, np.int32).reshape(-1)
        face_material = np.ascontiguousarray(face_material, np.int32).reshape(-1)
        materials_kd_ke = np.ascontiguousarray(materials_kd_ke, np.float32).reshape(-1, 6)
        if len(face_material) * 3 != len(corner_vertex):
            raise ValueError("one material id per face (3 corners)")
        self._check(self._L.bpt_upload_obj_arrays(self._h, _ptr(positions), len(positions), _ptr(corner_vertex), len(corner_vertex),
                                                  _ptr(face_material), _ptr(materials_kd_ke), len(materials_kd_ke)))

    def upload_soup(self, ntris, seed):
        self._check(self._L.bpt_upload_soup(self._h, ntris, seed))

    def set_instances(self, xforms):
        xforms = np.ascontiguousarray(xforms, np.float32).reshape(-1, 12)
        self._check(self._L.bpt_set_instances(self._h, _ptr(xforms), len(xforms)))

    def set_option(self, option, value):
        self._check(self._L.bpt_set_option(self._h, option, int(value)))

    def build_accel(self):
        self._check(self._L.bpt_build_accel(self._h))
        return self.accel_info()

    def accel_info(self):
        info = AccelInfo()
        self._check(self._L.bpt_accel_info_get(self._h, C.byref(info)))
        return info

    # -- render ------------------------------------------------------------------------------
    def trace(self, params):
        self._check(self._L.bpt_trace(self._h, C.byref(params)))

    def sync(self):
        self._check(self._L.bpt_sync(self._h))

    def read_image(self, width, height, out=None):
        if out is None:
            out = np.empty((height, width, 4), np.float32)
        self._check(self._L.bpt_read_image(self._h, _ptr(out), out.size))
        return out

    def read_image_bgra8(self, width, height):
        out = np.empty((height, width, 4), np.uint8)
        self._check(self._L.bpt_read_image_bgra8(self._h, _ptr(out), out.size))
        return out

    def read_image_async(self, out):
        """Enqueues the read-back of the image on the context's copy stream into `out` (H x W x 4 float32, ideally pinned);
        the next trace may be enqueued at once. `out` is valid after read_wait()."""
        self._check(self._L.bpt_read_image_async(self._h, _ptr(out), out.size))

    def read_image_bgra8_async(self, out):
        self._check(self._L.bpt_read_image_bgra8_async(self._h, _ptr(out), out.size))

    def read_wait(self):
        self._check(self._L.bpt_read_wait(self._h))

    def image_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._L.bpt_image_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def clear_image(self):
        self._check(self._L.bpt_clear_image(self._h))

    def render(self, params, frames=1):
        """Convenience: `frames` successive frames starting at params.frame, then the image."""
        p = Params.from_buffer_copy(params)
        for f in range(frames):
            p.frame = params.frame + f
            self.trace(p)
        return self.read_image(p.width, p.height)

    def stats(self):
        s = Stats()
        self._check(self._L.bpt_get_stats(self._h, C.byref(s)))
        return s

    def reset_stats(self):
        self._check(self._L.bpt_reset_stats(self._h))

    # -- stage level -------------------------------------------------------------------------
    def trace_rays(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros(len(rays), HIT_DTYPE)
        self._check(self._L.bpt_trace_rays(self._h, _ptr(rays), len(rays), _ptr(hits)))
        return hits

    def shade_step(self, params, rays, hits, weight, seed):
        """One closest-hit / miss + path-update step (bpt_shade_step); returns a dict of contrib, ray, weight, seed and alive,
        one row per path."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = len(rays)
        hits = np.ascontiguousarray(hits, HIT_DTYPE)
        weight = np.ascontiguousarray(weight, np.float32).reshape(n, 3)
        seed = np.ascontiguousarray(seed, np.uint32)
        out = dict(contrib=np.zeros((n, 3), np.float32), ray=np.zeros((n, 8), np.float32),
                   weight=np.zeros((n, 3), np.float32), seed=np.zeros(n, np.uint32), alive=np.zeros(n, np.uint8))
        self._check(self._L.bpt_shade_step(self._h, C.byref(params), n, _ptr(rays), _ptr(hits), _ptr(weight), _ptr(seed),
                                           _ptr(out["contrib"]), _ptr(out["ray"]), _ptr(out["weight"]), _ptr(out["seed"]),
                                           _ptr(out["alive"])))
        return out

    def generate_rays(self, params, sample_in_frame=0):
        rows = (params.height // params.tile_nranks if params.tile_block
                else (params.tile_rows if params.tile_rows else params.height - params.tile_y0))
        n = rows * params.width
        rays = np.zeros((n, 8), np.float32)
        seeds = np.zeros(n, np.uint32)
        self._check(self._L.bpt_generate_rays(self._h, C.byref(params), sample_in_frame, _ptr(rays), _ptr(seeds)))
        return rays, seeds

    # -- introspection -----------------------------------------------------------------------
    def download_accel(self):
        """(records as NODE8_DTYPE, the same bytes as WOOP_DTYPE, rec_prim, grid_bias, grid_step) of the mesh-level BVH;
        a node origin is fmaf(float(2^23 + c), grid_step, grid_bias) per axis."""
        info = self.accel_info()
        recs = np.zeros(info.num_records, NODE8_DTYPE)
        rec_prim = np.zeros(info.num_records, np.uint32)
        grid = np.zeros(6, np.float32)
        self._check(self._L.bpt_download_accel(self._h, _ptr(recs), _ptr(rec_prim), _ptr(grid)))
        return recs, recs.view(WOOP_DTYPE), rec_prim, grid[:3].copy(), grid[3:].copy()

    def download_mesh(self, ntris, nverts=None):
        nverts = 3 * ntris if nverts is None else nverts
        verts = np.zeros((nverts, 3), np.float32)
        idx = np.zeros(3 * ntris, np.uint32)
        faces = np.zeros((ntris, 6), np.float32)
        self._check(self._L.bpt_download_mesh(self._h, _ptr(verts), _ptr(idx), _ptr(faces)))
        return verts, idx, faces

    def download_morton(self):
        n = self.accel_info().num_tris
        keys = np.zeros(n, np.uint64)
        self._check(self._L.bpt_download_morton(self._h, _ptr(keys), n))
        return keys

    def download_leaf_order(self):
        n = self.accel_info().num_tris
        prims = np.zeros(n, np.uint32)
        self._check(self._L.bpt_download_leaf_order(self._h, _ptr(prims), n))
        return prims

    def download_lbvh(self):
        n = self.accel_info().num_tris
        left = np.zeros(max(n - 1, 0), np.uint32)
        right = np.zeros(max(n - 1, 0), np.uint32)
        aabbs = np.zeros((2 * n - 1, 6), np.float32)
        self._check(self._L.bpt_download_lbvh(self._h, _ptr(left), _ptr(right), _ptr(aabbs)))
        return left, right, aabbs

    # -- multi-GPU ---------------------------------------------------------------------------
    @staticmethod
    def nccl_unique_id():
        L = load_library()
        buf = (C.c_uint8 * NCCL_UNIQUE_ID_BYTES)()
        rc = L.bpt_nccl_unique_id(buf)
        if rc != 0:
            raise BptError(rc, L.bpt_last_error(None).decode())
        return bytes(buf)

    def nccl_init(self, unique_id, rank, nranks):
        buf = (C.c_uint8 * NCCL_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self._L.bpt_nccl_init(self._h, buf, rank, nranks))

    def allgather_image(self, width, height):
        self._check(self._L.bpt_allgather_image(self._h, width, height))
