#!/usr/bin/env python
"""bench.py — throughput of the hot path (build once, then trace W x H x spp paths per step) on 1..8 B200s.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path through the C ABI (libbpt.so)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference (oracle/)

Workload (BASELINE.json configs[3], the one `north_star` quotes its targets on): the synthetic 10 M-triangle soup
(seed 0x5EED0002), 4096 x 4096, depth 8. One *step* is one frame as the reference defines it — one trace call of
32 samples per pixel (maxSamples, raygen.rgen:43) over the whole image; 4 steps are the configuration's 128 spp (the
per-sample seeds are the reference's global sample index, so steps simply continue the same render). With N GPUs the image's row blocks are dealt round-robin to the ranks
(fixed image => "strong" scaling), every rank builds the same BVH, and one NCCL all-gather assembles the image
after the last step, inside the timed region.

One JSON line on stdout (rank 0). `value` is whole-job Mray/s with everything resident in HBM; `e2e` is the same
metric through the public API with host buffers (mesh upload from pinned host memory + build + per-step image
read-back inside the timed region, the read-back of frame f overlapping the trace of frame f+1 as a present loop
would run it); `roofline` describes the traversal kernel: what binds it (instruction issue, by ncu), its DRAM floor
against the measured HBM peak, and — only when profiles/k_trace_traffic.json holds an ncu capture of THIS csrc/ and
workload — the measured DRAM traffic; `cpu_baseline` is the oracle on the host cores on a bounded sample of the same
workload. Only the cpu_baseline / --impl reference legs touch oracle/. With N > 1 the gathered image is verified
against a single-tile re-render after the timed regions (`allgather_verified`).
"""
import argparse
import hashlib
import importlib
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: BASELINE.json config it is
    "soup10m": dict(tris=10_000_000, seed=0x5EED0002, width=4096, height=4096, spp=32, depth=8, full_spp=128,
                    config="configs[3]: synthetic 10 M triangle soup, 4096x4096, 128 spp (4 steps of the reference's 32 spp), depth 8"),
    "soup1m": dict(tris=1_000_000, seed=0x5EED0001, width=1920, height=1080, spp=32, depth=8, full_spp=64,
                   config="configs[2]: synthetic 1 M triangle soup, 1920x1080, 64 spp (2 steps of the reference's 32 spp), depth 8"),
    "cornell": dict(tris=0, seed=0, width=1024, height=1024, spp=32, depth=8, full_spp=256,
                    config="configs[1]: CornellBox-Original.obj, 1024x1024, 256 spp (8 steps of 32 spp), depth 8"),
    "cornell1000": dict(tris=0, seed=0, width=2048, height=2048, spp=32, depth=8, full_spp=512, instances=10,
                        cam_origin=(0.0, -1.0, 55.0), cam_target=(0.0, -1.0, 52.0),
                        config="configs[4]: Cornell box x1000 instances (10x10x10 grid, pitch 2.5, two-level BVH), 2048x2048, "
                               "512 spp (16 steps of 32 spp), depth 8, camera pulled back to z=55"),
}
METRIC = "Mray/s"
TILE_BLOCK = 8          # rows per interleaved block
CPU_SAMPLE_ROWS = 128   # rows of the image the CPU baseline renders per step (spread uniformly over the image)
CPU_SAMPLE_SPP = 12     # samples per pixel of those rows: 10-20 s of work on 16-32 cores for the 10 M soup


_RESULT_LINE = []  # rank 0's JSON line, printed by main() once stdout is restored


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="soup10m", choices=sorted(WORKLOADS))
    ap.add_argument("--tris", type=int, default=None, help="override the soup size (debugging)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--spp", type=int, default=None, help="samples per pixel per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--backend", default="nccl", help="torch.distributed backend (tests use gloo)")
    ap.add_argument("--opt", action="append", default=[], help="developer knob: OPTION_ID=VALUE for bpt_set_option (recorded in config)")
    return ap.parse_args(argv)


def workload_of(args):
    w = dict(WORKLOADS[args.workload])
    w["name"] = args.workload
    for k in ("tris", "width", "height", "spp"):
        if getattr(args, k) is not None:
            w[k] = getattr(args, k)
            w["name"] = args.workload + "-custom"
    return w


# ------------------------------------------------------------------------------------------ multi-process plumbing
class Dist:
    """torch.distributed for the plumbing only: barrier, max/sum over ranks, byte broadcast."""

    def __init__(self, n_gpus, backend="nccl"):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend
        self.active = self.world > 1
        if self.world != max(1, n_gpus) and self.active:
            raise SystemExit(f"--gpus {n_gpus} but WORLD_SIZE={self.world}")
        if self.active:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                kw["device_id"] = torch.device("cuda", self.local_rank)
            dist.init_process_group(backend, rank=self.rank, world_size=self.world, **kw)
            self.dist = dist

    def _tensor(self, values, dtype):
        import torch
        dev = torch.device("cuda", self.local_rank) if self.backend == "nccl" else torch.device("cpu")
        return torch.tensor(values, dtype=dtype, device=dev)

    def barrier(self):
        if self.active:
            self.dist.barrier()

    def max(self, x):
        if not self.active:
            return float(x)
        import torch
        t = self._tensor([float(x)], torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if not self.active:
            return float(x)
        import torch
        t = self._tensor([float(x)], torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_floats(self, x):
        """[x of rank 0, x of rank 1, ...] on every rank."""
        if not self.active:
            return [float(x)]
        import torch
        t = self._tensor([float(x)], torch.float64)
        out = [torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def broadcast_bytes(self, data, nbytes):
        """rank 0's `data` (bytes of length nbytes) to everyone."""
        if not self.active:
            return data
        import torch
        t = self._tensor(list(data) if self.rank == 0 else [0] * nbytes, torch.uint8)
        self.dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    def close(self):
        if self.active:
            self.dist.destroy_process_group()


def interleaved_rows(height, block, nranks, rank):
    """Image rows of rank `rank` under the round-robin row-block tiling (bpt_params.tile_block), in tile order."""
    rows = []
    for b in range(rank, height // block, nranks):
        rows.extend(range(b * block, (b + 1) * block))
    return rows


def tile_kwargs(height, nranks, rank):
    """bpt_params tiling fields for this rank (none for a single GPU)."""
    if nranks == 1:
        return {}
    block = TILE_BLOCK
    while height % (block * nranks):
        block //= 2
        if block == 0:
            raise SystemExit(f"height {height} cannot be dealt in row blocks to {nranks} ranks")
    return dict(tile_block=block, tile_nranks=nranks, tile_rank=rank)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons of the job's GPUs, sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, indices):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(i) for i in indices), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    @classmethod
    def summarise(cls, lines, t0, t1):
        """(whole-job summary for the `clocks` key, per-GPU median SM clock and peak power)"""
        per = {}
        sm, smax, power, reasons = [], [], [], set()
        for (t, line) in lines:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                gpu, c, cmax, w = int(f[0]), float(f[1]), float(f[2]), float(f[3])
            except ValueError:
                continue
            sm.append(c); smax.append(cmax); power.append(w)
            g = per.setdefault(gpu, {"sm": [], "w": []})
            g["sm"].append(c); g["w"].append(w)
            for name, v in zip(cls.NAMES, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        clocks = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                  "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}
        per_gpu = {str(g): {"sm_mhz": float(np.median(v["sm"])), "sm_mhz_min": min(v["sm"]), "power_w_max": max(v["w"])}
                   for g, v in sorted(per.items())}
        return clocks, per_gpu

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}, {}
        time.sleep(0.25)
        self.proc.terminate()
        return self.summarise(self.lines, t0, t1)


# ------------------------------------------------------------------------------------------ CPU legs (oracle)
def instance_grid(n, pitch=2.5):
    """n^3 translations (3x4 row-major) on a grid centred on the box centre (0,-1,0): BASELINE config 5."""
    c = 0.5 * (n - 1) * pitch
    xf = np.zeros((n * n * n, 3, 4), np.float32)
    xf[:, 0, 0] = xf[:, 1, 1] = xf[:, 2, 2] = 1.0
    g = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 3)
    xf[:, :, 3] = g * pitch - c
    return xf.reshape(-1, 12)


def camera_kwargs(w):
    return {k: w[k] for k in ("cam_origin", "cam_target") if k in w}


def oracle_scene(w):
    """The workload's scene in the oracle (CPU restatement of the reference). Checker / baseline only."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    if w["tris"]:
        verts, idx, faces = O.soup(w["tris"], w["seed"])
        return O, O.Scene(verts, idx, faces)
    verts, idx, faces, _ = O.load_cornell_golden()
    return O, O.Scene(verts, idx, faces, xforms=instance_grid(w["instances"]) if w.get("instances") else None)


def cpu_sample_params(O, w, frame, rows_target=None):
    rows = min(rows_target or CPU_SAMPLE_ROWS, w["height"])
    while w["height"] % rows:
        rows -= 1
    # `rows` single rows spread uniformly over the image: the interleaved tiling with 1-row blocks, rank 0
    return O.default_params(w["width"], w["height"], CPU_SAMPLE_SPP, w["depth"], frame, tile_block=1,
                            tile_nranks=w["height"] // rows, tile_rank=0, **camera_kwargs(w)), rows


def cpu_baseline(w, steps=1, warmup=0, rows_target=None):
    """The reference's algorithm on all host cores over a bounded sample: CPU_SAMPLE_ROWS rows x full width x
    CPU_SAMPLE_SPP spp per step. kind "reference": the reference's OWN shader text compiled as C++
    (oracle/_ref/libref_shade.so, built from /root/reference/shaders where they lie) does everything the reference's
    source states; traceRayEXT — the driver's closed-source traversal — is answered by the oracle's BVH. Workloads the
    shader text cannot express (another camera, instance transforms: SURVEY T11/T12) and boxes without that library
    run the oracle's restatement instead (kind "port")."""
    O, scene = oracle_scene(w)
    cores = int(O.lib().orc_hardware_threads())
    img = np.zeros((w["height"], w["width"], 4), np.float32)
    use_text = O.ref_shade_available() and not w.get("instances") and "cam_origin" not in w
    rays_total, secs = 0, 0.0
    for s in range(warmup + steps):
        p, rows = cpu_sample_params(O, w, s, rows_target)
        t0 = time.perf_counter()
        if use_text:
            _, rays = O.ref_shade_render(scene.verts, scene.indices, scene.faces, w["width"], w["height"], 1, CPU_SAMPLE_SPP,
                                         w["depth"], False, (0, 0), scene, False, cores, img, s, w["height"] // rows)
        else:
            _, rays = scene.render(p, 32, nthreads=cores, image=img)
        dt = time.perf_counter() - t0
        if s >= warmup:
            rays_total += rays
            secs += dt
    return {"value": rays_total / secs / 1e6, "unit": METRIC, "cores": cores, "kind": "reference" if use_text else "port",
            "sample": f"{rows} rows spread uniformly over the {w['width']}x{w['height']} image x {CPU_SAMPLE_SPP} spp x depth "
                      f"{w['depth']} per step, {steps} step(s): {rays_total} rays in {secs:.2f} s ("
                      + ("the reference's shader text compiled as C++, traceRayEXT answered by the oracle's median-split BVH)"
                         if use_text else "oracle's restatement, own median-split BVH)")}, \
        rays_total, secs


def run_reference(args):
    """--impl reference: the reference has no CPU path and its Vulkan build cannot run in this image, so this arm
    times the reference's shader text compiled as C++ over the oracle's BVH (cpu_baseline above; the oracle's
    restatement where the text cannot express the workload) on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload_of(args)
    # the per-step sample shrinks with the number of steps so that the whole run stays within about a minute of CPU time
    n = args.steps + args.warmup
    rows_target = CPU_SAMPLE_ROWS if n <= 6 else max(64, CPU_SAMPLE_ROWS * 6 // n)   # >= 4 rows per thread on a 16-core host
    base, rays, secs = cpu_baseline(w, steps=args.steps, warmup=args.warmup, rows_target=rows_target)
    val = base["value"]
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": w["name"], "baseline_config": w["config"], "tris": w["tris"], "width": w["width"],
                      "height": w["height"], "depth": w["depth"]},
           "cpu_baseline": base,
           "e2e": {"value": val, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _RESULT_LINE.append(json.dumps(out))


# ------------------------------------------------------------------------------------------ ncu evidence
def csrc_hash():
    """sha256 over the CUDA sources the library is built from: an ncu capture describes THIS code or it is not used."""
    d = os.path.join(ROOT, "single-file-vulkan-pathtracing_b200", "csrc")
    h = hashlib.sha256()
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".cpp", ".h")):
            h.update(name.encode())
            with open(os.path.join(d, name), "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


def ncu_capture(workload, n_gpus):
    """The k_trace entry of profiles/k_trace_traffic.json (written by profiles/summarize.py traffic from an ncu capture
    of `bench.py --workload W`, all bounces of a frame) for this workload, GPU count and csrc hash, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "k_trace_traffic.json")) as f:
            entries = json.load(f)
    except (OSError, ValueError):
        return None
    for e in entries:
        if e.get("workload") == workload and int(e.get("n_gpus", 1)) == n_gpus and e.get("csrc_hash") == csrc_hash():
            return e
    return None


def verify_allgather(bpt, pt, d, w, params, tile, build_scene, stream):
    """N > 1: the image NCCL assembled must be the single-GPU image. Two frames are rendered tiled and all-gathered;
    every rank checksums the full image (all ranks must hold the same bytes), and rank 0 re-renders, on a second
    context as ordinary contiguous tiles, eight row blocks that belong to OTHER ranks and compares them bit for bit."""
    W, H = w["width"], w["height"]
    pt.clear_image()
    for f in range(2):
        pt.trace(params(f))
    pt.allgather_image(W, H)
    img = pt.read_image(W, H)
    crc = float(zlib.crc32(img.tobytes()))
    crcs = d.gather_floats(crc)
    ok = len(set(crcs)) == 1
    detail = {"crc_equal_on_all_ranks": ok, "rows_checked": 0}
    if d.rank == 0:
        block = tile["tile_block"]
        nblocks = H // block
        # row blocks of the other ranks, spread over the image
        cand = [b for b in range(nblocks) if b % d.world != 0]
        pick = [cand[i * len(cand) // 8] for i in range(8)] if len(cand) >= 8 else cand
        pt2 = bpt.PathTracer(d.local_rank, stream.cuda_stream)
        build_scene(pt2)
        same = True
        for b in pick:
            pt2.clear_image()
            for f in range(2):
                pt2.trace(bpt.default_params(W, H, w["spp"], w["depth"], f, tile_y0=b * block, tile_rows=block, **camera_kwargs(w)))
            part = pt2.read_image(W, H)[b * block:(b + 1) * block]
            same = same and bool(np.array_equal(part, img[b * block:(b + 1) * block]))
            detail["rows_checked"] += block
        pt2.close()
        detail["foreign_rows_bit_identical"] = same
        ok = ok and same
    ok = d.max(0.0 if ok else 1.0) == 0.0
    pt.clear_image()
    return ok, detail


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    bpt = importlib.import_module("single-file-vulkan-pathtracing_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libbpt has no CPU fallback")
    w = workload_of(args)
    d = Dist(args.gpus, args.backend)
    torch.cuda.set_device(d.local_rank)
    stream = torch.cuda.Stream(device=d.local_rank)
    pt = bpt.PathTracer(d.local_rank, stream.cuda_stream)
    for o in args.opt:
        k, v = o.split("=")
        pt.set_option(int(k), int(v))
    lanes = 1
    for o in args.opt:
        if o.split("=")[0] == str(bpt.OPT_STREAMS):
            lanes = int(o.split("=")[1])
    W, H, K, WU = w["width"], w["height"], args.steps, args.warmup
    tile = tile_kwargs(H, d.world, d.rank)

    def params(frame):
        return bpt.default_params(W, H, w["spp"], w["depth"], frame, **tile, **camera_kwargs(w))

    # ---- scene + build (once per job; reported, not part of `value`)
    def build_scene(t):
        if w["tris"]:
            t.upload_soup(w["tris"], w["seed"])
        else:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as O  # fixture loader only (tests/golden/cornell_scene.json)
            verts, idx, faces, _ = O.load_cornell_golden()
            t.upload_mesh(verts, idx, faces)
            if w.get("instances"):
                t.set_instances(instance_grid(w["instances"]))
        return t.build_accel()
    info = build_scene(pt)
    build_ms = pt.stats().build_ms
    if d.active:
        uid = d.broadcast_bytes(bpt.PathTracer.nccl_unique_id() if d.rank == 0 else b"", bpt.NCCL_UNIQUE_ID_BYTES)
        pt.nccl_init(uid, d.rank, d.world)

    # ---- device-resident timed region
    frame = 0
    for _ in range(WU):
        pt.trace(params(frame)); frame += 1
    if d.active:
        pt.allgather_image(W, H)
    pt.sync()
    pt.reset_stats()
    clocks = ClockSampler(range(d.world)) if d.rank == 0 else None   # one node: the job's GPUs are 0..world-1
    time.sleep(0.3 if clocks else 0.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K):
        pt.trace(params(frame)); frame += 1
    if d.active:
        pt.allgather_image(W, H)
    e1.record(stream)
    torch.cuda.synchronize(); d.barrier()
    t1 = time.perf_counter()
    ms_own = e0.elapsed_time(e1)
    ms = d.max(ms_own)
    st = pt.stats()
    # the slowest and the fastest rank's own frames (all-gather excluded): a gap between them is rank imbalance or a
    # slow GPU, not the sharding (every rank gets the same rows modulo 8-row blocks)
    rank_frames_ms = d.gather_floats(st.frame_ms)
    rank_region_ms = d.gather_floats(ms_own)
    if os.environ.get("BPT_BENCH_DEBUG"):
        print(f"[rank {d.rank}] timed region {ms_own:.2f} ms, frames {st.frame_ms:.2f} ms, rays {st.rays_traced}",
              file=sys.stderr, flush=True)
    clk, clk_per_gpu = clocks.stop(t0, t1) if clocks else (None, None)
    rays = d.sum(st.rays_traced)
    paths = d.sum(st.paths)
    value = rays / (ms * 1e-3) / 1e6
    kernel_launches = int(st.kernel_launches)

    # ---- per-kernel timing of the traversal kernel: the same K frames again with ONE sample lane, two CUDA events
    # around every traversal launch on its stream (with several lanes in flight the kernels of different lanes share
    # the SMs and an event pair around one of them also counts the time it waits for the others)
    pt.set_option(bpt.OPT_STREAMS, 1)
    pt.set_option(bpt.OPT_PROFILE, 1)
    pt.reset_stats()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    KT = min(K, 4)  # frames of the kernel-timing region
    k0.record(stream)
    for _ in range(KT):
        pt.trace(params(frame)); frame += 1
    k1.record(stream)
    torch.cuda.synchronize()
    st = pt.stats()
    single_lane_ms = k0.elapsed_time(k1)

    # ---- traversal-kernel roofline (rank 0's kernel): per-ray node/triangle fetch counts from one instrumented frame
    pt.set_option(bpt.OPT_PROFILE, 0)
    pt.set_option(bpt.OPT_COUNT_TRAVERSAL, 1)
    pt.reset_stats()
    pt.trace(params(frame)); frame += 1
    sc = pt.stats()
    pt.set_option(bpt.OPT_COUNT_TRAVERSAL, 0)
    pt.set_option(bpt.OPT_STREAMS, lanes)
    nodes_per_ray = sc.nodes_visited / max(sc.rays_traced, 1)
    tris_per_ray = sc.tris_tested / max(sc.rays_traced, 1)
    # what the quality stage of the build (BPT_OPT_BVH_SAH_SUBTREE, default 32) buys: the same instrumented frame through
    # a plain Morton LBVH, on a second context (rank 0 of a single-GPU run only; soups only — the Cornell box has 36 triangles)
    nodes_per_ray_plain = None
    if d.world == 1 and w["tris"] and not any(o.split("=")[0] == str(bpt.OPT_BVH_SAH_SUBTREE) for o in args.opt):
        pt2 = bpt.PathTracer(d.local_rank, stream.cuda_stream)
        pt2.set_option(bpt.OPT_BVH_SAH_SUBTREE, 0)
        build_scene(pt2)
        pt2.set_option(bpt.OPT_COUNT_TRAVERSAL, 1)
        pt2.reset_stats()
        pt2.trace(params(frame - 1))
        s2 = pt2.stats()
        nodes_per_ray_plain = s2.nodes_visited / max(s2.rays_traced, 1)
        pt2.close()
    # Requested bytes (SURVEY 8d "modelled"): 32 B ray + 16 B hit + one 64 B record per node visited and per triangle
    # tested. Almost all of them are served by L1/L2 (ncu: L2 hit rate 83-90 %), so this is NOT DRAM traffic and is
    # reported under its own name only.
    launches = max(st.trace_launches, 1)
    # Which frame loop ran: the fused path kernel (BPT_OPT_FUSED_PATHS, default) launches the traversal kernel once per
    # sample pass, the wavefront once per bounce of every pass.
    fused = w["depth"] > 1 and launches / KT < w["depth"]
    ray_record_bytes = 0.0 if fused else 48.0
    requested_per_ray = ray_record_bytes + 64.0 * nodes_per_ray + 64.0 * tris_per_ray
    avg_launch_ms = st.trace_kernel_ms / launches
    kernel_s = max(st.trace_kernel_ms * 1e-3, 1e-12)
    rays_per_launch = st.rays_traced / launches
    # Algorithmic DRAM bytes of one launch (the floor): every acceleration-structure record the launch touches fetched
    # from DRAM once — at most the whole record array, at most one record per visit — plus, for the wavefront's traversal
    # launches, every ray record read and every hit record written once (48 B per ray); for the fused path kernel, which
    # has no ray or hit records in HBM, the 64-byte shading record of every hit (at most the whole array, at most one per
    # ray) and the 16-byte colour of every path written once.
    bvh_bytes = float(info.bytes_nodes + info.bytes_tris)
    algo_per_launch = min(bvh_bytes, 64.0 * (nodes_per_ray + tris_per_ray) * rays_per_launch)
    if fused:
        paths_per_launch = float(w["spp"]) * W * (H // d.world) * KT / launches
        algo_per_launch += min(64.0 * max(info.num_tris, 1), 64.0 * rays_per_launch) + 16.0 * paths_per_launch
    else:
        algo_per_launch += 48.0 * rays_per_launch
    achieved = algo_per_launch / max(avg_launch_ms * 1e-3, 1e-12) / 1e9
    peak, peak_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        pass
    cap = ncu_capture(w["name"], d.world)
    traffic = cap["dram_bytes_per_launch"] if cap else None
    roofline = {"kernel": ("k_trace<FUSED> (persistent path kernel: primary rays, BVH8 traversal, shade and bounce of a sample pass)"
                           if fused else "k_trace (persistent BVH8 traversal)"),
                "frame_loop": "fused path kernel, one launch per sample pass" if fused else "wavefront, one launch per bounce",
                # what binds the kernel by ncu (profiles/*_k_trace_ncu_*.txt): instruction issue and the ALU pipe; DRAM
                # throughput is a few per cent of peak. `achieved`/`frac` are the kernel's algorithmic DRAM bytes per
                # launch over its measured duration against the measured HBM peak: far below 1 because the kernel is
                # not an HBM-bound kernel on this machine (the hot top of the BVH lives in the 126 MB L2).
                "bound": "issue", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_per_launch, "compulsory_bytes_per_ray": ray_record_bytes,
                "hbm_frac_ncu": (traffic / max(avg_launch_ms * 1e-3, 1e-12) / 1e9 / peak) if traffic else None,
                "issue_frac": cap.get("issue_active_frac") if cap else None,
                "alu_pipe_frac": cap.get("alu_pipe_frac") if cap else None,
                "l2_hit_rate": cap.get("l2_hit_rate") if cap else None,
                "ncu_capture": ({"file": cap.get("source"), "csrc_hash": cap.get("csrc_hash"), "launches": cap.get("launches")}
                                if cap else f"no capture of csrc {csrc_hash()} / {w['name']} / {d.world} GPU(s) under profiles/"),
                "requested_bytes_per_ray": requested_per_ray,
                "requested_frac": requested_per_ray * st.rays_traced / kernel_s / 1e9 / peak,
                "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                "nodes_per_ray_plain_lbvh": nodes_per_ray_plain,
                "lanes_per_node_step": sc.nodes_visited / max(sc.warp_node_steps, 1),
                "lanes_per_tri_step": sc.tris_tested / max(sc.warp_tri_steps, 1), "avg_launch_ms": avg_launch_ms,
                "launches": int(st.trace_launches), "rays_per_launch": rays_per_launch,
                "mrays_trace_kernel": st.rays_traced / kernel_s / 1e6,
                "timing": f"{KT} frames with one sample lane, CUDA events around every launch on its stream",
                "trace_share_of_step": st.trace_kernel_ms / max(single_lane_ms, 1e-9),
                "ms_per_step_single_lane": single_lane_ms / KT}
    # every rank's own device-to-device copy bandwidth (1 GiB, best of 5), next to its frame times: a rank whose HBM-bound
    # kernels (shade, generate, gather) run slower than its peers' shows here whether its memory is slower as well
    a = torch.empty(1 << 28, dtype=torch.float32, device=f"cuda:{d.local_rank}")
    b = torch.empty_like(a)
    best = 0.0
    with torch.cuda.stream(stream):
        for _ in range(6):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream); b.copy_(a); c1.record(stream)
            torch.cuda.synchronize()
            best = max(best, 2.0 * a.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9)
    del a, b
    ranks = {"frames_ms": rank_frames_ms, "timed_region_ms": rank_region_ms,
             "frames_ms_min": min(rank_frames_ms), "frames_ms_max": max(rank_frames_ms),
             "hbm_copy_gbs": d.gather_floats(best), "gpu_clocks": clk_per_gpu}

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        nt = info.num_tris
        hv, hi, hf = pt.download_mesh(nt)            # the caller's host arrays (untimed preparation)
        pin = [torch.from_numpy(a).pin_memory() for a in (hv, hi, hf)]
        hv, hi, hf = [p.numpy() for p in pin]
        himg = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
        himg_np = himg.numpy()
        himg2 = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()   # the presenter's second frame buffer
        pt.reset_stats()
        d.barrier(); torch.cuda.synchronize()
        tw0 = time.perf_counter()
        e0.record(stream)
        dbg = os.environ.get("BPT_BENCH_DEBUG")
        pt.upload_mesh(hv, hi, hf)                   # H2D: 72 B per triangle
        if dbg: print(f"[rank {d.rank}] upload {1e3 * (time.perf_counter() - tw0):.1f} ms", file=sys.stderr, flush=True)
        if w.get("instances"):
            pt.set_instances(instance_grid(w["instances"]))
        pt.build_accel()
        if dbg: print(f"[rank {d.rank}] +build {1e3 * (time.perf_counter() - tw0):.1f} ms", file=sys.stderr, flush=True)
        host_frames = [himg_np, himg2.numpy()]
        for s in range(K):
            pt.trace(params(s))
            if d.active:
                pt.allgather_image(W, H)
            if d.rank == 0:
                # D2H of the frame the reference would present (one presenter), double-buffered: the copy of frame s runs
                # on the copy stream while frame s+1 is traced; the host only waits for the copy of frame s-1
                pt.read_wait()
                pt.read_image_async(host_frames[s & 1])
            if dbg: print(f"[rank {d.rank}] +step {s} {1e3 * (time.perf_counter() - tw0):.1f} ms", file=sys.stderr, flush=True)
        pt.read_wait()
        pt.sync()
        e1.record(stream)
        torch.cuda.synchronize(); d.barrier()
        tw1 = time.perf_counter()
        ems = d.max(max(e0.elapsed_time(e1), (tw1 - tw0) * 1e3))
        s2 = pt.stats()
        e2e = {"value": d.sum(s2.rays_traced) / (ems * 1e-3) / 1e6, "unit": METRIC,
               "h2d_bytes_per_step": int(nt * 72 / K), "d2h_bytes_per_step": int(W * H * 16),
               "includes": f"mesh upload from pinned host memory + BVH build (once, amortised over {K} steps) + per-step "
                           "bpt_trace" + (" + all-gather" if d.active else "") + " + full-image read-back to pinned host memory (double-buffered, "
                           "overlapping the next step's trace)"
                           + (" on rank 0 (the presenter)" if d.active else ""),
               "ms_total": ems}

    verified, vdetail = None, None
    if d.active:
        verified, vdetail = verify_allgather(bpt, pt, d, w, params, tile, build_scene, stream)

    cpu = None
    if d.rank == 0 and d.world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_baseline(w)

    if d.rank == 0:
        out = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": d.world, "steps": K, "warmup": WU,
               "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "ranks": ranks,
               "config": {"workload": w["name"], "baseline_config": w["config"], "tris": info.num_tris, "instances": info.num_instances, "width": W, "height": H,
                          "spp_per_step": w["spp"], "depth": w["depth"], "sampler": "uniform hemisphere (reference)",
                          "sample_lanes": lanes,
                          "tiling": (f"{tile['tile_block']}-row blocks round-robin over {d.world} GPUs + 1 NCCL all-gather"
                                     if tile else "single tile"),
                          "frame_loop": roofline["frame_loop"],
                          "l2": "working set per step (BVH + shading records + per-path colours, in the wavefront also the path queues) is far larger than the 126 MB L2; no flush needed",
                          "bvh8_nodes": info.num_nodes8, "bvh_bytes": int(info.bytes_nodes + info.bytes_tris),
                          **({"options": args.opt} if args.opt else {})},
               "samples_per_s": paths / (ms * 1e-3), "rays": int(rays), "paths": int(paths), "build_ms": build_ms,
               "gpu_launches": kernel_launches, "clocks": clk, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
               **({"allgather_verified": verified, "allgather_check": vdetail} if d.active else {})}
        _RESULT_LINE.append(json.dumps(out))
    pt.close()
    d.close()


def main(argv=None):
    args = parse_args(argv)
    # stdout carries exactly ONE line, the JSON result: libraries that write to file descriptor 1 on their own
    # (NCCL prints its version banner there) are sent to stderr while the bench runs
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(keep, 1)
        os.close(keep)
    if _RESULT_LINE:
        print(_RESULT_LINE[-1], flush=True)


if __name__ == "__main__":
    main()
