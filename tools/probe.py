#!/usr/bin/env python
"""Developer probe: traversal-loop statistics and per-bounce timing of the soup workloads on one GPU.

    python tools/probe.py [--tris 10000000] [--size 2048] [--spp 2] [--opt K=V ...]

Prints one JSON line: Mray/s of bpt_trace (device events), and from one instrumented frame the per-ray node /
triangle counts and the SIMD efficiency of the traversal loop (bpt_stats)."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bpt = importlib.import_module("single-file-vulkan-pathtracing_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--seed", type=lambda x: int(x, 0), default=0x5EED0002)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--spp", type=int, default=2)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[], help="OPTION_ID=VALUE passed to bpt_set_option before the build")
    ap.add_argument("--tag", default="")
    ap.add_argument("--height", type=int, default=None, help="image height (default: --size)")
    ap.add_argument("--tile", default=None, help="NRANKS,RANK[,BLOCK]: render only that rank's row blocks (multi-GPU tiling on one GPU)")
    a = ap.parse_args()
    with bpt.PathTracer(0) as pt:
        for o in a.opt:
            k, v = o.split("=")
            pt.set_option(int(k), int(v))
        pt.upload_soup(a.tris, a.seed)
        info = pt.build_accel()
        kw = {}
        if a.tile:
            t = [int(x) for x in a.tile.split(",")]
            kw = dict(tile_nranks=t[0], tile_rank=t[1], tile_block=t[2] if len(t) > 2 else 8)
        p = lambda f: bpt.default_params(a.size, a.height or a.size, a.spp, a.depth, f, **kw)
        pt.trace(p(0)); pt.sync()                      # warm-up
        pt.set_option(bpt.OPT_PROFILE, 1)
        pt.reset_stats()
        for f in range(a.frames):
            pt.trace(p(1 + f))
        st = pt.stats()
        out = {"tag": a.tag, "tris": a.tris, "nodes8": info.num_nodes8, "depth8": info.max_depth8,
               "build_ms": round(st.build_ms, 2),
               "mrays_frame": round(st.rays_traced / st.frame_ms / 1e3, 1),
               "mrays_trace_kernel": round(st.rays_traced / st.trace_kernel_ms / 1e3, 1),
               "trace_share": round(st.trace_kernel_ms / st.frame_ms, 3)}
        pt.set_option(bpt.OPT_PROFILE, 0)
        pt.set_option(bpt.OPT_COUNT_TRAVERSAL, 1)
        pt.reset_stats()
        pt.trace(p(100))
        sc = pt.stats()
        r = max(sc.rays_traced, 1)
        out.update({"nodes_per_ray": round(sc.nodes_visited / r, 2), "tris_per_ray": round(sc.tris_tested / r, 2),
                    "iters_per_ray": round(sc.lane_iterations / r, 2),
                    "lanes_per_node_step": round(sc.nodes_visited / max(sc.warp_node_steps, 1), 2),
                    "lanes_per_tri_step": round(sc.tris_tested / max(sc.warp_tri_steps, 1), 2),
                    "live_lanes_per_iter": round(sc.lane_iterations / max(sc.warp_iterations, 1), 2),
                    "node_steps_per_warp_iter": round(sc.warp_node_steps / max(sc.warp_iterations, 1), 3),
                    "tri_steps_per_warp_iter": round(sc.warp_tri_steps / max(sc.warp_iterations, 1), 3)})
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
