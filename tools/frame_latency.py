#!/usr/bin/env python
"""Developer probe: microseconds per frame of small Cornell-box frames (the launch-bound end of the workload range)
through the per-bounce wavefront and through the fused path kernel, with and without CUDA-graph replay.

    python tools/frame_latency.py
"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
bpt = importlib.import_module("single-file-vulkan-pathtracing_b200")


def main():
    import oracle_lib as O  # fixture loader only (tests/golden/cornell_scene.json)
    verts, idx, faces, _ = O.load_cornell_golden()
    with bpt.PathTracer(0) as pt:
        pt.upload_mesh(verts, idx, faces)
        pt.build_accel()
        for (w, h, spp, depth, frames) in ((256, 256, 1, 2, 400), (256, 256, 1, 8, 400), (512, 512, 1, 8, 300),
                                           (1024, 1024, 1, 8, 200), (1024, 1024, 4, 8, 100), (1024, 1024, 32, 8, 20)):
            row = {"width": w, "height": h, "spp": spp, "depth": depth}
            for fused in (0, 1):
                for graph in (0, 1):
                    pt.set_option(bpt.OPT_FUSED_PATHS, fused)
                    pt.set_option(bpt.OPT_USE_GRAPH, graph)
                    pt.clear_image()
                    for f in range(5):
                        pt.trace(bpt.default_params(w, h, spp, depth, f))
                    pt.sync()
                    t0 = time.perf_counter()
                    for f in range(frames):
                        pt.trace(bpt.default_params(w, h, spp, depth, 5 + f))
                    pt.sync()
                    us = 1e6 * (time.perf_counter() - t0) / frames
                    row[("fused" if fused else "wavefront") + ("_graph" if graph else "")] = round(us, 1)
            print(json.dumps(row), flush=True)
        pt.set_option(bpt.OPT_USE_GRAPH, 0)


if __name__ == "__main__":
    main()
