#!/bin/bash
# bench.py after it learned to report nodes per ray through a plain LBVH next to the default build (SAH stage)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2ag_bench_soup10m.json 2> gpurun_out/r2ag_bench.err
timeout 300 python bench.py --workload soup1m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2ag_bench_soup1m.json 2>> gpurun_out/r2ag_bench.err
timeout 300 python bench.py --workload cornell --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2ag_bench_cornell.json 2>> gpurun_out/r2ag_bench.err
tail -n 3 gpurun_out/r2ag_bench.err
python - <<'PY'
import json
for w in ("soup10m", "soup1m", "cornell"):
    d = json.load(open(f"gpurun_out/r2ag_bench_{w}.json"))
    r = d["roofline"]
    print(w, round(d["value"], 1), r["nodes_per_ray"], r["nodes_per_ray_plain_lbvh"], r["build_ms_plain_lbvh"], d["build_ms"], r["traffic"])
PY
