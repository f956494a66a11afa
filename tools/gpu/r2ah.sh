#!/bin/bash
# the final bench.py once more, as the driver calls it (default workload, cpu baseline included)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python bench.py --steps 4 --warmup 3 > gpurun_out/r2ah_bench_soup10m.json 2> gpurun_out/r2ah_bench.err
echo "exit $?"; tail -n 3 gpurun_out/r2ah_bench.err; cut -c1-150 gpurun_out/r2ah_bench_soup10m.json
python -c "
import json; d=json.load(open('gpurun_out/r2ah_bench_soup10m.json')); r=d['roofline']
print(d['value'], d['e2e']['value'], r['traffic'], r['hbm_frac_ncu'], r['nodes_per_ray'], r['nodes_per_ray_plain_lbvh'], d['cpu_baseline']['value'], d['gpu_launches'])"
