#!/bin/bash
# refill policy of the traversal kernel on the tile one rank of 8 renders (32-spp frame): does the launch tail care?
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
for cfg in "30 2" "32 1" "32 2" "28 2" "24 2" "30 4" "16 2"; do
  set -- $cfg
  timeout 300 python tools/probe.py --size 4096 --spp 32 --frames 2 --tile 8,0 --opt 8=$1 --opt 9=$2 --tag tile8_refill$1_steps$2
done
timeout 300 python tools/probe.py --size 4096 --spp 32 --frames 2 --tag full_refill30_steps2
} > gpurun_out/r2r_probe.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2r_probe.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['tag'], d['mrays_frame'], d['mrays_trace_kernel'], d['live_lanes_per_iter'])
PY
