#!/bin/bash
# ncu captures of the traversal kernel for the other workloads (8 launches = the 8 bounces of one timed frame each), summarised
# on the box into gpurun_out/r2p_k_trace_traffic.json (merged into profiles/k_trace_traffic.json afterwards)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
cp profiles/k_trace_traffic.json $O/r2p_k_trace_traffic.json
for w in cornell cornell1000 soup1m; do
  timeout 900 ncu --set full --clock-control none -k regex:k_trace -s 8 -c 8 -f -o /tmp/r2p_$w \
    python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2p_ncu_$w.log 2>&1
  ncu -i /tmp/r2p_$w.ncu-rep --page raw --csv > $O/r2p_k_trace_${w}_raw.csv 2>/dev/null
  python profiles/summarize.py traffic /tmp/r2p_$w.ncu-rep $w 1 $O/r2p_k_trace_traffic.json > $O/r2p_traffic_$w.log 2>&1
  head -20 $O/r2p_traffic_$w.log
done
