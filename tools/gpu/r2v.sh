#!/bin/bash
# round 2, call V: frame latency of small frames, wavefront vs fused path kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/frame_latency.py > gpurun_out/r2v_frame_latency.txt 2>&1
cat gpurun_out/r2v_frame_latency.txt
