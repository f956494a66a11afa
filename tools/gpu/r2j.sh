#!/bin/bash
# e2e timeline of the bench on 1 GPU (where do the ~150 ms outside the frames go?)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
BPT_BENCH_DEBUG=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
grep "rank 0" gpurun_out/r2j_bench.err; cut -c1-400 gpurun_out/r2j_bench.json
