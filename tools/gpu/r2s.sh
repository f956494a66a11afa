#!/bin/bash
# what the driver does at round end, on one box: smoke(), the GPU test suite, the default bench of both arms
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out


SECONDS=0; timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $O/r2s_bench_reference.json 2> $O/r2s_bench_reference.err
echo "reference arm: $SECONDS s"; SECONDS=0; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 3 > $O/r2s_bench_ours.json 2> $O/r2s_bench_ours.err; echo "our arm: $SECONDS s"


cut -c1-300 $O/r2s_bench_reference.json; cut -c1-300 $O/r2s_bench_ours.json
