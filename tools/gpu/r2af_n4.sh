#!/bin/bash
# 4 GPUs, final code: the scaling line at N = 4 (4 steps) and the 1-GPU line of the same box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu-baseline > $O/r2af_bench_soup10m_4gpu.json 2> $O/r2af_bench_4gpu.err
cut -c1-400 $O/r2af_bench_soup10m_4gpu.json
timeout 300 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2af_bench_soup10m_1gpu.json 2> $O/r2af_bench_1gpu.err
cut -c1-200 $O/r2af_bench_soup10m_1gpu.json
