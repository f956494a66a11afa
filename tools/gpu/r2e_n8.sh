#!/bin/bash
# 8 GPUs: the scaling line the driver will take (N = 8, then 4 on the same box), and Cfg5 at 8 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2e_env.txt
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 4 --warmup 3 > $O/r2e_bench_soup10m_${n}gpu.json 2> $O/r2e_bench_${n}gpu.err
cut -c1-1200 $O/r2e_bench_soup10m_${n}gpu.json
done
timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline > $O/r2e_bench_soup10m_1gpu.json 2> $O/r2e_bench_1gpu.err
cut -c1-600 $O/r2e_bench_soup10m_1gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --workload cornell1000 --steps 4 --warmup 3 > $O/r2e_bench_cornell1000_8gpu.json 2> $O/r2e_bench_c1000.err
cut -c1-600 $O/r2e_bench_cornell1000_8gpu.json
