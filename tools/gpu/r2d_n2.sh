#!/bin/bash
# 2 GPUs: the NCCL data path on hardware (2-rank tests), then the bench at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2d_env.txt
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s -p no:cacheprovider > $O/r2d_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2d_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > $O/r2d_bench_soup10m_2gpu.json 2> $O/r2d_bench_2gpu.err
tail -3 $O/r2d_pytest.txt; cut -c1-1500 $O/r2d_bench_soup10m_2gpu.json; tail -5 $O/r2d_bench_2gpu.err
