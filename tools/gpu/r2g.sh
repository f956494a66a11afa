#!/bin/bash
# round 2, call G: 2 CTAs x 576 threads (36 warps per SM at 56 registers) against 1 x 1024 (32 warps at 64) for the traversal kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
{
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag b1024_full
BPT_LIB_VARIANT=b576 timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag b576_full
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag b1024_full_again
BPT_LIB_VARIANT=b576 timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag b576_full_again
timeout 300 python tools/probe.py --size 4096 --spp 32 --frames 2 --tile 8,0 --tag b1024_tile8_32spp
BPT_LIB_VARIANT=b576 timeout 300 python tools/probe.py --size 4096 --spp 32 --frames 2 --tile 8,0 --tag b576_tile8_32spp
timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 2048 --spp 8 --frames 3 --tag b1024_soup1m
BPT_LIB_VARIANT=b576 timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 2048 --spp 8 --frames 3 --tag b576_soup1m
} > $O/r2g_probe.txt 2>&1
BPT_LIB_VARIANT=b576 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "soup or tiny or instanced_trace or random_rays" > $O/r2g_pytest_b576.txt 2>&1
cat $O/r2g_probe.txt; tail -3 $O/r2g_pytest_b576.txt
