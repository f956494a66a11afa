#!/bin/bash
# round 2, call AD: k_shade with the next tile's hit / state / path id prefetched into shared memory by cp.async
# (variant build: make VARIANT=spf EXTRA_shade="--fmad=false -DBPT_SHADE_PREFETCH=1") against the kept kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
BPT_LIB_VARIANT=spf timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "cfg1_image or soup_image or instanced_image or determinism or rgba8 or tiling" > $O/r2ad_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2ad_pytest.txt
tail -n 3 $O/r2ad_pytest.txt
rm -f $O/r2ad_probe.txt
for v in "" spf "" spf; do
  BPT_LIB_VARIANT=$v timeout 300 python tools/probe.py --tris 10000000 --size 4096 --spp 8 --frames 3 --tag "soup10m_${v:-kept}" >> $O/r2ad_probe.txt 2>&1
  BPT_LIB_VARIANT=$v timeout 300 python bench.py --workload cornell --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>> $O/r2ad_bench.err | cut -c1-120 >> $O/r2ad_probe.txt
done
cut -c1-170 $O/r2ad_probe.txt
