#!/bin/bash
# round 2, call D: ring shade kernel after the fixes (no RED, pipelined compaction atomic) vs the plain kernel, 32-spp steps
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "ring or cfg1 or cfg2 or next_event or roulette or shade_step or instanced or rgba8 or determinism" > $O/r2d_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2d_pytest.txt
{
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag default_full
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tile 8,0 --tag default_tile8
timeout 300 python tools/probe.py --size 4096 --spp 32 --frames 2 --tile 8,0 --tag default_tile8_32spp
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --opt 13=0 --tag noring_full
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tile 8,0 --opt 13=0 --tag noring_tile8
} > $O/r2d_probe.txt 2>&1
for w in cornell cornell1000; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e > $O/r2d_bench_$w.json 2>> $O/r2d_bench.err
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --opt 13=0 > $O/r2d_bench_${w}_noring.json 2>> $O/r2d_bench.err
done
timeout 900 python bench.py --steps 4 --warmup 3 > $O/r2d_bench_soup10m.json 2>> $O/r2d_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 2 -f -o $O/r2d_k_shade \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --spp 8 > $O/r2d_ncu_shade.log 2>&1
du -sh $O; tail -5 $O/r2d_pytest.txt; cat $O/r2d_probe.txt; tail -3 $O/r2d_bench.err
