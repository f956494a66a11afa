#!/bin/bash
# round 2, call A: full GPU test suite, lanes A/B (whole image and the tile one of 8 ranks renders), bench, ncu captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
(nvidia-smi -L; nproc; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.limit --format=csv) > $O/r2a_env.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $O/r2a_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2a_pytest.txt
for lanes in 1 2 3 4; do
  timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --opt 4=$lanes --tag full_l$lanes
  timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tile 8,0 --opt 4=$lanes --tag tile8_l$lanes
done > $O/r2a_probe.txt 2>&1
timeout 600 python bench.py --steps 8 --warmup 3 > $O/r2a_bench_soup10m.json 2> $O/r2a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2a_bench_under_ncu.log 2>&1
# all 16 traversal launches (2 lanes x 8 bounces) of the timed frame: launches 0..15 are the warm-up frame
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 16 -c 16 -f -o $O/r2a_k_trace \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2a_ncu_trace.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 16 -c 4 -f -o $O/r2a_k_shade \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2a_ncu_shade.log 2>&1
ls -la $O | tail -20
tail -5 $O/r2a_pytest.txt
cat $O/r2a_probe.txt
