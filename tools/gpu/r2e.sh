#!/bin/bash
# round 2, call E: triangle-record prefetch A/B (default = on, variant nopf = off), full suite on the candidate final code,
# bench lines at the 32-spp step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
{
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag pf_full
BPT_LIB_VARIANT=nopf timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag nopf_full
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag pf_full_again
BPT_LIB_VARIANT=nopf timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag nopf_full_again
timeout 300 python tools/probe.py --size 4096 --spp 32 --frames 2 --tile 8,0 --tag pf_tile8_32spp
BPT_LIB_VARIANT=nopf timeout 300 python tools/probe.py --size 4096 --spp 32 --frames 2 --tile 8,0 --tag nopf_tile8_32spp
timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 2048 --spp 8 --frames 3 --tag pf_soup1m
BPT_LIB_VARIANT=nopf timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 2048 --spp 8 --frames 3 --tag nopf_soup1m
} > $O/r2e_probe.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2e_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2e_pytest.txt
for w in cornell cornell1000; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > $O/r2e_bench_$w.json 2>> $O/r2e_bench.err
  BPT_LIB_VARIANT=nopf timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e > $O/r2e_bench_${w}_nopf.json 2>> $O/r2e_bench.err
done
timeout 900 python bench.py --steps 4 --warmup 3 > $O/r2e_bench_soup10m.json 2>> $O/r2e_bench.err
du -sh $O; tail -5 $O/r2e_pytest.txt; cat $O/r2e_probe.txt; tail -3 $O/r2e_bench.err
