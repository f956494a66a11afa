#!/bin/bash
# round 2, call C: the whole GPU suite on the final kernels, A/B of the ring shade kernel and of the eager triangle-record load,
# bench lines (soup10m headline, Cornell, Cornell x1000, soup1m), ncu evidence
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $O/r2c_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2c_pytest.txt
{
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag default_full
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tile 8,0 --tag default_tile8
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --opt 13=0 --tag noring_full
timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tile 8,0 --opt 13=0 --tag noring_tile8
BPT_LIB_VARIANT=lazyuv timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tag lazyuv_full
BPT_LIB_VARIANT=lazyuv timeout 300 python tools/probe.py --size 4096 --spp 8 --frames 3 --tile 8,0 --tag lazyuv_tile8
} > $O/r2c_probe.txt 2>&1
for w in cornell cornell1000 soup1m; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > $O/r2c_bench_$w.json 2>> $O/r2c_bench.err
done
timeout 300 python bench.py --workload cornell --no-e2e --no-cpu-baseline --opt 13=0 > $O/r2c_bench_cornell_noring.json 2>> $O/r2c_bench.err
timeout 600 python bench.py --steps 8 --warmup 3 > $O/r2c_bench_soup10m.json 2>> $O/r2c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2c_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_trace -s 8 -c 8 -f -o /tmp/r2c_k_trace_all \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2c_ncu_trace.log 2>&1
ncu -i /tmp/r2c_k_trace_all.ncu-rep --page raw --csv > $O/r2c_k_trace_all_raw.csv 2>/dev/null
cp /tmp/r2c_k_trace_all.ncu-rep $O/r2c_k_trace_all.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 9 -c 1 -f -o $O/r2c_k_trace_src \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 9 -c 2 -f -o $O/r2c_k_shade \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2c_ncu_shade.log 2>&1
du -sh $O; tail -15 $O/r2c_pytest.txt; cat $O/r2c_probe.txt
