#!/bin/bash
# round 2, the final call on 1 GPU (third version of the fused path kernel; the default frame loop is the one r2y measured):
# full GPU suite, smoke, ncu launch list of the bench command, full ncu capture of the 8 traversal launches of one sample
# pass of a timed frame (all 8 bounces) and of the other workloads, summarised on the box (csrc hash of THIS code), then
# the bench lines of every BASELINE.json configuration and of the reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
python -c "import bench; print(bench.csrc_hash())" > $O/r2ac_csrc_hash.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2ac_smoke.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2ac_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2ac_pytest.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2ac_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2ac_bench_under_ncu.log 2>&1
cp profiles/k_trace_traffic.json $O/r2ac_k_trace_traffic.json
# warm-up frame = launches 0..31, timed frame = launches 32..63; its first pass = 32..39
timeout 1500 ncu --set full --clock-control none -k regex:k_trace -s 32 -c 8 -f -o /tmp/r2ac_soup10m \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2ac_ncu_soup10m.log 2>&1
ncu -i /tmp/r2ac_soup10m.ncu-rep --page raw --csv > $O/r2ac_k_trace_soup10m_raw.csv 2>/dev/null
python profiles/summarize.py traffic /tmp/r2ac_soup10m.ncu-rep soup10m 1 $O/r2ac_k_trace_traffic.json > $O/r2ac_traffic_soup10m.log 2>&1
for w in cornell cornell1000 soup1m; do
  timeout 900 ncu --set full --clock-control none -k regex:k_trace -s 8 -c 8 -f -o /tmp/r2ac_$w \
    python bench.py --workload $w --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2ac_ncu_$w.log 2>&1
  ncu -i /tmp/r2ac_$w.ncu-rep --page raw --csv > $O/r2ac_k_trace_${w}_raw.csv 2>/dev/null
  python profiles/summarize.py traffic /tmp/r2ac_$w.ncu-rep $w 1 $O/r2ac_k_trace_traffic.json > $O/r2ac_traffic_$w.log 2>&1
done
cp $O/r2ac_k_trace_traffic.json profiles/k_trace_traffic.json
timeout 900 python bench.py > $O/r2ac_bench_soup10m_1gpu.json 2>> $O/r2ac_bench.err
for w in cornell cornell1000 soup1m; do
  timeout 600 python bench.py --workload $w > $O/r2ac_bench_$w.json 2>> $O/r2ac_bench.err
done
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/r2ac_bench_reference_arm.json 2>> $O/r2ac_bench.err
du -sh $O; tail -n 2 $O/r2ac_smoke.txt; tail -n 3 $O/r2ac_pytest.txt; cut -c1-200 $O/r2ac_bench_soup10m_1gpu.json; tail -n 3 $O/r2ac_bench.err; head -n 8 $O/r2ac_traffic_soup10m.log
