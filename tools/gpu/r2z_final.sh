#!/bin/bash
# round 2, final call on 1 GPU: full GPU suite, the bench lines of every BASELINE.json configuration, the ncu launch list of
# the bench command and the full capture of ALL traversal launches of one timed frame (4 passes x 8 bounces), summarised
# on the box (csrc hash of THIS code) because the .ncu-rep of 32 launches does not fit the 64 MiB that come back
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
python -c "import bench; print(bench.csrc_hash())" > $O/r2z_csrc_hash.txt
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2z_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2z_pytest.txt
timeout 900 python bench.py > $O/r2z_bench_soup10m_1gpu.json 2>> $O/r2z_bench.err
for w in cornell cornell1000 soup1m; do
  timeout 600 python bench.py --workload $w > $O/r2z_bench_$w.json 2>> $O/r2z_bench.err
done
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/r2z_bench_reference_arm.json 2>> $O/r2z_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2z_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2z_bench_under_ncu.log 2>&1
# warm-up frame = launches 0..31, timed frame = launches 32..63
timeout 1500 ncu --set full --clock-control none -k regex:k_trace -s 32 -c 32 -f -o /tmp/r2z_k_trace_all \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2z_ncu_trace.log 2>&1
ncu -i /tmp/r2z_k_trace_all.ncu-rep --page raw --csv > $O/r2z_k_trace_all_raw.csv 2>/dev/null
python profiles/summarize.py traffic /tmp/r2z_k_trace_all.ncu-rep soup10m 1 $O/r2z_k_trace_traffic.json > $O/r2z_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 33 -c 1 -f -o $O/r2z_k_trace_src \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_shade -s 33 -c 2 -f -o $O/r2z_k_shade \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
du -sh $O; tail -4 $O/r2z_pytest.txt; cut -c1-700 $O/r2z_bench_soup10m_1gpu.json; tail -3 $O/r2z_bench.err; cat $O/r2z_traffic.log | head -30
