#!/bin/bash
# round 2, final call on 1 GPU (after the fused path kernel went in): full GPU suite, smoke, the bench lines of every
# BASELINE.json configuration and of the reference arm, the ncu launch list of the bench command and the full capture of
# ALL traversal launches of one timed frame (4 passes x 8 bounces), summarised on the box (csrc hash of THIS code)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
python -c "import bench; print(bench.csrc_hash())" > $O/r2y_csrc_hash.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2y_smoke.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2y_pytest.txt 2>&1
echo "pytest exit $?" >> $O/r2y_pytest.txt
timeout 900 python bench.py > $O/r2y_bench_soup10m_1gpu_prelim.json 2>> $O/r2y_bench.err
for w in cornell cornell1000 soup1m; do
  timeout 600 python bench.py --workload $w > $O/r2y_bench_$w.json 2>> $O/r2y_bench.err
done
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/r2y_bench_reference_arm.json 2>> $O/r2y_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2y_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2y_bench_under_ncu.log 2>&1
# warm-up frame = launches 0..31, timed frame = launches 32..63
timeout 1500 ncu --set full --clock-control none -k regex:k_trace -s 32 -c 32 -f -o /tmp/r2y_k_trace_all \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r2y_ncu_trace.log 2>&1
ncu -i /tmp/r2y_k_trace_all.ncu-rep --page raw --csv > $O/r2y_k_trace_all_raw.csv 2>/dev/null
cp profiles/k_trace_traffic.json $O/r2y_k_trace_traffic.json
python profiles/summarize.py traffic /tmp/r2y_k_trace_all.ncu-rep soup10m 1 $O/r2y_k_trace_traffic.json > $O/r2y_traffic.log 2>&1
# the headline line once more, now that the capture of this csrc exists (traffic / hbm_frac_ncu filled in)
cp $O/r2y_k_trace_traffic.json profiles/k_trace_traffic.json
timeout 900 python bench.py > $O/r2y_bench_soup10m_1gpu.json 2>> $O/r2y_bench.err
timeout 600 ncu --set full --clock-control none -k regex:k_shade -s 33 -c 2 -f -o $O/r2y_k_shade \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu -i $O/r2y_k_shade.ncu-rep --page raw --csv > $O/r2y_k_shade_raw.csv 2>/dev/null; rm -f $O/r2y_k_shade.ncu-rep
du -sh $O; cat $O/r2y_smoke.txt | tail -2; tail -4 $O/r2y_pytest.txt; cut -c1-300 $O/r2y_bench_soup10m_1gpu.json; tail -3 $O/r2y_bench.err; head -12 $O/r2y_traffic.log
