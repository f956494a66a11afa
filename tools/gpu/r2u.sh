#!/bin/bash
# round 2, call U: first run of the fused path kernel (BPT_OPT_FUSED_PATHS): bit-identity tests against the wavefront,
# memcheck on a small case, then A/B probes fused (default) vs wavefront (13=0)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
  -k "fused or graph_replay or sample_lanes or cfg1_image_parity or long_frame_loop" > $O/r2u_pytest_fused.txt 2>&1
echo "pytest exit $?" >> $O/r2u_pytest_fused.txt
tail -15 $O/r2u_pytest_fused.txt
if grep -q "pytest exit 0" $O/r2u_pytest_fused.txt; then
  for t in "fused" "wavefront 13=0"; do
    set -- $t
    opt=""; [ -n "$2" ] && opt="--opt $2"
    timeout 300 python tools/probe.py --tris 10000000 --size 4096 --spp 8 --frames 2 --tag soup10m_$1 $opt >> $O/r2u_probe.txt 2>&1
    timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 1920 --height 1080 --spp 32 --frames 2 --tag soup1m_$1 $opt >> $O/r2u_probe.txt 2>&1
    timeout 300 python bench.py --workload cornell --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $opt > $O/r2u_bench_cornell_$1.json 2>> $O/r2u_bench.err
    timeout 300 python bench.py --workload cornell1000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline $opt > $O/r2u_bench_cornell1000_$1.json 2>> $O/r2u_bench.err
  done
  cat $O/r2u_probe.txt | cut -c1-600
  for f in $O/r2u_bench_cornell*.json; do echo $f; cut -c1-160 $f; done
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
    -k "fused_path_kernel_on_instanced" > $O/r2u_memcheck.txt 2>&1
  echo "memcheck exit $?" >> $O/r2u_memcheck.txt
  tail -5 $O/r2u_memcheck.txt
fi
