#!/bin/bash
# round 2, call X: fused path kernel, second version (vector RED for the colour, chunked path ids, L2 prefetch of the
# shading record at retire, smaller park): bit-identity tests, then A/B probes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "fused or graph_replay" > $O/r2x_pytest_fused.txt 2>&1
echo "pytest exit $?" >> $O/r2x_pytest_fused.txt
tail -8 $O/r2x_pytest_fused.txt
if grep -q "pytest exit 0" $O/r2x_pytest_fused.txt; then
  rm -f $O/r2x_probe.txt
  for t in "fused 13=1" "wavefront 13=0"; do
    set -- $t
    timeout 300 python tools/probe.py --tris 10000000 --size 4096 --spp 8 --frames 2 --tag soup10m_$1 --opt $2 >> $O/r2x_probe.txt 2>&1
    timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 1920 --height 1080 --spp 32 --frames 2 --tag soup1m_$1 --opt $2 >> $O/r2x_probe.txt 2>&1
    timeout 300 python bench.py --workload cornell --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --opt $2 > $O/r2x_bench_cornell_$1.json 2>> $O/r2x_bench.err
  done
  cut -c1-200 $O/r2x_probe.txt
  for f in $O/r2x_bench_cornell*.json; do echo $f; cut -c1-100 $f; done
fi
