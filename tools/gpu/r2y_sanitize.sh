#!/bin/bash
# compute-sanitizer over the fused path kernel (shared-memory path slots, done / ready lists, warp-level hand-over)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused" > $O/r2y_memcheck.txt 2>&1
echo "memcheck exit $?" >> $O/r2y_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_path_kernel_on_instanced" > $O/r2y_racecheck.txt 2>&1
echo "racecheck exit $?" >> $O/r2y_racecheck.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused" > $O/r2y_synccheck.txt 2>&1
echo "synccheck exit $?" >> $O/r2y_synccheck.txt
for f in memcheck racecheck synccheck; do tail -n 4 $O/r2y_$f.txt; done
