#!/bin/bash
# compute-sanitizer over the kernels that are new in round 2 (ring shade, NEE, SAH stage, device front-end, lanes, async read)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
K="next_event or sah or front_end or shade_step or lanes or async or roulette or cfg1 or instanced_image or soup_image or rgba8"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "$K" > $O/r2c_memcheck.txt 2>&1
echo "memcheck exit $?" >> $O/r2c_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "sah or cfg1_image or next_event or lanes" > $O/r2c_racecheck.txt 2>&1
echo "racecheck exit $?" >> $O/r2c_racecheck.txt
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "sah or cfg1_image or next_event or lanes" > $O/r2c_synccheck.txt 2>&1
echo "synccheck exit $?" >> $O/r2c_synccheck.txt
tail -5 $O/r2c_memcheck.txt $O/r2c_racecheck.txt $O/r2c_synccheck.txt
