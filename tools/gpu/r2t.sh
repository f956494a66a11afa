#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider > gpurun_out/r2t_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2t_pytest.txt
tail -4 gpurun_out/r2t_pytest.txt
