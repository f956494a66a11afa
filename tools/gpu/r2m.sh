#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_host_app.py -m gpu -q -p no:cacheprovider > gpurun_out/r2m_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2m_pytest.txt
tail -15 gpurun_out/r2m_pytest.txt
python tools/probe.py --size 4096 --spp 8 --frames 3 --tag gridfix_full
