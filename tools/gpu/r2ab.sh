#!/bin/bash
# round 2, call AB: fused path kernel, third version (second version minus the chunked path ids): tests, frame latency, probes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -p no:cacheprovider -k "fused or graph_replay or invariances" > $O/r2ab_pytest_fused.txt 2>&1
echo "pytest exit $?" >> $O/r2ab_pytest_fused.txt
tail -n 3 $O/r2ab_pytest_fused.txt
timeout 300 python tools/frame_latency.py > $O/r2ab_frame_latency.txt 2>&1
cat $O/r2ab_frame_latency.txt
rm -f $O/r2ab_probe.txt
timeout 300 python tools/probe.py --tris 10000000 --size 4096 --spp 8 --frames 2 --tag soup10m_fused --opt 13=1 >> $O/r2ab_probe.txt 2>&1
timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 1920 --height 1080 --spp 32 --frames 2 --tag soup1m_fused --opt 13=1 >> $O/r2ab_probe.txt 2>&1
cut -c1-200 $O/r2ab_probe.txt
