#!/bin/bash
# round 2, call W: ncu captures of the fused path kernel (source-level on the 1 M soup; counters on the 10 M soup)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
if [ "$1" = "soup1m" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -f -o $O/r2w_fused_src \
  python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 1920 --height 1080 --spp 8 --frames 1 --opt 13=1 > $O/r2w_ncu.log 2>&1
else
timeout 600 ncu --set full --clock-control none -k regex:k_trace -s 1 -c 1 -f -o $O/r2w_fused_10m \
  python tools/probe.py --tris 10000000 --size 4096 --spp 4 --frames 1 --opt 13=1 > $O/r2w_ncu10m.log 2>&1
ncu -i $O/r2w_fused_10m.ncu-rep --page raw --csv > $O/r2w_fused_10m_raw.csv 2>/dev/null
rm -f $O/r2w_fused_10m.ncu-rep
fi
tail -3 $O/r2w_ncu*.log | cut -c1-300
