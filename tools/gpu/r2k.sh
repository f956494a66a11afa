#!/bin/bash
# where does one rank of 8 lose time against 1/8 of the single-GPU frame? ncu launch lists of a 32-spp frame: the whole image
# and the tile rank 0 of 8 renders (on one GPU)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2k_launches_full.csv \
  python tools/probe.py --size 4096 --spp 32 --frames 1 --tag full > $O/r2k_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2k_launches_tile8.csv \
  python tools/probe.py --size 4096 --spp 32 --frames 1 --tile 8,0 --tag tile8 > $O/r2k_tile8.log 2>&1
python tools/probe.py --size 4096 --spp 32 --frames 2 --tag full
python tools/probe.py --size 4096 --spp 32 --frames 2 --tile 8,0 --tag tile8
ls -la $O
