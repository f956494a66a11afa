#!/bin/bash
# round 2, call AA: per-lane stack entries in shared memory 8 (default) / 6 / 4 — fewer entries = smaller shared-memory
# carve-out = more L1 (git apply tools/experiments/smem_stack_entries.patch; make VARIANT=ssN EXTRA_ALL=-DBPT_TRACE_SSTACK=N)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r2aa_probe.txt
for v in "" ss6 ss4; do
  BPT_LIB_VARIANT=$v timeout 300 python tools/probe.py --tris 10000000 --size 4096 --spp 8 --frames 2 --tag "soup10m_sstack_${v:-8}" >> $O/r2aa_probe.txt 2>&1
  BPT_LIB_VARIANT=$v timeout 300 python tools/probe.py --tris 1000000 --seed 0x5EED0001 --size 1920 --height 1080 --spp 32 --frames 2 --tag "soup1m_sstack_${v:-8}" >> $O/r2aa_probe.txt 2>&1
done
cut -c1-330 $O/r2aa_probe.txt
