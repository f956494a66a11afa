#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box free): tools/gpu/retry.sh <timeout> <script> [gpus]
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then out=$(gpurun --timeout $T -- "bash $S" 2>&1); else out=$(gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
  echo "$out" | tail -100
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then echo "[retry $i] pod busy, sleeping 90 s"; sleep 90; continue; fi
  break
done
