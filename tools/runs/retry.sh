#!/bin/bash
# usage: tools/runs/retry.sh <timeout> <script> [gpus]   -- retries a gpurun call while the pod answers "transient" / busy
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 12); do
    if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "bash $S" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" 2>&1); fi
    echo "$OUT" | tail -60
    if echo "$OUT" | grep -q "status=transient\|rc=3\|no box\|busy"; then sleep 150; continue; fi
    break
done
