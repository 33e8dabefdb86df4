#!/bin/bash
# gpurun_retry.sh TIMEOUT 'command' [--gpus N]: retry while the pod answers "transient" (nothing charged), 3 min apart.
t=$1; cmd=$2; shift 2
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t "$@" -- "$cmd" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 180; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after 12 transient answers"
