#!/bin/bash
# usage: scripts/gpurun_retry.sh <gpurun args...> -- '<command>'
# Retries while the pod answers "transient" (no box / slot free; nothing charged).
for attempt in $(seq 1 15); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry] attempt $attempt transient; sleeping 150 s"
    sleep 150
    continue
  fi
  break
done
