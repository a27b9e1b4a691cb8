#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <timeout> <script> [--gpus N]   -- retries while the pod answers "busy" (exit code 3)
T=$1; S=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "bash $S"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
