#!/usr/bin/env bash
# Runs a gpurun call, retrying while the pod answers "no box or slot free" (exit code 3: nothing was charged).
#   tools/gpu_retry.sh <timeout-seconds> [--gpus N] -- '<command>'
limit=$1; shift
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout "$limit" "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
