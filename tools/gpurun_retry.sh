#!/bin/bash
# gpurun with retries while the pool answers "no box / slot right now" (exit 3): tools/gpurun_retry.sh <timeout> '<command>'
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
