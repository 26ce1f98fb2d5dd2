#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (nothing is charged for those)
#   tools/gpurun_retry.sh <timeout-seconds> <out-file> <command...>
T=$1; OUT=$2; shift 2
for attempt in $(seq 1 30); do
    /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$OUT" 2>&1
    rc=$?
    if grep -q "status=transient" "$OUT" || [ $rc -eq 3 ]; then
        echo "attempt $attempt: busy, retrying in 90 s" >> "$OUT.retries"
        sleep 90
        continue
    fi
    exit $rc
done
exit 3
