#!/bin/bash
# gpurun with retries on "no box / slot free right now" (exit code 3) and transient refusals.
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    verdict=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status',''))" 2>/dev/null)
    if [ "$rc" != "3" ] && [ "$verdict" != "transient" ]; then exit $rc; fi
    echo "[gpurun_retry] attempt $attempt: rc=$rc status=$verdict; retrying in 90 s" >&2
    sleep 90
done
exit 3
