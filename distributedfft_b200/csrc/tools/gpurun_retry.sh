#!/bin/bash
# developer helper: retry a gpurun call while the pod answers "busy" (exit code 3); usage: gpurun_retry.sh <gpurun args...>
for i in $(seq 1 30); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] attempt $i answered busy; sleeping 90 s"
    sleep 90
done
exit 3
