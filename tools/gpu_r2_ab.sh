#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for sm in -1 100000; do
echo "== CHI_B200_SPLIT_MAX=$sm"
if [ $sm = -1 ]; then timeout 600 python tools/latency_breakdown.py 2000 3000 5000 10000 2>&1 | grep -E "n_ids|raw C|solve kernel"; else CHI_B200_SPLIT_MAX=$sm timeout 600 python tools/latency_breakdown.py 3000 5000 10000 2>&1 | grep -E "n_ids|raw C|solve kernel"; fi
done
