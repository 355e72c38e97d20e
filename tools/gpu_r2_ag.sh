#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python tools/latency_breakdown.py 1000 1500 2000 2>&1 | grep -E "n_ids|solve kernel|post\.|raw C"
