#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
echo "== lane G2 + lane Miller"; LHB_G2_WARP=0 LHB_MILLER_WARP=0 timeout 300 python scripts/quick_gossip_concurrency.py 1 2>&1 | cut -c1-200
echo "== lane G2 + warp Miller"; LHB_G2_WARP=0 timeout 300 python scripts/quick_gossip_concurrency.py 1 2>&1 | cut -c1-200
