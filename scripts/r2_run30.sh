#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
echo "== lane G2, warp Miller, 32 queues"; LHB_G2_WARP=0 timeout 300 python scripts/quick_gossip_concurrency.py 1 2>&1 | cut -c1-160
echo "== lane G2, coop Miller"; LHB_G2_WARP=0 LHB_MILLER_WARP=0 timeout 300 python scripts/quick_gossip_concurrency.py 1 2>&1 | tail -n 3 | cut -c1-160
