#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python scripts/quick_gossip_concurrency.py 1 > gpurun_out/r2_gossip_concurrency.jsonl 2> gpurun_out/r2_gossip.err; cat gpurun_out/r2_gossip_concurrency.jsonl; tail -n 3 gpurun_out/r2_gossip.err
