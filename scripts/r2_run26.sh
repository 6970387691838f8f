#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
echo "== 32 connections"; CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python scripts/quick_gossip_concurrency.py 1 2>&1 | cut -c1-200
