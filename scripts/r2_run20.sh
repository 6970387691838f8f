#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
for n in 592 800 1024 1184 1500; do timeout 300 python scripts/quick_bls_bench.py $n 128 2>&1 | tail -n 1; done
timeout 600 python -m pytest tests/test_bls_gpu.py -m gpu -q -x 2>&1 | tail -n 3
