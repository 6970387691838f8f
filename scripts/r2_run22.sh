#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_bls_gpu.py -m gpu -q -x 2>&1 | tail -n 3
for n in 100000 10000 3000 64; do timeout 300 python scripts/quick_bls_bench.py $n 128 2>&1 | tail -n 1; done
timeout 300 python scripts/quick_cfg3_bench.py 2>&1 | tail -n 1
timeout 300 python scripts/quick_bls_bench.py 100000 128 2>&1 | tail -n 1
