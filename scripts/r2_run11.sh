#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_block_root.py tests/test_bls_gpu.py -m gpu -q -x 2>&1 | tail -n 30
timeout 600 python scripts/quick_keyimport_bench.py 16384 500000
