#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_bls_gpu.py -m gpu -q -x 2>&1 | tail -n 12
python scripts/mode_probe.py | cut -c1-100
