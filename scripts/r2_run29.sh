#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_block_root.py -m gpu -q -x 2>&1 | tail -n 15
