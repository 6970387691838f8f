#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 120 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 3
