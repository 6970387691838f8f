#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -n 2 gpurun_out/r2_bench_n1.err
timeout 600 python -m pytest tests/test_bls_gpu.py -m gpu -q 2>&1 | tail -n 2
