#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_final_warp -s 1 -c 1 -o gpurun_out/r2_ncu_k_final_warp -f python scripts/quick_bls_bench.py 64 128 > gpurun_out/r2_ncu18.log 2>&1; tail -n 1 gpurun_out/r2_ncu18.log
