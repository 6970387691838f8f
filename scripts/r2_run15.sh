#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_bls_stages_gpu.py -m gpu -q -x 2>&1 | tail -n 15
for n in 64 300 600 1024 3000; do timeout 300 python scripts/quick_bls_bench.py $n 128 2>&1 | tail -n 1; done
LHB_MILLER_WARP=0 timeout 300 python scripts/quick_bls_bench.py 1024 128 2>&1 | tail -n 1
for n in 64 1024; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bls_$n.csv python scripts/quick_bls_bench.py $n 128 > /dev/null 2>&1
grep -E "k_miller" gpurun_out/r2_launches_bls_$n.csv | tail -n 1 | awk -F'","' '{print $5, $NF}' | cut -c1-100
done
