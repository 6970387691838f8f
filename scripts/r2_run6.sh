#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bls_gpu.py tests/test_bls_stages_gpu.py -m gpu -q > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6.log; tail -n 4 gpurun_out/r2_t6.log
for n in 100000 10000 1024 64; do timeout 300 python scripts/quick_bls_bench.py $n 128 2>&1 | tail -n 1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_100k_v6.csv python scripts/quick_bls_bench.py 100000 128 > /dev/null 2>&1
grep -E "k_miller_coop|k_fp12_reduce|k_final" gpurun_out/r2_launches_100k_v6.csv | tail -n 4 | awk -F'","' '{print $5, $NF}' | cut -c1-160
