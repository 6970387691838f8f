#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_merkle_snapshot.py tests/test_bls_gpu.py tests/test_merkle_gpu.py -m gpu -q > gpurun_out/r2_t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t5.log; tail -n 5 gpurun_out/r2_t5.log
for v in 1 0; do LHB_PK_TMA=$v timeout 300 python scripts/quick_bls_bench.py 100000 128 > gpurun_out/r2_q5_tma$v.log 2>&1; tail -n 1 gpurun_out/r2_q5_tma$v.log; done
for n in 1024 64; do timeout 300 python scripts/quick_bls_bench.py $n 128 2>&1 | tail -n 1; done
LHB_PK_TMA=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_100k_v5.csv python scripts/quick_bls_bench.py 100000 128 > /dev/null 2>&1
grep -E "k_pk_aggregate|k_fp12_reduce|k_final" gpurun_out/r2_launches_100k_v5.csv | tail -n 6 | cut -c1-200
