#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_bls_stages_gpu.py -m gpu -q -x 2>&1 | tail -n 15
for n in 64 1024 3000 10000 100000; do timeout 300 python scripts/quick_bls_bench.py $n 128 2>&1 | tail -n 1; done
timeout 600 python scripts/quick_cfg3_bench.py 2>&1 | tail -n 1
for n in 64 1024; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bls_$n.csv python scripts/quick_bls_bench.py $n 128 > /dev/null 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg3.csv python scripts/quick_cfg3_bench.py > /dev/null 2>&1
