#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_bls_stages_gpu.py -m gpu -q -x 2>&1 | tail -n 15
for n in 64 300 592 1024 100000; do timeout 300 python scripts/quick_bls_bench.py $n 128 2>&1 | tail -n 1; done
for n in 64; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bls_$n.csv python scripts/quick_bls_bench.py $n 128 > /dev/null 2>&1
tail -n 10 gpurun_out/r2_launches_bls_$n.csv | awk -F'","' '{print substr($5,1,40), $(NF-4), $NF}'
done
