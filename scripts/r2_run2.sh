#!/bin/bash
# round-2 GPU session: parity + timing of the cooperative Miller kernel (Karatsuba 3-sum f-updates), small-batch launch lists
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_bls_stages_gpu.py -m gpu -q > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t2.log
tail -n 5 gpurun_out/r2_t2.log
for n in 100000 10000 1024 64; do
  timeout 600 python scripts/quick_bls_bench.py $n 128 > gpurun_out/r2_q3_$n.log 2>&1; tail -n 1 gpurun_out/r2_q3_$n.log
done
for n in 1024 64; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_${n}_v3.csv python scripts/quick_bls_bench.py $n 128 > /dev/null 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_miller_coop -c 1 -f -o gpurun_out/r2_miller_coop_v3 python scripts/quick_bls_bench.py 100000 128 > gpurun_out/r2_ncu3.log 2>&1
ls gpurun_out | tail -n 5
