#!/bin/bash
# round-2 GPU session 1: parity of the cooperative Miller kernel, A/B timing, ncu capture
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_bls_stages_gpu.py tests/test_shuffle.py -m gpu -q > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
tail -5 gpurun_out/r2_t1.log
for n in 100000 10000 1024 64; do
  LHB_MILLER_COOP=1 timeout 600 python scripts/quick_bls_bench.py $n 128 > gpurun_out/r2_q_coop_$n.log 2>&1
  LHB_MILLER_COOP=0 timeout 600 python scripts/quick_bls_bench.py $n 128 > gpurun_out/r2_q_old_$n.log 2>&1
  tail -n 2 gpurun_out/r2_q_coop_$n.log; tail -n 2 gpurun_out/r2_q_old_$n.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_100k_v2.csv python scripts/quick_bls_bench.py 100000 128 > gpurun_out/r2_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_miller_coop -c 1 -f -o gpurun_out/r2_miller_coop_v2 python scripts/quick_bls_bench.py 100000 128 > gpurun_out/r2_ncu2.log 2>&1
ls -la gpurun_out | tail -5
