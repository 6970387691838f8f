#!/bin/bash
# round-2 validation pass on one B200: whole GPU suite, bench (both arms), launch lists, ncu captures of the latency-mode kernels
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t19.log; tail -n 5 gpurun_out/r2_t19.log
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2_bench_n1.err; cut -c1-300 gpurun_out/r2_bench_n1.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
for n in 100000 1024 64; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bls_$n.csv python scripts/quick_bls_bench.py $n 128 > /dev/null 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg3.csv python scripts/quick_cfg3_bench.py > /dev/null 2>&1
for k in k_miller_warp k_sig_prepare_warp k_hash_to_g2_warp; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r2_ncu_$k -f python scripts/quick_bls_bench.py 64 128 > gpurun_out/r2_ncu19_$k.log 2>&1; tail -n 1 gpurun_out/r2_ncu19_$k.log
done
for n in 100000 10000 3000 1024 64; do timeout 300 python scripts/quick_bls_bench.py $n 128 2>&1 | tail -n 1; done
timeout 300 python scripts/quick_cfg3_bench.py 2>&1 | tail -n 1
