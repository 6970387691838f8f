#!/bin/bash
# round-2 GPU session: full parity suite, per-kernel ncu captures (counters for bench.py's roofline), warm-path bench
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log
tail -n 8 gpurun_out/r2_t4.log
for k in k_sig_prepare k_hash_to_g2 k_pk_aggregate_tma k_miller_coop; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip 1 -c 1 -f -o gpurun_out/r2_ncu_$k python scripts/quick_bls_bench.py 100000 128 > gpurun_out/r2_ncu_$k.log 2>&1
done
LHB_PK_TMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_pk_aggregate" --launch-skip 1 -c 1 -f -o gpurun_out/r2_ncu_k_pk_aggregate_plain python scripts/quick_bls_bench.py 100000 128 > gpurun_out/r2_ncu_pkplain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_validator_roots" --launch-skip 2 -c 1 -f -o gpurun_out/r2_ncu_k_validator_roots python scripts/quick_merkle_bench.py > gpurun_out/r2_ncu_vr.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bls_100k_final.csv python scripts/quick_bls_bench.py 100000 128 > /dev/null 2>&1
timeout 600 python scripts/quick_incremental_bench.py > gpurun_out/r2_incremental_bench.json 2> gpurun_out/r2_incr.err; cat gpurun_out/r2_incremental_bench.json | cut -c1-1500
ls -la gpurun_out | grep r2_ncu | head -20
