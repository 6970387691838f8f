#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
# guard: the TMA-staged key ingest must not hang — quick subset first, short timeout
timeout 240 python -m pytest tests/test_bls_gpu.py -m gpu -q -x -k "verify_signature_sets_cases or ragged or golden" > gpurun_out/r2_t3a.log 2>&1; rc=$?; echo "guard rc=$rc"; tail -n 3 gpurun_out/r2_t3a.log
if [ $rc -ne 0 ]; then echo "guard failed: falling back to LHB_PK_TMA=0 for the rest"; export LHB_PK_TMA=0; fi
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
tail -n 6 gpurun_out/r2_t3.log
for v in 1 0; do LHB_PK_TMA=$v timeout 300 python scripts/quick_bls_bench.py 100000 128 > gpurun_out/r2_q4_tma$v.log 2>&1; tail -n 1 gpurun_out/r2_q4_tma$v.log; done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2_bench_n1.err; head -c 1500 gpurun_out/r2_bench_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; tail -n 3 gpurun_out/r2_bench_ref.err; head -c 600 gpurun_out/r2_bench_ref.json
