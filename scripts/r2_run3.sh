#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t3.log
tail -n 6 gpurun_out/r2_t3.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2_bench_n1.err; head -c 1500 gpurun_out/r2_bench_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; tail -n 3 gpurun_out/r2_bench_ref.err; head -c 600 gpurun_out/r2_bench_ref.json
