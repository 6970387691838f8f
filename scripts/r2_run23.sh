#!/bin/bash
# final validation of the round: whole GPU suite, smoke, bench (both arms)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_t23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t23.log; tail -n 5 gpurun_out/r2_t23.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2_bench_n1.err; cut -c1-200 gpurun_out/r2_bench_n1.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2_bench_reference_arm.json
