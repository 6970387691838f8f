#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_merkle_gpu.py -m gpu -q > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log; tail -n 4 gpurun_out/r2_t7.log
for n in 100000 30000 10000 3000 1024 64; do timeout 300 python scripts/quick_bls_bench.py $n 128 2>&1 | tail -n 1; done
timeout 600 python scripts/quick_incremental_bench.py > gpurun_out/r2_incremental_bench.json 2> gpurun_out/r2_incr.err; python3 -c "
import json; d=json.load(open('gpurun_out/r2_incremental_bench.json'))
for k,v in d.items(): print(k, v if not isinstance(v,dict) else {a:b for a,b in v.items() if a!='note'})"
