#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python scripts/quick_gossip_concurrency.py 1 > gpurun_out/r2_gossip_concurrency.jsonl 2> gpurun_out/r2_gossip.err; cat gpurun_out/r2_gossip_concurrency.jsonl | cut -c1-200
timeout 600 python -m pytest tests/test_bls_gpu.py -m gpu -q -x -k "concurrent or modes_agree or cpp" 2>&1 | tail -n 3
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2_bench_n1.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['cfg0']['gossip_concurrent'], d['cfg3']['ms_per_segment'])"
