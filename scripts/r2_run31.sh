#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bls_gpu.py -m gpu -q -x -k "streamed or full_size or concurrent" 2>&1 | tail -n 4
export LHB_BENCH_SKIP=cfg0,cfg3,cfg4
timeout 900 python bench.py 2> gpurun_out/r2_b31.err | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'pageable', d['e2e_pageable']['value'], d['e2e_pageable']['ms_per_step'])"
LHB_STAGE_PAGEABLE=0 timeout 900 python bench.py 2>> gpurun_out/r2_b31.err | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('driver staging: pageable', d['e2e_pageable']['value'], d['e2e_pageable']['ms_per_step'])"
tail -n 3 gpurun_out/r2_b31.err
