#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
LHB_PK_TMA=0 timeout 600 python scripts/quick_cfg3_bench.py 2>&1 | tail -n 1
LHB_PK_TMA=1 timeout 600 python scripts/quick_cfg3_bench.py 2>&1 | tail -n 1
LHB_PK_TMA=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg3_plain.csv python scripts/quick_cfg3_bench.py > /dev/null 2>&1
grep -E "k_pk_aggregate" gpurun_out/r2_launches_cfg3_plain.csv | tail -n 1 | awk -F'","' '{print $5, $NF}' | cut -c1-140
timeout 600 nsys --version 2>&1 | head -n 1
