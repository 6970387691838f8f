#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python scripts/quick_cfg3_bench.py 2>&1 | tail -n 3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg3.csv python scripts/quick_cfg3_bench.py > /dev/null 2>&1
tail -n 12 gpurun_out/r2_launches_cfg3.csv | awk -F'","' '{print $5, $NF}' | cut -c1-140
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_keyimport.csv python scripts/quick_keyimport_bench.py 500000 > /dev/null 2>&1
grep -E "decompress_validate" gpurun_out/r2_launches_keyimport.csv | tail -n 2 | awk -F'","' '{print $5, $NF}' | cut -c1-140
