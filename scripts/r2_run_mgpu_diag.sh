#!/bin/bash
# N-GPU diagnosis of the weak-scaling step time: default, then without the per-step collective
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-8}
export LHB_BENCH_SKIP=cfg0,cfg3,cfg4,state_sharded
for mode in default nocoll; do
  if [ $mode = nocoll ]; then export LHB_BENCH_NO_STEP_COLLECTIVE=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 6 --warmup 3 2> gpurun_out/diag_$mode.err | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$mode', d['value'], d['ms_per_step'], d['clocks'], d['e2e']['ms_per_step'])"
done
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
