#!/bin/bash
# round-2 multi-GPU session (run with gpurun --gpus N): library-owned NCCL parity check + bench at N
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-2}
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/mgpu_check.py > gpurun_out/r2_mgpu_check_n$N.log 2>&1; echo "mgpu_check rc=$?"; tail -n 4 gpurun_out/r2_mgpu_check_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"; tail -n 5 gpurun_out/r2_bench_n$N.err; head -c 400 gpurun_out/r2_bench_n$N.json
