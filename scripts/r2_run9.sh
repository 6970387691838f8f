#!/bin/bash
# A/B of the K<=2 unrolled sum-of-products rows (variants/ holds the two builds), then ncu of k_miller_coop on the winner
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in rolled unroll2 rolled unroll2; do
  echo "== $v"; LHB200_LIB_PATH=$PWD/variants/lib_$v.so timeout 300 python scripts/quick_bls_bench.py 100000 128 2>&1 | tail -n 1
done
for v in rolled unroll2; do
  echo "== $v 1024"; LHB200_LIB_PATH=$PWD/variants/lib_$v.so timeout 300 python scripts/quick_bls_bench.py 1024 128 2>&1 | tail -n 1
done
LHB200_LIB_PATH=$PWD/variants/lib_unroll2.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_miller_coop -s 1 -c 1 -o gpurun_out/r2_miller_unroll2 -f python scripts/quick_bls_bench.py 100000 128 > gpurun_out/r2_ncu9.log 2>&1; tail -n 2 gpurun_out/r2_ncu9.log
LHB200_LIB_PATH=$PWD/variants/lib_unroll2.so timeout 600 python -m pytest tests/test_bls_gpu.py -m gpu -q -x 2>&1 | tail -n 3
