#!/bin/bash
for env in "" "HEDIT_GEMM_BN=160" "HEDIT_GEMM_BN=256" "HEDIT_GEMM_CLUSTER=1" "HEDIT_GEMM_CLUSTER=1 HEDIT_GEMM_BN=256"; do
  echo "== $env"; env $env timeout 300 python tools/op_bench.py linear --iters 20 2>&1 | tail -6 | cut -c1-90
done
