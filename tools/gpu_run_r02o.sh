#!/bin/bash
# round-2 pass O (one GPU): M >= 128 GEMM and prefill attention throughput, tensor-pipe evidence for the dense GEMM
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python tools/dense_check.py 5 > gpurun_out/dense_r02o.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 40 -c 1 -f -o gpurun_out/dense_r02o python tools/dense_check.py 1 > gpurun_out/ncu_dense_r02o.log 2>&1
cat gpurun_out/dense_r02o.log
tail -3 gpurun_out/ncu_dense_r02o.log
