#!/bin/bash
# round-2 pass P (one GPU): FP4 linears on the tcgen05 pipeline, dense GEMM at mid-size m
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fp_linear_gpu.py tests/test_dense_gemm_gpu.py -q -m gpu > gpurun_out/pytest_r02p.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02p.log
timeout 300 python tools/dense_check.py 5 2>&1 | head -6 > gpurun_out/dense_r02p.log
timeout 300 python tools/fp4_check.py > gpurun_out/fp4_r02p.log 2>&1
grep -E "exit|passed|failed|Error|error|assert" gpurun_out/pytest_r02p.log | head -20
cat gpurun_out/dense_r02p.log gpurun_out/fp4_r02p.log
