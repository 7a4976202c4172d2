#!/bin/bash
# round-2 pass L (one GPU): FP8 attention on the register path (no staging tile)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 python tools/attn_check.py 32 4400 32 8 12 fp8 > gpurun_out/attn_fp8_r02l.log 2>&1
timeout 120 python tools/attn_check.py 32 1000 32 8 12 fp8 >> gpurun_out/attn_fp8_r02l.log 2>&1
timeout 120 python tools/attn_check.py 7 333 32 8 12 fp8 >> gpurun_out/attn_fp8_r02l.log 2>&1
timeout 900 python -m pytest tests/test_attention_gpu.py -q -m gpu -x -k "fp8" > gpurun_out/pytest_r02l.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02l.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paged_attn_decode_kernel -s 2 -c 1 -f -o gpurun_out/fp8attn_r02l python tools/attn_check.py 32 4400 32 8 2 fp8 > gpurun_out/ncu_fp8attn_r02l.log 2>&1
cat gpurun_out/attn_fp8_r02l.log
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02l.log | head -20
