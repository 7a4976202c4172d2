#!/bin/bash
# round-2 pass M (one GPU): FP8 attention tuning -- deferred queue claims, chunk size sweep, warps x stages variants
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
L=gpurun_out/attn_fp8_r02m.log
: > $L
for cp in 0 2 4 8; do echo "== 8 warps x 3 stages, chunk pages $cp (0 = heuristic)" >> $L; B200_ATTN_CHUNK_PAGES=$cp timeout 120 python tools/attn_check.py 32 4400 32 8 12 fp8 >> $L 2>&1; done
echo "== ctx 4150" >> $L; timeout 120 python tools/attn_check.py 32 4150 32 8 12 fp8 >> $L 2>&1
cp candle-vllm_b200/libb200backend.so /tmp/lib_default.so
for v in w9s3 w10s2 w8s2; do
  cp candle-vllm_b200/build/variants/lib_$v.so candle-vllm_b200/libb200backend.so
  for cp in 4 8; do echo "== variant $v chunk pages $cp" >> $L; B200_ATTN_CHUNK_PAGES=$cp timeout 120 python tools/attn_check.py 32 4400 32 8 12 fp8 >> $L 2>&1; done
done
cp /tmp/lib_default.so candle-vllm_b200/libb200backend.so
timeout 900 python -m pytest tests/test_attention_gpu.py -q -m gpu -x > gpurun_out/pytest_r02m.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02m.log
cat $L
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02m.log | head -20
