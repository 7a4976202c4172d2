#!/bin/bash
# round-2 pass N (one GPU): static queue walk (neighbouring kv heads side by side) for FP8 and 16-bit KV
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
L=gpurun_out/attn_r02n.log
: > $L
for sw in 1 0; do for kv in fp8 ""; do for cp in 0 4; do echo "== static $sw kv '$kv' chunk pages $cp" >> $L; B200_ATTN_STATIC=$sw B200_ATTN_CHUNK_PAGES=$cp timeout 120 python tools/attn_check.py 32 4400 32 8 12 $kv >> $L 2>&1; done; done; done
cp candle-vllm_b200/libb200backend.so /tmp/lib_default.so
cp candle-vllm_b200/build/variants/lib_w8s2.so candle-vllm_b200/libb200backend.so
for cp in 0 4; do echo "== variant w8s2 static chunk pages $cp" >> $L; B200_ATTN_CHUNK_PAGES=$cp timeout 120 python tools/attn_check.py 32 4400 32 8 12 fp8 >> $L 2>&1; done
cp /tmp/lib_default.so candle-vllm_b200/libb200backend.so
timeout 900 python -m pytest tests/test_attention_gpu.py -q -m gpu -x > gpurun_out/pytest_r02n.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02n.log
cat $L
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02n.log | head -20
