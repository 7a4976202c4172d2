#!/bin/bash
# round-2 pass R (one GPU): CTA-level ticket mailbox for the FP8 attention queue (modes 0 / 1 / 2 at a balanced and an unbalanced context),
# MoE / MLA operator timing
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
L=gpurun_out/attn_r02r.log
: > $L
for ctx in 4400 4664 4150; do for sw in 1 2 0; do echo "== ctx $ctx queue mode $sw" >> $L; B200_ATTN_STATIC=$sw timeout 120 python tools/attn_check.py 32 $ctx 32 8 12 fp8 >> $L 2>&1; done; done
echo "== ragged" >> $L; timeout 120 python tools/attn_check.py 7 333 32 8 4 fp8 >> $L 2>&1
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_llama_gpu.py -q -m gpu -x > gpurun_out/pytest_r02r.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02r.log
timeout 600 python bench.py --config gptq_fp8kv --steps 32 --no-cpu-baseline > gpurun_out/bench_r02r_gptq.log 2>&1
timeout 600 python tools/moe_mla_check.py > gpurun_out/moe_mla_r02r.log 2>&1
cat $L
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02r.log | head
tail -1 gpurun_out/bench_r02r_gptq.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}; g=d.get('roofline_gemm') or {}
    print('  ms', d.get('ms_per_step'), 'value', d.get('value'), 'attn', r.get('ms_per_launch'), r.get('frac'), r.get('traffic'), 'gemm', g.get('ms_per_launch'), g.get('frac'), 'parity', d.get('parity'))
except Exception as e: print('ERR', e)
"
cat gpurun_out/moe_mla_r02r.log
