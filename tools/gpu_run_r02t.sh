#!/bin/bash
# round-2 pass T (one GPU): chunk size decided in the kernel from the device-side context lengths
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
L=gpurun_out/attn_r02t.log
: > $L
for ctx in 4096 4608 4664 5120; do for ad in 1 0; do echo "== fp8 ctx $ctx adaptive $ad" >> $L; B200_ATTN_ADAPTIVE=$ad timeout 120 python tools/attn_check.py 32 $ctx 32 8 12 fp8 2>&1 | tail -2 >> $L; done; done
for ctx in 4096 4664 5120; do for ad in 1 0; do echo "== bf16 ctx $ctx adaptive $ad" >> $L; B200_ATTN_ADAPTIVE=$ad timeout 120 python tools/attn_check.py 32 $ctx 32 8 12 2>&1 | tail -2 >> $L; done; done
echo "== bf16 TP8 shard (4 heads / 1 kv head) ctx 4400" >> $L; timeout 120 python tools/attn_check.py 32 4400 4 1 12 2>&1 | tail -2 >> $L
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_llama_gpu.py -q -m gpu -x > gpurun_out/pytest_r02t.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02t.log
timeout 600 python bench.py --steps 64 --no-cpu-baseline > gpurun_out/bench_r02t.log 2>&1
timeout 600 python bench.py --config gptq_fp8kv --steps 32 --no-cpu-baseline > gpurun_out/bench_r02t_gptq.log 2>&1
cat $L
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02t.log | head
for f in gpurun_out/bench_r02t.log gpurun_out/bench_r02t_gptq.log; do tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}; g=d.get('roofline_gemm') or {}
    print('  ms', d.get('ms_per_step'), 'value', d.get('value'), 'attn', r.get('ms_per_launch'), r.get('frac'), r.get('traffic'), 'gemm', g.get('ms_per_launch'), g.get('frac'), 'parity', (d.get('parity') or {}).get('logits_max_err'), (d.get('parity') or {}).get('within_tolerance'))
except Exception as e: print('ERR', e)
"; done
