#!/bin/bash
# round-2 pass U (one GPU): chunk decision folded into the existing prologue passes (no extra barriers) -- engine step time check
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -m gpu -x > gpurun_out/pytest_r02u.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02u.log
timeout 600 python bench.py --steps 64 --no-cpu-baseline > gpurun_out/bench_r02u.log 2>&1
timeout 600 python bench.py --config gptq_fp8kv --steps 32 --no-cpu-baseline > gpurun_out/bench_r02u_gptq.log 2>&1
for ctx in 4664 4096; do timeout 120 python tools/attn_check.py 32 $ctx 32 8 12 fp8 2>&1 | tail -1; done
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02u.log | head
for f in gpurun_out/bench_r02u.log gpurun_out/bench_r02u_gptq.log; do tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}; g=d.get('roofline_gemm') or {}
    print('  ms', d.get('ms_per_step'), 'value', d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'attn', r.get('ms_per_launch'), r.get('frac'), 'gemm', g.get('ms_per_launch'), g.get('frac'), 'parity', (d.get('parity') or {}).get('logits_max_err'), (d.get('parity') or {}).get('within_tolerance'))
except Exception as e: print('ERR', e)
"; done
