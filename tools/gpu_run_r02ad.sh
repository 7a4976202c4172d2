#!/bin/bash
# round-2 pass AD (one GPU): int4 scale fetch without the integer division -- parity (GPTQ sym, AWQ zero points, engine) and config 3 step
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_marlin.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python -m pytest tests/test_llama_gpu.py -q -m gpu -x -k "linear_kinds or other" 2>&1 | tail -2
timeout 300 python bench.py --config gptq_fp8kv --steps 48 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_r02ad_gptq.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; g=d['roofline_gemm']; print('gptq: ms/step %.3f  value %.0f  e2e %.0f attn %.1f us (%.3f)  gemm chain %.3f ms (%.3f) parity %s %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['ms_per_launch']*1e3, r['frac'], g['ms_per_launch'], g['frac'], (d.get('parity') or {}).get('logits_max_err'), (d.get('parity') or {}).get('within_tolerance')))"
