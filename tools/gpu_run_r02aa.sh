#!/bin/bash
# round-2 pass AA (one GPU): 128-byte TMA swizzle for the int4 / FP4 weight tiles (8-way LDS bank conflicts before)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_marlin.py tests/test_fp_linear_gpu.py tests/test_llama_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python tools/fp4_check.py 2>&1 | grep -E "nvfp4|mxfp4" | head -6
timeout 600 python bench.py --config gptq_fp8kv --steps 48 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_r02aa_gptq.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; g=d['roofline_gemm']; print('gptq: ms/step %.3f  value %.0f  e2e %.0f attn %.1f us (%.3f)  gemm chain %.3f ms (%.3f) parity %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['ms_per_launch']*1e3, r['frac'], g['ms_per_launch'], g['frac'], (d.get('parity') or {}).get('within_tolerance')))"
