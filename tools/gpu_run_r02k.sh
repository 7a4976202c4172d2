#!/bin/bash
# round-2 pass K (one GPU): ldmatrix.b8 layout + e4m3 expansion rate probes, FP8 attention with pipelined expansion (both routes),
# GPTQ engine with fused int4 launches
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 60 tools/probes/ldm8 > gpurun_out/ldm8_probe.log 2>&1
B200_FP8_CVT=1 timeout 120 python tools/attn_check.py 32 4400 32 8 12 fp8 > gpurun_out/attn_fp8_r02k.log 2>&1
B200_FP8_CVT=2 timeout 120 python tools/attn_check.py 32 4400 32 8 12 fp8 >> gpurun_out/attn_fp8_r02k.log 2>&1
timeout 600 python -m pytest tests/test_llama_gpu.py tests/test_marlin.py -q -m gpu -x > gpurun_out/pytest_r02k.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02k.log
timeout 600 python bench.py --config gptq_fp8kv --steps 32 --no-cpu-baseline > gpurun_out/bench_r02k_gptq.log 2>&1
cat gpurun_out/ldm8_probe.log
cat gpurun_out/attn_fp8_r02k.log
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02k.log | head -20
for f in gpurun_out/bench_r02k_gptq.log; do echo "== $f"; tail -1 $f | cut -c1-200; tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('  ms', d['ms_per_step'], 'attn', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'gemm', d['roofline_gemm']['ms_per_launch'], d['roofline_gemm']['frac'], 'parity', d.get('parity'))
except Exception as e: print('ERR', e)
"; done
