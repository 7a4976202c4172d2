#!/bin/bash
# round-2 pass Q (one GPU): int4 / FP4 GEMMs with scales fetched one unit ahead
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fp_linear_gpu.py tests/test_marlin.py tests/test_llama_gpu.py -q -m gpu > gpurun_out/pytest_r02q.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02q.log
timeout 300 python tools/fp4_check.py > gpurun_out/fp4_r02q.log 2>&1
timeout 600 python bench.py --config gptq_fp8kv --steps 32 --no-cpu-baseline > gpurun_out/bench_r02q_gptq.log 2>&1
grep -E "exit|passed|failed|Error|error|assert" gpurun_out/pytest_r02q.log | head -20
cat gpurun_out/fp4_r02q.log
for f in gpurun_out/bench_r02q_gptq.log; do echo "== $f"; tail -1 $f | cut -c1-200; tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('  ms', d['ms_per_step'], 'attn', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'gemm', d['roofline_gemm']['ms_per_launch'], d['roofline_gemm']['frac'], 'parity', d.get('parity'))
except Exception as e: print('ERR', e)
"; done
