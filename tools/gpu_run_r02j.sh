#!/bin/bash
# round-2 pass J (one GPU): FP8 attention after the integer e4m3 expansion (tests, timing, ncu), full GPU suite, benches with
# teacher-forced parity
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 python tools/attn_check.py 32 4400 32 8 12 fp8 > gpurun_out/attn_fp8_r02j.log 2>&1
timeout 120 python tools/attn_check.py 32 4400 32 8 12 >> gpurun_out/attn_fp8_r02j.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_r02j.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02j.log
timeout 600 python bench.py --steps 64 --no-cpu-baseline > gpurun_out/bench_r02j.log 2>&1
timeout 600 python bench.py --config gptq_fp8kv --steps 32 --no-cpu-baseline > gpurun_out/bench_r02j_gptq.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paged_attn_decode_kernel -s 2 -c 1 -f -o gpurun_out/fp8attn_r02j python tools/attn_check.py 32 4400 32 8 2 fp8 > gpurun_out/ncu_fp8attn_r02j.log 2>&1
cat gpurun_out/attn_fp8_r02j.log
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02j.log | head -20
for f in gpurun_out/bench_r02j.log gpurun_out/bench_r02j_gptq.log; do echo "== $f"; tail -1 $f | cut -c1-200; tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('  ms', d['ms_per_step'], 'attn', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'gemm', d['roofline_gemm']['ms_per_launch'], d['roofline_gemm']['frac'], 'parity', d.get('parity'))
except Exception as e: print('ERR', e)
"; done
