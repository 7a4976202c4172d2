#!/bin/bash
# round-2 pass W (one GPU): static first ticket (no atomic before the first TMA); attention + engine tests, benches
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_llama_gpu.py -q -m gpu -x > gpurun_out/pytest_r02w.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02w.log
run() { tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('   ms/step %.3f  value %.0f  attn %.1f us frac %.3f' % (d['ms_per_step'], d['value'], r['ms_per_launch']*1e3, r['frac']))"; }
for q in 1 2; do echo "gptq_fp8kv queue=$q"; B200_ATTN_STATIC=$q timeout 300 python bench.py --config gptq_fp8kv --steps 48 --no-cpu-baseline --parity-steps 0 2>/dev/null | run; done
echo "q4k"; timeout 300 python bench.py --steps 64 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_r02w.log | run
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02w.log | head
