#!/bin/bash
# round-2 pass C (one GPU): timeline of the persistent layer kernel, FP8-KV attention on the TMA path, engine tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B200_MEGA_TRACE=3 timeout 300 python tools/mega_trace.py 32 4096 6 > gpurun_out/mega_trace_r02c.log 2>&1
timeout 900 python -m pytest tests/test_attention_gpu.py -q -m gpu > gpurun_out/pytest_r02c_attn.log 2>&1
echo "pytest(attn) exit $?" >> gpurun_out/pytest_r02c_attn.log
timeout 900 python -m pytest tests/test_llama_gpu.py -q -s > gpurun_out/pytest_r02c_llama.log 2>&1
echo "pytest(llama) exit $?" >> gpurun_out/pytest_r02c_llama.log
timeout 600 python bench.py --kv fp8 --no-cpu-baseline --steps 32 > gpurun_out/bench_r02c_fp8.log 2>&1
cat gpurun_out/mega_trace_r02c.log
grep -E "exit|passed|failed|metric shapes|spread|Error|error" gpurun_out/pytest_r02c_attn.log gpurun_out/pytest_r02c_llama.log | head -30
tail -1 gpurun_out/bench_r02c_fp8.log | cut -c1-2600
