#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fp_linear_gpu.py tests/test_marlin.py tests/test_attention_gpu.py tests/test_qmatmul_gpu.py tests/test_llama_gpu.py -m gpu -q --durations=5 > gpurun_out/pytest_e.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_e.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v14.log 2>&1
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_e.log | tail -30; tail -1 gpurun_out/bench_v14.log | cut -c1-200
