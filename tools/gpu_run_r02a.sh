#!/bin/bash
# round-2 first GPU pass (one GPU): full GPU test suite (new metric-shape engine test), smoke, default bench (parity, roofline_gemm)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_r02a.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02a.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r02a.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r02a.log 2>&1
timeout 600 python bench.py --kv fp8 --no-cpu-baseline --steps 32 > gpurun_out/bench_r02a_fp8.log 2>&1
tail -6 gpurun_out/pytest_r02a.log; tail -1 gpurun_out/smoke_r02a.log; tail -1 gpurun_out/bench_r02a.log; tail -1 gpurun_out/bench_r02a_fp8.log | cut -c1-600
