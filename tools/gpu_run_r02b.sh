#!/bin/bash
# round-2 pass B (one GPU): persistent layer kernel -- engine tests on both paths, then benches
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
echo "== mega on" > gpurun_out/pytest_r02b.log
timeout 900 python -m pytest tests/test_llama_gpu.py -q -x -s --durations=5 >> gpurun_out/pytest_r02b.log 2>&1
echo "pytest(mega) exit $?" >> gpurun_out/pytest_r02b.log
echo "== mega off" >> gpurun_out/pytest_r02b.log
B200_MEGA=0 timeout 900 python -m pytest tests/test_llama_gpu.py -q -x -s >> gpurun_out/pytest_r02b.log 2>&1
echo "pytest(legacy) exit $?" >> gpurun_out/pytest_r02b.log
timeout 300 python -m pytest tests/test_cache_ops_gpu.py tests/test_marlin.py -q -x -m gpu >> gpurun_out/pytest_r02b.log 2>&1
echo "pytest(cache,marlin) exit $?" >> gpurun_out/pytest_r02b.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r02b_mega.log 2>&1
B200_MEGA=0 timeout 600 python bench.py --no-cpu-baseline --steps 32 > gpurun_out/bench_r02b_legacy.log 2>&1
grep -E "exit|passed|failed|metric shapes|spread|Error|error" gpurun_out/pytest_r02b.log | head -40
tail -1 gpurun_out/bench_r02b_mega.log | cut -c1-1800
tail -1 gpurun_out/bench_r02b_legacy.log | cut -c1-400
