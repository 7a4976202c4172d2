#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qmatmul_gpu.py tests/test_llama_gpu.py -m gpu -x -q > gpurun_out/pytest_c.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_c.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v12.log 2>&1
: > gpurun_out/gemm_v12.log
for shp in "28672 4096 12" "4096 14336 12" "6144 4096 12" "4096 4096 12"; do
  B200_TRACE=1 timeout 120 python tools/gemm_check.py 32 $shp 24 >> gpurun_out/gemm_v12.log 2>&1
done
tail -3 gpurun_out/pytest_c.log; tail -1 gpurun_out/bench_v12.log | cut -c1-330; cat gpurun_out/gemm_v12.log
