#!/bin/bash
# scratch driver for one gpurun call: slab-mode validation + Q6_K investigation
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qmatmul_gpu.py tests/test_llama_gpu.py tests/test_tp.py -m gpu -x -q --durations=8 > gpurun_out/pytest_slabs.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_slabs.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v10.log 2>&1
: > gpurun_out/gemm_v10.log
for shp in "28672 4096 12" "4096 14336 12" "6144 4096 12" "4096 4096 12"; do
  timeout 120 python tools/gemm_check.py 32 $shp 24 >> gpurun_out/gemm_v10.log 2>&1
  B200_SLABS=1 timeout 120 python tools/gemm_check.py 32 $shp 24 >> gpurun_out/gemm_v10.log 2>&1
done
for dbg in 0 1 2 3; do B200_GEMM_DEBUG=$dbg timeout 200 python tools/gemm_check.py 32 128256 4096 14 6 >> gpurun_out/gemm_v10.log 2>&1; done
B200_SLABS=1 timeout 200 python tools/gemm_check.py 32 128256 4096 14 6 >> gpurun_out/gemm_v10.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qmatmul_tc_kernel -c 1 -f -o /tmp/prof_q6k python tools/gemm_check.py 32 128256 4096 14 0 > gpurun_out/ncu_q6k.log 2>&1
ncu -i /tmp/prof_q6k.ncu-rep --page raw --csv > gpurun_out/q6k_raw.csv 2>/dev/null
ncu -i /tmp/prof_q6k.ncu-rep --page source --csv > gpurun_out/q6k_source.csv 2>/dev/null
ncu -i /tmp/prof_q6k.ncu-rep --page details > gpurun_out/q6k_details.txt 2>/dev/null
ls -la /tmp/prof_q6k.ncu-rep >> gpurun_out/ncu_q6k.log
tail -3 gpurun_out/pytest_slabs.log; cat gpurun_out/bench_v10.log | tail -2; cat gpurun_out/gemm_v10.log
