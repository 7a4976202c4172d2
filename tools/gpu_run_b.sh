#!/bin/bash
# scratch driver for one gpurun call: validate the Q6_K fix + legacy engine, launch-gap probe, fresh Q4_K ncu capture
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_qmatmul_gpu.py tests/test_llama_gpu.py tests/test_tp.py -m gpu -x -q > gpurun_out/pytest_b.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_b.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v11.log 2>&1
./tools/probes/launch_gap_probe > gpurun_out/launch_gap.log 2>&1
: > gpurun_out/gemm_v11.log
for dbg in 0 2; do B200_GEMM_DEBUG=$dbg timeout 200 python tools/gemm_check.py 32 128256 4096 14 6 >> gpurun_out/gemm_v11.log 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qmatmul_tc_kernel -c 1 -f -o /tmp/prof_q4k python tools/gemm_check.py 32 28672 4096 12 0 > gpurun_out/ncu_q4k_v11.log 2>&1
ncu -i /tmp/prof_q4k.ncu-rep --page source --csv > gpurun_out/q4k_source_v11.csv 2>/dev/null
ncu -i /tmp/prof_q4k.ncu-rep --page details > gpurun_out/q4k_details_v11.txt 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qmatmul_tc_kernel -c 1 -f -o /tmp/prof_q6k python tools/gemm_check.py 32 128256 4096 14 0 > gpurun_out/ncu_q6k_v11.log 2>&1
ncu -i /tmp/prof_q6k.ncu-rep --page source --csv > gpurun_out/q6k_source_v11.csv 2>/dev/null
ncu -i /tmp/prof_q6k.ncu-rep --page details > gpurun_out/q6k_details_v11.txt 2>/dev/null
tail -3 gpurun_out/pytest_b.log; tail -1 gpurun_out/bench_v11.log | cut -c1-400; cat gpurun_out/launch_gap.log gpurun_out/gemm_v11.log
