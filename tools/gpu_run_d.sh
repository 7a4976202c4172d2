#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_qmatmul_gpu.py tests/test_llama_gpu.py tests/test_tp.py -m gpu -x -q --durations=5 > gpurun_out/pytest_d.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_d.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v13.log 2>&1
: > gpurun_out/gemm_v13.log
for shp in "28672 4096 12" "4096 14336 12" "6144 4096 12" "4096 4096 12" "128256 4096 14"; do
  timeout 200 python tools/gemm_check.py 32 $shp 12 >> gpurun_out/gemm_v13.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paged_attn_decode_kernel -s 3 -c 1 -f -o /tmp/prof_attn python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_attn_v13.log 2>&1
ncu -i /tmp/prof_attn.ncu-rep --page details > gpurun_out/attn_details_v13.txt 2>/dev/null
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > gpurun_out/attn_raw_v13.csv 2>/dev/null
tail -8 gpurun_out/pytest_d.log; tail -1 gpurun_out/bench_v13.log; grep "graph of" gpurun_out/gemm_v13.log
