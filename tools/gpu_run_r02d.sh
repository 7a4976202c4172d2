#!/bin/bash
# round-2 pass D (one GPU): layer-kernel timeline after the slab-load fix, dense tcgen05 GEMM tests, FP8 attention profile
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B200_MEGA_TRACE=3 timeout 300 python tools/mega_trace.py 32 4096 6 > gpurun_out/mega_trace_r02d.log 2>&1
timeout 600 python -m pytest tests/test_dense_gemm_gpu.py -q -x > gpurun_out/pytest_r02d_dense.log 2>&1
echo "pytest(dense) exit $?" >> gpurun_out/pytest_r02d_dense.log
timeout 900 python -m pytest tests/test_llama_gpu.py -q -s > gpurun_out/pytest_r02d_llama.log 2>&1
echo "pytest(llama) exit $?" >> gpurun_out/pytest_r02d_llama.log
timeout 600 python bench.py --no-cpu-baseline --steps 64 > gpurun_out/bench_r02d_mega.log 2>&1
timeout 120 python tools/attn_check.py 32 4400 32 8 12 fp8 > gpurun_out/attn_fp8_r02d.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paged_attn_decode_kernel -s 2 -c 1 -f -o gpurun_out/fp8attn_r02d python tools/attn_check.py 32 4400 32 8 2 fp8 > gpurun_out/ncu_fp8attn_r02d.log 2>&1
cat gpurun_out/mega_trace_r02d.log
grep -E "exit|passed|failed|metric shapes|spread|Error|error" gpurun_out/pytest_r02d_dense.log gpurun_out/pytest_r02d_llama.log | head -30
tail -1 gpurun_out/bench_r02d_mega.log | cut -c1-300
cat gpurun_out/attn_fp8_r02d.log
