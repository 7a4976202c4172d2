#!/bin/bash
# round-2 pass H (one GPU): MoE + MLA parity, split-phase deterministic engine path (B200_MEGA=2), CSR, benches
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_moe_gpu.py tests/test_mla_gpu.py -q > gpurun_out/pytest_r02h_moe_mla.log 2>&1
echo "pytest(moe,mla) exit $?" >> gpurun_out/pytest_r02h_moe_mla.log
timeout 900 python -m pytest tests/test_llama_gpu.py -q -s -k "split_phases or other_linear" > gpurun_out/pytest_r02h_llama.log 2>&1
echo "pytest(llama split) exit $?" >> gpurun_out/pytest_r02h_llama.log
B200_MEGA=2 timeout 600 python bench.py --no-cpu-baseline --steps 64 > gpurun_out/bench_r02h_split.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --steps 64 > gpurun_out/bench_r02h_legacy.log 2>&1
grep -E "exit|passed|failed|Error|error|rel_fro|assert" gpurun_out/pytest_r02h_moe_mla.log gpurun_out/pytest_r02h_llama.log | head -40
for f in gpurun_out/bench_r02h_split.log gpurun_out/bench_r02h_legacy.log; do echo "== $f"; tail -1 $f | cut -c1-420; done
