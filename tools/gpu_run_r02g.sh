#!/bin/bash
# round-2 pass G (2 GPUs): MoE, engine kinds + both engine paths, TP parity (both paths), TP=2 bench with the concurrent-poll exchange
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_moe_gpu.py -q > gpurun_out/pytest_r02g_moe.log 2>&1
echo "pytest(moe) exit $?" >> gpurun_out/pytest_r02g_moe.log
timeout 900 python -m pytest tests/test_llama_gpu.py -q -s > gpurun_out/pytest_r02g_llama.log 2>&1
echo "pytest(llama) exit $?" >> gpurun_out/pytest_r02g_llama.log
timeout 900 python -m pytest tests/test_tp.py -q -m gpu > gpurun_out/pytest_r02g_tp.log 2>&1
echo "pytest(tp) exit $?" >> gpurun_out/pytest_r02g_tp.log
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --no-cpu-baseline "$@"; }
run2 > gpurun_out/bench_r02g_tp2.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --steps 64 > gpurun_out/bench_r02g_n1.log 2>&1
grep -E "exit|passed|failed|Error|error|step [0-9]:" gpurun_out/pytest_r02g_moe.log gpurun_out/pytest_r02g_llama.log gpurun_out/pytest_r02g_tp.log | head -40
for f in gpurun_out/bench_r02g_tp2.log gpurun_out/bench_r02g_n1.log; do echo "== $f"; tail -1 $f | cut -c1-400; done
