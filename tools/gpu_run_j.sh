#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_attention_gpu.py tests/test_llama_gpu.py tests/test_tp.py -m gpu -q -x > gpurun_out/pytest_j.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_j.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 64 --warmup 4 > gpurun_out/bench_tp2_v18.log 2>&1
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_j.log | tail -8; tail -1 gpurun_out/bench_tp2_v18.log | cut -c1-260
