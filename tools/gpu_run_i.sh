#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_tp.py -m gpu -q -x > gpurun_out/pytest_i.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_i.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 48 --warmup 4 > gpurun_out/bench_tp4_v17.log 2>&1
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_i.log | tail -8; tail -2 gpurun_out/bench_tp4_v17.log | cut -c1-600
