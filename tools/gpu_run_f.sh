#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fp_linear_gpu.py tests/test_tp.py -m gpu -q > gpurun_out/pytest_f.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_f.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 4 > gpurun_out/bench_tp2_v15.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_f.log | tail -12; tail -2 gpurun_out/bench_tp2_v15.log | cut -c1-300
