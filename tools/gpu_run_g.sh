#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tp.py tests/test_fp_linear_gpu.py -m gpu -q -x > gpurun_out/pytest_g.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_g.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 4 > gpurun_out/bench_tp2_v16.log 2>&1
B200_TP_NCCL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 64 --warmup 4 > gpurun_out/bench_tp2_v16_nccl.log 2>&1
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_g.log | tail -12; tail -3 gpurun_out/bench_tp2_v16.log | cut -c1-700; tail -1 gpurun_out/bench_tp2_v16_nccl.log | cut -c1-250
