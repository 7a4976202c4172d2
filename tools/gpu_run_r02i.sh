#!/bin/bash
# round-2 pass I (8 GPUs): TP parity at world 4 / 8, strong-scaling bench at N = 8 and 4
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp.py -q -m gpu -k "per_gemm" > gpurun_out/pytest_r02i_tp.log 2>&1
echo "pytest(tp) exit $?" >> gpurun_out/pytest_r02i_tp.log
runN() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 64 --no-cpu-baseline "$@"; }
runN 8 > gpurun_out/bench_r02i_tp8.log 2>&1
runN 4 > gpurun_out/bench_r02i_tp4.log 2>&1
B200_MEGA=1 runN 8 --parity-steps 0 > gpurun_out/bench_r02i_tp8_mega.log 2>&1
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02i_tp.log | head
for f in gpurun_out/bench_r02i_tp8.log gpurun_out/bench_r02i_tp4.log gpurun_out/bench_r02i_tp8_mega.log; do echo "== $f"; tail -1 $f | cut -c1-330; tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('  attn', d['roofline']['ms_per_launch'], d['roofline']['frac'], 'gemm', d['roofline_gemm']['ms_per_launch'], 'parity', d.get('parity'))
except Exception as e: print('ERR', e)
"; done
