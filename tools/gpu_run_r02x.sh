#!/bin/bash
# round-2 pass X (2 GPUs): TP tests at world 2 on both engine paths, bench N = 2 with the teacher-forced parity (broadcast of the reference tokens)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp.py -q -m gpu > gpurun_out/pytest_r02x_tp.log 2>&1
echo "pytest(tp) exit $?" >> gpurun_out/pytest_r02x_tp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --no-cpu-baseline > gpurun_out/bench_r02x_tp2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r02x_tp2_reference.log 2>&1
grep -E "exit|passed|failed|skipped|Error|error" gpurun_out/pytest_r02x_tp.log | head
for f in gpurun_out/bench_r02x_tp2.log gpurun_out/bench_r02x_tp2_reference.log; do echo "== $f"; tail -1 $f | cut -c1-200; tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}; g=d.get('roofline_gemm') or {}
    print('  ms', d.get('ms_per_step'), 'value', d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'attn', r.get('ms_per_launch'), r.get('frac'), 'gemm', g.get('ms_per_launch'), g.get('frac'), 'parity', d.get('parity'))
except Exception as e: print('ERR', e)
"; done
