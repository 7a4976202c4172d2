#!/bin/bash
# round-2 pass Z (8 GPUs): bench at N = 8 with the final kernels and the teacher-forced parity
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 64 --no-cpu-baseline > gpurun_out/bench_r02z_tp8.log 2>&1
tail -1 gpurun_out/bench_r02z_tp8.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}; g=d.get('roofline_gemm') or {}
    print('  ms', d.get('ms_per_step'), 'value', d.get('value'), 'e2e', (d.get('e2e') or {}).get('value'), 'attn', r.get('ms_per_launch'), r.get('frac'), 'gemm', g.get('ms_per_launch'), g.get('frac'), 'parity', d.get('parity'))
except Exception as e: print('ERR', e)
"
tail -3 gpurun_out/bench_r02z_tp8.log | cut -c1-300
