#!/bin/bash
# round-2 pass V (one GPU): A/B of the attention queue / chunk knobs inside the engine step (same box)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('   ms/step %.3f  attn %.1f us' % (d['ms_per_step'], r['ms_per_launch']*1e3))"; }
for v in "1 1" "0 1" "1 2" "1 0" "0 0"; do set -- $v; echo "gptq_fp8kv adaptive=$1 queue=$2"; B200_ATTN_ADAPTIVE=$1 B200_ATTN_STATIC=$2 timeout 300 python bench.py --config gptq_fp8kv --steps 48 --no-cpu-baseline --parity-steps 0 2>/dev/null | run; done
for v in "1" "0"; do echo "q4k adaptive=$v"; B200_ATTN_ADAPTIVE=$v timeout 300 python bench.py --steps 48 --no-cpu-baseline --parity-steps 0 2>/dev/null | run; done
