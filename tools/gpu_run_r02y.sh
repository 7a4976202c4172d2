#!/bin/bash
# round-2 pass Y (one GPU): L2 prefetch of the next GEMM's weights from the tail of each GEMM -- A/B inside the decode step
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; g=d['roofline_gemm']; print('   ms/step %.3f  value %.0f  attn %.1f us  gemm chain %.3f ms (%.3f)' % (d['ms_per_step'], d['value'], r['ms_per_launch']*1e3, g['ms_per_launch'], g['frac']))"; }
for pf in 1 0 1 0; do echo "q4k prefetch=$pf"; B200_GEMM_PREFETCH=$pf timeout 300 python bench.py --steps 64 --no-cpu-baseline --parity-steps 0 2>/dev/null | run; done
for pf in 1 0; do echo "gptq prefetch=$pf"; B200_GEMM_PREFETCH=$pf timeout 300 python bench.py --config gptq_fp8kv --steps 48 --no-cpu-baseline --parity-steps 0 2>/dev/null | run; done
timeout 600 python -m pytest tests/test_llama_gpu.py tests/test_qmatmul_gpu.py -q -m gpu -x 2>&1 | tail -2
