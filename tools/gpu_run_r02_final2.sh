#!/bin/bash
# round-2 closing pass (one GPU): what the driver runs at round end (tests, smoke, both bench arms) + refreshed ncu exports
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_r02_final2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r02_final2.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_r02_final2.log
timeout 900 python bench.py > gpurun_out/bench_r02_final2.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_final2_reference.log 2>&1
timeout 600 python bench.py --config dense_bf16 --steps 32 --no-cpu-baseline > gpurun_out/bench_r02_final2_dense.log 2>&1
timeout 600 python bench.py --config gptq_fp8kv --steps 32 --no-cpu-baseline > gpurun_out/bench_r02_final2_gptq.log 2>&1
K='regex:qmatmul|paged_attn|rms_norm|rope_and|silu_mul|argmax|embedding|zero_f32|advance_meta|finish_slabs|layer_mega|dense_gemm'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1100 --csv --log-file gpurun_out/launches_r02_gptq.csv python bench.py --config gptq_fp8kv --no-cpu-baseline --steps 2 --warmup 1 --parity-steps 0 > gpurun_out/ncu_launchlist_r02_gptq.log 2>&1
python tools/agg_launches.py gpurun_out/launches_r02_gptq.csv > gpurun_out/launches_r02_gptq.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paged_attn_decode_kernel -s 40 -c 1 -f -o /tmp/p_attn python bench.py --no-cpu-baseline --steps 2 --warmup 1 --parity-steps 0 > gpurun_out/ncu_attn_r02b.log 2>&1
ncu -i /tmp/p_attn.ncu-rep --page raw --csv > gpurun_out/attn_raw_r02b.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paged_attn_decode_kernel -s 40 -c 1 -f -o /tmp/p_attn8 python bench.py --config gptq_fp8kv --no-cpu-baseline --steps 2 --warmup 1 --parity-steps 0 > gpurun_out/ncu_attn_fp8_r02b.log 2>&1
ncu -i /tmp/p_attn8.ncu-rep --page raw --csv > gpurun_out/attn_fp8_raw_r02b.csv 2>/dev/null
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02_final2.log gpurun_out/smoke_r02_final2.log | head
for f in gpurun_out/bench_r02_final2.log gpurun_out/bench_r02_final2_reference.log gpurun_out/bench_r02_final2_dense.log gpurun_out/bench_r02_final2_gptq.log; do echo "== $f"; tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}; g=d.get('roofline_gemm') or {}; p=d.get('parity') or {}
    print('  ms', d.get('ms_per_step'), 'value', d.get('value'), 'steps', d.get('steps'), 'e2e', (d.get('e2e') or {}).get('value'), 'attn', r.get('ms_per_launch'), r.get('frac'), r.get('traffic'), 'gemm', g.get('ms_per_launch'), g.get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'parity', p.get('logits_max_err'), p.get('within_tolerance'))
    print('  config', d.get('config'))
except Exception as e: print('ERR', e)
"; done
cat gpurun_out/launches_r02_gptq.txt | head -16
