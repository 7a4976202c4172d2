#!/bin/bash
# round-2 pass E (one GPU): per-kernel launch lists (mega vs legacy), tensor-core prefill attention + fp prefill tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_fp_linear_gpu.py tests/test_dense_gemm_gpu.py -q -m gpu > gpurun_out/pytest_r02e.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_r02e.log
K='regex:layer_mega|qmatmul|paged_attn|rms_norm|rope_and|silu_mul|argmax|embedding|zero_f32|advance_meta'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 200 -c 400 --csv --log-file gpurun_out/launches_r02e_mega.csv python bench.py --no-cpu-baseline --steps 3 --warmup 3 --parity-steps 0 --layers 8 > gpurun_out/ncu_ll_mega.log 2>&1
B200_MEGA=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 400 -c 600 --csv --log-file gpurun_out/launches_r02e_legacy.csv python bench.py --no-cpu-baseline --steps 3 --warmup 3 --parity-steps 0 --layers 8 > gpurun_out/ncu_ll_legacy.log 2>&1
python tools/agg_launches.py gpurun_out/launches_r02e_mega.csv > gpurun_out/launches_r02e_mega.txt 2>&1
python tools/agg_launches.py gpurun_out/launches_r02e_legacy.csv > gpurun_out/launches_r02e_legacy.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline --steps 64 --parity-steps 0 --layers 8 > gpurun_out/bench_r02e_mega8.log 2>&1
B200_MEGA=0 timeout 300 python bench.py --no-cpu-baseline --steps 64 --parity-steps 0 --layers 8 > gpurun_out/bench_r02e_legacy8.log 2>&1
B200_PDL=0 timeout 300 python bench.py --no-cpu-baseline --steps 64 --parity-steps 0 --layers 8 > gpurun_out/bench_r02e_mega8_nopdl.log 2>&1
tail -3 gpurun_out/pytest_r02e.log
cat gpurun_out/launches_r02e_mega.txt gpurun_out/launches_r02e_legacy.txt
for f in gpurun_out/bench_r02e_mega8.log gpurun_out/bench_r02e_legacy8.log gpurun_out/bench_r02e_mega8_nopdl.log; do tail -1 $f | cut -c1-200; done
