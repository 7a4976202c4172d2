#!/bin/bash
# round-1 final profiling pass (one GPU): launch list of a bench run + full captures of the two dominant kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
K='regex:qmatmul|paged_attn|rms_norm|rope_and|silu_mul|argmax|embedding|zero_f32|advance_meta|finish_slabs'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1100 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_launchlist.log 2>&1
python tools/agg_launches.py gpurun_out/launches_r01_final.csv > gpurun_out/launches_r01_final.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:paged_attn_decode_kernel -s 40 -c 1 -f -o /tmp/p_attn python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_attn_final.log 2>&1
ncu -i /tmp/p_attn.ncu-rep --page details > gpurun_out/attn_details_final.txt 2>/dev/null
ncu -i /tmp/p_attn.ncu-rep --page raw --csv > gpurun_out/attn_raw_final.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qmatmul_tc_kernel -c 1 -f -o /tmp/p_q4k python tools/gemm_check.py 32 28672 4096 12 0 > gpurun_out/ncu_q4k_final.log 2>&1
ncu -i /tmp/p_q4k.ncu-rep --page details > gpurun_out/q4k_details_final.txt 2>/dev/null
ncu -i /tmp/p_q4k.ncu-rep --page raw --csv > gpurun_out/q4k_raw_final.csv 2>/dev/null
ncu -i /tmp/p_q4k.ncu-rep --page source --csv > gpurun_out/q4k_source_final.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:qmatmul_tc_kernel -c 1 -f -o /tmp/p_q6k python tools/gemm_check.py 32 128256 4096 14 0 > gpurun_out/ncu_q6k_final.log 2>&1
ncu -i /tmp/p_q6k.ncu-rep --page details > gpurun_out/q6k_details_final.txt 2>/dev/null
ncu -i /tmp/p_q6k.ncu-rep --page raw --csv > gpurun_out/q6k_raw_final.csv 2>/dev/null
cat gpurun_out/launches_r01_final.txt
