#!/bin/bash
# round-2 pass AC (one GPU): ncu capture of the int4 GEMM at the gate|up shape (why is it 1.5x the Q4_K kernel per byte?)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qmatmul_tc_kernel -s 3 -c 1 -f -o /tmp/p_m4 python tools/marlin_gemm_check.py 28672 4096 2 > gpurun_out/ncu_m4_r02.log 2>&1
ncu -i /tmp/p_m4.ncu-rep --page raw --csv > gpurun_out/m4_raw_r02.csv 2>/dev/null
ncu -i /tmp/p_m4.ncu-rep --page source --csv --print-source sass > gpurun_out/m4_source_r02.csv 2>/dev/null
ncu -i /tmp/p_m4.ncu-rep --page details > gpurun_out/m4_details_r02.txt 2>/dev/null
tail -2 gpurun_out/ncu_m4_r02.log
grep -E "Duration|DRAM Throughput|Issue Slots Busy|No Eligible|bank conflict|L2 Hit|Registers Per" gpurun_out/m4_details_r02.txt | head
