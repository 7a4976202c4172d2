#!/bin/bash
# round-2 pass F (2 GPUs): tensor-parallel parity on the layer kernel, TP=2 bench both paths; N=1 knobs; new engine kinds; fp8 attention timing
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp.py -q -m gpu -k "world" > gpurun_out/pytest_r02f_tp.log 2>&1
echo "pytest(tp) exit $?" >> gpurun_out/pytest_r02f_tp.log
timeout 600 python -m pytest tests/test_llama_gpu.py -q -k "other_linear or fp8" > gpurun_out/pytest_r02f_llama.log 2>&1
echo "pytest(llama kinds) exit $?" >> gpurun_out/pytest_r02f_llama.log
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --no-cpu-baseline "$@"; }
run2 > gpurun_out/bench_r02f_tp2_mega.log 2>&1
B200_MEGA=0 run2 > gpurun_out/bench_r02f_tp2_legacy.log 2>&1
for v in "B200_MEGA_PDL=0" "B200_MEGA_TRIGGER=0" "B200_MEGA_PDL=0 B200_MEGA_TRIGGER=0"; do
  env $v timeout 300 python bench.py --no-cpu-baseline --steps 64 --parity-steps 0 --layers 8 2>&1 | tail -1 | cut -c1-160 | sed "s/^/$v: /" >> gpurun_out/mega_knobs_r02f.log
done
timeout 120 python tools/attn_check.py 32 4400 32 8 12 fp8 > gpurun_out/attn_fp8_r02f.log 2>&1
timeout 600 python bench.py --config gptq_fp8kv --no-cpu-baseline --steps 32 > gpurun_out/bench_r02f_gptq.log 2>&1
timeout 600 python bench.py --config dense_bf16 --no-cpu-baseline --steps 32 > gpurun_out/bench_r02f_dense.log 2>&1
grep -E "exit|passed|failed|Error|error" gpurun_out/pytest_r02f_tp.log gpurun_out/pytest_r02f_llama.log | head
for f in gpurun_out/bench_r02f_tp2_mega.log gpurun_out/bench_r02f_tp2_legacy.log gpurun_out/bench_r02f_gptq.log gpurun_out/bench_r02f_dense.log; do echo "== $f"; tail -1 $f | cut -c1-1200; done
cat gpurun_out/mega_knobs_r02f.log gpurun_out/attn_fp8_r02f.log
