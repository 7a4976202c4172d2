#!/bin/bash
# what the driver does at round end (one GPU): full GPU test suite, smoke, reference arm, default bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
: > gpurun_out/attn_sweep.log
for c in 2 4 8; do for shape in "32 4400 32 8" "32 4400 16 4" "32 4400 4 1"; do
  B200_ATTN_CHUNK_PAGES=$c timeout 120 python tools/attn_check.py $shape 12 2>&1 | tail -1 | sed "s/^/chunk=$c $shape: /" >> gpurun_out/attn_sweep.log
done; done
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_final.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1
cat gpurun_out/attn_sweep.log; tail -4 gpurun_out/pytest_final.log; tail -1 gpurun_out/smoke_final.log; tail -1 gpurun_out/bench_reference_final.log | cut -c1-400; tail -1 gpurun_out/bench_final.log
