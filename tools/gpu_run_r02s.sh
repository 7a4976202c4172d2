#!/bin/bash
# round-2 pass S (one GPU): FP8 attention vs context length (why ctx 4664 is slower than 4400)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
L=gpurun_out/attn_r02s.log
: > $L
for ctx in 4096 4480 4608 4664 4736 5120; do for cp in 0 4; do echo "== ctx $ctx chunk pages $cp" >> $L; B200_ATTN_CHUNK_PAGES=$cp timeout 120 python tools/attn_check.py 32 $ctx 32 8 12 fp8 2>&1 | tail -1 >> $L; done; done
for ctx in 4608 4664; do echo "== bf16 ctx $ctx" >> $L; timeout 120 python tools/attn_check.py 32 $ctx 32 8 12 2>&1 | tail -1 >> $L; done
cat $L
