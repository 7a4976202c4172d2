"""One rank's shard of a TP=t Llama-3-8B decode step on ONE GPU (no collectives): what the per-rank kernels cost at the
shard shapes.  python tools/shard_probe.py <t> [steps]   (run under ncu for a launch list)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import candle_vllm_b200 as pkg
from candle_vllm_b200 import synthetic

t = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B, bs, ctx = 32, 64, 4100
dev = torch.device("cuda", 0)
blocks_per_seq = -(-(ctx + 2 * steps + 80) // bs)
cfg = pkg.LlamaConfig(num_heads=32 // t, num_kv_heads=max(1, 8 // t), ffn=14336 // t, vocab=128256 // t, max_num_seqs=B,
                      max_blocks_per_seq=blocks_per_seq, max_pos=ctx + 2 * steps + 128, block_size=bs)
w = synthetic.make_weights(cfg, dev, seed=0)
num_blocks = B * blocks_per_seq + 16
eng = pkg.CacheEngine(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim, pkg.CacheConfig(bs, num_blocks), device=dev)
synthetic.fill_kv_cache(eng.gpu_cache, seed=1)
tables = synthetic.random_block_tables(B, blocks_per_seq, num_blocks, seed=2)
model = pkg.GGUFLLaMa(cfg, w, eng.gpu_cache)
prep = pkg.prepare_decode([ctx + 1] * B, [1] * B, tables, bs)
model.decode(prep)
for _ in range(4): model.decode_resident(B, advance=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(model.stream):
    e0.record(model.stream)
    for _ in range(steps): model.decode_resident(B, advance=True)
    e1.record(model.stream)
torch.cuda.synchronize()
print(f"shard of tp{t}: {e0.elapsed_time(e1) / steps:.3f} ms/step (no collectives), heads {cfg.num_heads}, kv {cfg.num_kv_heads}, ffn {cfg.ffn}, vocab {cfg.vocab}")
