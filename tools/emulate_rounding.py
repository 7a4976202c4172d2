#!/usr/bin/env python
"""CPU-only: where does the logits error of the decode path come from at the METRIC's shapes?

Runs the numpy oracle (oracle/llama.py) on a 2-layer model with Llama-3-8B layer shapes (hidden 4096, ffn 14336, 32 q / 8 kv heads,
Q6_K lm_head with 16 384 rows, ctx ~ 4 k) three ways and reports max|err| / max|logit| and rel-Frobenius against the exact (fp64
dequant-matmul) result:
  * fp16 operands: activations AND dequantised weights rounded once to fp16, exact accumulation -- the numerical contract of this
    repo's QMatMul with every other source of error removed: the floor of the design;
  * q8k: the reference's own CPU/GGML semantics (activations quantised to Q8_K, integer dot) -- the reference's noise floor.
Used to justify the tolerance of tests/test_llama_gpu.py::test_decode_at_metric_shapes_matches_oracle.  ~3 minutes on 8 cores.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import candle_vllm_b200 as pkg  # noqa: E402
from candle_vllm_b200 import synthetic  # noqa: E402
from oracle import ggml_quants as G, llama as OL  # noqa: E402
from tests.gpu_util import weights_to_oracle  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    cfg = pkg.LlamaConfig(hidden=4096, num_layers=2, num_heads=32, num_kv_heads=8, head_dim=128, ffn=14336, vocab=16384,
                          max_pos=5248, block_size=64, max_num_seqs=32, max_blocks_per_seq=80)
    ow = weights_to_oracle(synthetic.make_weights(cfg, "cpu", seed=0))
    nb = B * 80 + 8
    rng = np.random.default_rng(1)
    kc = [OL.bf16_round(rng.standard_normal((nb, 64, 8, 128)).astype(np.float32)) for _ in range(2)]
    vc = [OL.bf16_round(rng.standard_normal((nb, 64, 8, 128)).astype(np.float32)) for _ in range(2)]
    lens = [int(x) for x in rng.integers(3900, 4300, B)]
    tables = synthetic.random_block_tables(B, 80, nb, seed=2)
    prep = pkg.prepare_decode(lens, [int(t) for t in rng.integers(0, cfg.vocab, B)], tables, 64)
    ocfg = dict(hidden=cfg.hidden, heads=cfg.num_heads, kv_heads=cfg.num_kv_heads, head_dim=cfg.head_dim, rms_eps=cfg.rms_eps,
                max_pos=cfg.max_pos, rope_theta=cfg.rope_theta)
    meta = dict(slot_mapping=prep["slot_mapping"], block_tables=prep["block_tables"], context_lens=prep["context_lens"])

    def run(mode="dequant"):
        return OL.forward(ocfg, ow, prep["tokens"].astype(np.int64), prep["positions"], [k.copy() for k in kc], [v.copy() for v in vc], meta, mode=mode)

    t0 = time.time()
    ref = run()
    print(f"exact oracle: {time.time() - t0:.1f} s, max|logit| = {np.abs(ref).max():.3f}")
    f16 = lambda a: np.asarray(a, np.float32).astype(np.float16).astype(np.float32)
    orig = OL.qmm

    def qmm_f16(x, wt, mode):
        wb, t, n, k = wt
        return (f16(x).astype(np.float64) @ f16(G.dequantize_weight(wb, t, n, k)).T.astype(np.float64)).astype(np.float32)

    OL.qmm = qmm_f16
    out = run()
    OL.qmm = orig
    print(f"fp16 operands (this repo's QMatMul contract): max err / max = {np.abs(out - ref).max() / np.abs(ref).max():.2e}, "
          f"rel-Fro = {np.linalg.norm(out - ref) / np.linalg.norm(ref):.2e}")
    t0 = time.time()
    out = run("q8k")
    print(f"q8k (reference CPU/GGML semantics, {time.time() - t0:.0f} s): max err / max = {np.abs(out - ref).max() / np.abs(ref).max():.2e}, "
          f"rel-Fro = {np.linalg.norm(out - ref) / np.linalg.norm(ref):.2e}")


if __name__ == "__main__":
    main()
