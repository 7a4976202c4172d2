"""GGUF-LLaMA decode/prefill step restated in numpy.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.  Follows the op order and dtype
flow of ``GGUFLLaMa::forward_inner`` (src/openai/models/quantized_llama.rs:424-506),
``QuantizedAttention::forward`` (src/openai/models/layers/attention.rs:910-1011) and
``Mlp::forward`` (quantized_llama.rs:32-44):

    x(f32) = tok_embeddings[tokens]
    per layer:  h = rms_norm(x); q,k,v = QMatMul(h) (f32); rope_i(q,k) (f32, interleaved);
                q,k,v -> bf16; cache write; paged attention -> bf16; -> f32; wo; x += .
                h = rms_norm(x); x += w2(silu(w1 h) * w3 h)
    logits(f32) = output(rms_norm(x))

Weights: dict with keys tok_embeddings f32[V,H], norm f32[H], output (bytes, type), and
layers[i] = dict(attn_norm, ffn_norm f32[H]; wq, wk, wv, wo, w1, w2, w3 = (u8 bytes, ggml_type,
N, K)).  ``mode`` selects QMatMul semantics: "dequant" (fp64 accumulate; the target) or "q8k"
(reference CPU path: Q8_K-quantised activations, integer dot [UPSTREAM]).
"""
from __future__ import annotations

import numpy as np

from . import attention as A
from . import cache_ops as C
from . import ggml_quants as G


def bf16_round(x: np.ndarray) -> np.ndarray:
    """f32 -> nearest-even bf16 -> f32 (the ``to_dtype(self.dtype)`` casts, attention.rs:971-975)."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16).reshape(np.shape(x))


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.asarray(b, np.uint16).astype(np.uint32) << 16).view(np.float32)


def qmm(x, w, mode):
    """w = (ggml bytes, type, n, k) -- the GGUF models -- or ("dense", W [n, k]) (Linear, linear.rs:124-172) or
    ("gptq", qweight u32 [K/8, N], scales [K/g, N], group) (GPTQ symmetric int4, oracle/gptq.py)."""
    if isinstance(w[0], str):
        if w[0] == "dense":
            return (np.asarray(x, np.float64) @ np.asarray(w[1], np.float64).T).astype(np.float32)
        if w[0] == "gptq":
            from . import gptq as OG
            return OG.gptq_matmul(np.asarray(x, np.float32), w[1], w[2], w[3]).astype(np.float32)
        raise ValueError(w[0])
    wbytes, t, n, k = w
    if mode == "q8k":
        return G.qmatmul_q8k(x, wbytes, t, n, k)
    return G.qmatmul_dequant(x, wbytes, t, n, k)


def forward(cfg, weights, tokens, positions, k_caches, v_caches, meta, mode="dequant",
            is_prefill=False, fp8_kv=False):
    """One forward.  k_caches/v_caches: per-layer f32 arrays holding bf16-representable values
    (or u8 e4m3 bits when ``fp8_kv``) in flash layout [nb, bs, kvh, hd]; updated in place.
    meta: dict(slot_mapping, block_tables, context_lens) for decode, plus cu_seqlens_q/k for
    prefill.  Returns logits f32 [T or n_seqs, V]."""
    H, nh, nkv, hd = cfg["hidden"], cfg["heads"], cfg["kv_heads"], cfg["head_dim"]
    eps = cfg["rms_eps"]
    cos, sin = A.rope_tables(hd, cfg["max_pos"], cfg["rope_theta"])
    scale = 1.0 / np.sqrt(np.float32(hd))
    x = weights["tok_embeddings"][np.asarray(tokens)].astype(np.float32)
    T = x.shape[0]
    for li, lw in enumerate(weights["layers"]):
        h = A.rms_norm(x, lw["attn_norm"], eps)
        q = qmm(h, lw["wq"], mode).reshape(T, nh, hd)
        k = qmm(h, lw["wk"], mode).reshape(T, nkv, hd)
        v = qmm(h, lw["wv"], mode).reshape(T, nkv, hd)
        il = not cfg.get("rope_neox", False)            # GGUF llama: rope_i (quantized_llama.rs:313-318); safetensors llama: NeoX (llama.rs:222)
        q = A.apply_rope(q, cos, sin, positions, interleaved=il)
        k = A.apply_rope(k, cos, sin, positions, interleaved=il)
        q, k, v = bf16_round(q), bf16_round(k), bf16_round(v)
        C.reshape_and_cache_flash(k, v, k_caches[li], v_caches[li], meta["slot_mapping"], fp8=fp8_kv)
        if is_prefill:
            y = A.paged_attention_prefill(q, k_caches[li], v_caches[li], meta["block_tables"],
                                          meta["cu_seqlens_q"], meta["cu_seqlens_k"], scale,
                                          fp8=fp8_kv)
        else:
            y = A.paged_attention_decode(q, k_caches[li], v_caches[li], meta["block_tables"],
                                         meta["context_lens"], scale, fp8=fp8_kv)
        y = bf16_round(y).reshape(T, nh * hd)
        x = x + qmm(y, lw["wo"], mode)
        h = A.rms_norm(x, lw["ffn_norm"], eps)
        g = A.silu_mul(qmm(h, lw["w1"], mode), qmm(h, lw["w3"], mode))
        x = x + qmm(g, lw["w2"], mode)
    if is_prefill:
        last = np.asarray(meta["cu_seqlens_q"][1:], np.int64) - 1     # quantized_llama.rs:495-499
        x = x[last]
    x = A.rms_norm(x, weights["norm"], eps)
    return qmm(x, weights["output"], mode)
