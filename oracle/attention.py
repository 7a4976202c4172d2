"""Paged attention (decode and chunked prefill), RoPE, RMSNorm, SwiGLU -- numpy fp32/fp64.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.  PARITY UNPINNED: the kernels
(paged_attention_v1/v2, flash varlen, FusedRope, rms_norm) are in attention-rs / candle-nn, not
in /root/reference.  Semantics follow ``NaiveAttention::forward``
(src/openai/models/mod.rs:1268-1307): softmax(q k^T * scale [-> tanh softcap]) v with GQA
``repeat_kv`` (mod.rs:1240-1247) and scale = 1/sqrt(head_dim) (layers/attention.rs:569,891),
applied to K/V gathered through the block table.
"""
from __future__ import annotations

import numpy as np

from . import cache_ops


def _gather_kv_flash(k_cache, v_cache, table, ctx_len, fp8=False):
    """caches [nb, bs, kvh, hd]; table: block ids; -> K,V [ctx_len, kvh, hd] f64."""
    bs = k_cache.shape[1]
    nblk = -(-ctx_len // bs)
    ids = np.asarray(table[:nblk], np.int64)
    k = k_cache[ids].reshape(nblk * bs, *k_cache.shape[2:])[:ctx_len]
    v = v_cache[ids].reshape(nblk * bs, *v_cache.shape[2:])[:ctx_len]
    if fp8:
        k, v = cache_ops.e4m3_to_f32(k), cache_ops.e4m3_to_f32(v)
    return k.astype(np.float64), v.astype(np.float64)


def _gather_kv_paged(k_cache, v_cache, table, ctx_len, fp8=False):
    """K [nb, kvh, hd/x, bs, x], V [nb, kvh, hd, bs] -> [ctx_len, kvh, hd]."""
    nb, kvh, hdx, bs, x = k_cache.shape
    nblk = -(-ctx_len // bs)
    ids = np.asarray(table[:nblk], np.int64)
    k = k_cache[ids].transpose(0, 3, 1, 2, 4).reshape(nblk * bs, kvh, hdx * x)[:ctx_len]
    v = v_cache[ids].transpose(0, 3, 1, 2).reshape(nblk * bs, kvh, -1)[:ctx_len]
    if fp8:
        k, v = cache_ops.e4m3_to_f32(k), cache_ops.e4m3_to_f32(v)
    return k.astype(np.float64), v.astype(np.float64)


def _attend(q, k, v, scale, softcap=None, mask=None):
    """q [Tq, H, hd], k/v [Tk, kvh, hd] f64 -> [Tq, H, hd].  mask [Tq, Tk] additive or None."""
    H, kvh = q.shape[1], k.shape[1]
    rep = H // kvh
    # repeat_kv (mod.rs:1240-1247): head h of q uses kv head h // rep
    kk = np.repeat(k, rep, axis=1)
    vv = np.repeat(v, rep, axis=1)
    s = np.einsum("qhd,khd->hqk", q, kk) * scale
    if softcap is not None:
        s = np.tanh(s / softcap) * softcap
    if mask is not None:
        s = s + mask[None]
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(axis=-1, keepdims=True)
    return np.einsum("hqk,khd->qhd", p, vv)


def paged_attention_decode(q, k_cache, v_cache, block_tables, context_lens, scale,
                           layout="flash", softcap=None, sliding_window=None, fp8=False):
    """One query token per sequence.  q [B, H, hd] -> out f32 [B, H, hd].

    ``context_lens`` INCLUDES the token being decoded (inputs.rs:393-397); its K/V must already
    be in the cache (PagedAttention.forward writes the cache first).
    sliding_window w: only the last w tokens are visible.
    """
    gather = _gather_kv_flash if layout == "flash" else _gather_kv_paged
    out = np.zeros(q.shape, np.float32)
    for b in range(q.shape[0]):
        L = int(context_lens[b])
        if L == 0:
            continue
        k, v = gather(k_cache, v_cache, block_tables[b], L, fp8)
        if sliding_window is not None and L > sliding_window:
            k, v = k[L - sliding_window:], v[L - sliding_window:]
        out[b] = _attend(q[b:b + 1].astype(np.float64), k, v, scale, softcap)[0]
    return out


def paged_attention_prefill(q, k_cache, v_cache, block_tables, cu_seqlens_q, cu_seqlens_k, scale,
                            layout="flash", softcap=None, sliding_window=None, fp8=False):
    """Varlen causal attention where all K/V (cached prefix + this chunk) is read from the paged
    cache (chunked prefill, ``use_cached_kv`` inputs.rs:133-143).  q [sum_q, H, hd];
    sequence i has q rows cu_seqlens_q[i]:cu_seqlens_q[i+1] which are the LAST q_len positions of
    its k_len = cu_seqlens_k[i+1]-cu_seqlens_k[i] context (causal, bottom-right aligned)."""
    gather = _gather_kv_flash if layout == "flash" else _gather_kv_paged
    out = np.zeros(q.shape, np.float32)
    n = len(cu_seqlens_q) - 1
    for i in range(n):
        q0, q1 = int(cu_seqlens_q[i]), int(cu_seqlens_q[i + 1])
        klen = int(cu_seqlens_k[i + 1]) - int(cu_seqlens_k[i])
        qlen = q1 - q0
        if qlen == 0:
            continue
        k, v = gather(k_cache, v_cache, block_tables[i], klen, fp8)
        qpos = np.arange(klen - qlen, klen)[:, None]
        kpos = np.arange(klen)[None, :]
        allowed = kpos <= qpos
        if sliding_window is not None:
            allowed &= kpos > qpos - sliding_window
        mask = np.where(allowed, 0.0, -np.inf)
        out[q0:q1] = _attend(q[q0:q1].astype(np.float64), k, v, scale, softcap, mask)
    return out


# --------------------------------------------------------------------------------------
# RoPE  (src/openai/models/layers/rotary_emb.rs:52-101; K15)
# --------------------------------------------------------------------------------------
def rope_tables(head_dim: int, max_pos: int, theta: float = 500000.0):
    """cos/sin f32 [max_pos, head_dim/2].  ``calculate_default_inv_freq`` (rotary_emb.rs:14-19):
    inv_freq[i] = 1f32 / (base^(i/dim) computed in f64, cast to f32); angle = pos(f32)*inv(f32)
    (rotary_emb.rs:31-36)."""
    i = np.arange(0, head_dim, 2, dtype=np.float64)
    inv = (np.float32(1.0) / np.power(np.float64(theta), i / head_dim).astype(np.float32)).astype(np.float32)
    t = (np.arange(max_pos, dtype=np.float32)[:, None] * inv[None, :]).astype(np.float32)
    return np.cos(t).astype(np.float32), np.sin(t).astype(np.float32)


def apply_rope(x, cos, sin, positions, interleaved: bool):
    """x [T, h, hd] f32; interleaved (``rope_i``, GGUF llama, quantized_llama.rs:313-318) rotates
    pairs (2i, 2i+1); NeoX (``rope``, llama.rs:222) rotates (i, i+hd/2)."""
    x = np.asarray(x, np.float32)
    c = cos[np.asarray(positions)][:, None, :]
    s = sin[np.asarray(positions)][:, None, :]
    out = np.empty_like(x)
    if interleaved:
        x0, x1 = x[..., 0::2], x[..., 1::2]
        out[..., 0::2] = x0 * c - x1 * s
        out[..., 1::2] = x0 * s + x1 * c
    else:
        h = x.shape[-1] // 2
        x0, x1 = x[..., :h], x[..., h:]
        out[..., :h] = x0 * c - x1 * s
        out[..., h:] = x0 * s + x1 * c
    return out


def rms_norm(x, w, eps):
    """candle_nn::ops::rms_norm (qrmsnorm.rs:28-31): x / sqrt(mean(x^2)+eps) * w, f32."""
    x = np.asarray(x, np.float32)
    ms = (x.astype(np.float64) ** 2).mean(axis=-1, keepdims=True)
    return (x / np.sqrt(ms + eps) * w).astype(np.float32)


def silu_mul(gate, up):
    """silu(w1 x) * (w3 x)  (quantized_llama.rs:32-37)."""
    g = np.asarray(gate, np.float64)
    return (g / (1.0 + np.exp(-g)) * np.asarray(up, np.float64)).astype(np.float32)
