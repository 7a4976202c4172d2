"""Fused mixture-of-experts (GGUF experts) restated in numpy.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.  Follows ``FusedMoe::forward`` (GGUF flavour,
/root/reference/src/openai/models/layers/moe.rs:1429-1482): router logits -> ``topk_softmax`` (softmax over ALL experts, the k largest
probabilities, not renormalised) -> optional ``norm_topk_prob`` / ``routed_scaling_factor`` -> per (token, slot) pair:
``down_e(silu(gate_e x) * up_e x) * weight`` -> sum over the k slots.  PARITY UNPINNED for the kernels (attention-rs); the routing
convention is vLLM's topk_softmax, which attention-rs ports.  ``sort_expert_assignments`` (moe.rs:35-45) is a plain ascending sort of
the flattened expert ids; any permutation that sorts them is a valid result, so tests check sortedness + permutation, not equality.
"""
from __future__ import annotations

import numpy as np

from . import ggml_quants as G


def topk_softmax(logits: np.ndarray, k: int):
    x = np.asarray(logits, np.float64)
    p = np.exp(x - x.max(axis=-1, keepdims=True))
    p = p / p.sum(axis=-1, keepdims=True)
    # the k largest, ties -> smaller expert id (stable sort on -p)
    ids = np.argsort(-p, axis=-1, kind="stable")[:, :k]
    return np.take_along_axis(p, ids, axis=-1).astype(np.float32), ids.astype(np.uint32)


def expert_weight(stacked: np.ndarray, ggml_type: int, e: int, n: int, k: int) -> np.ndarray:
    bb, be = G.BLOCK_BYTES[ggml_type], G.BLOCK_ELEMS[ggml_type]
    per = n * (k // be) * bb
    flat = np.asarray(stacked, np.uint8).reshape(-1)
    return G.dequantize_weight(flat[e * per:(e + 1) * per], ggml_type, n, k)


def moe_gemm(x, stacked, ggml_type, E, n, k, topk_ids_flat, topk, weights_flat=None):
    """out [P, n]: row p = W[e_p] . x[row(p)] (* weight_p); x has P rows (one per pair) or P / topk (one per token)."""
    P = len(topk_ids_flat)
    x = np.asarray(x, np.float64)
    per_token = x.shape[0] != P
    out = np.zeros((P, n), np.float64)
    for e in np.unique(topk_ids_flat):
        w = expert_weight(stacked, ggml_type, int(e), n, k).astype(np.float64)
        for p in np.nonzero(topk_ids_flat == e)[0]:
            out[p] = w @ x[p // topk if per_token else p]
            if weights_flat is not None:
                out[p] *= weights_flat[p]
    return out.astype(np.float32)


def moe_gemm_fp8(x, experts_u8, scale, by, bx, topk_ids_flat, topk, weights_flat=None):
    """moe.rs:1447-1473 with block-FP8 experts: experts e4m3 bytes [E, n, k], scale f32 [E, ceil(n/by), ceil(k/bx)]."""
    from . import fp_formats as F
    P = len(topk_ids_flat)
    x = np.asarray(x, np.float64)
    E, n, k = experts_u8.shape
    per_token = x.shape[0] != P
    out = np.zeros((P, n), np.float64)
    for e in np.unique(topk_ids_flat):
        w = F.dequant_fp8_block(experts_u8[int(e)], scale[int(e)], by, bx).astype(np.float64)
        for p in np.nonzero(topk_ids_flat == e)[0]:
            out[p] = w @ x[p // topk if per_token else p]
            if weights_flat is not None:
                out[p] *= weights_flat[p]
    return out.astype(np.float32)


def fused_moe(x, gate, ge, ue, de, types, E, H, I, k, norm_topk_prob=True, routed_scaling_factor=None):
    x = np.asarray(x, np.float32)
    w, ids = topk_softmax(x.astype(np.float64) @ np.asarray(gate, np.float64).T, k)
    if norm_topk_prob:
        w = w / w.sum(axis=-1, keepdims=True)
    if routed_scaling_factor is not None:
        w = w * routed_scaling_factor
    flat = ids.reshape(-1)
    g = moe_gemm(x, ge, types[0], E, I, H, flat, k).astype(np.float64)
    u = moe_gemm(x, ue, types[1], E, I, H, flat, k).astype(np.float64)
    d_in = (g / (1.0 + np.exp(-g)) * u).astype(np.float32)
    y = moe_gemm(d_in, de, types[2], E, H, I, flat, k, w.reshape(-1))
    return y.reshape(x.shape[0], k, H).sum(axis=1), (w, ids)
