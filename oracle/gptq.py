"""GPTQ / Marlin int4 weight-only GEMM restated in numpy.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.  Formats and host-side transforms follow the
reference: qweight u32 [K/8, N], 8 nibbles along K per word, LSB first
(/root/reference/src/openai/models/linear.rs:225-242), scales [K/g, N]; symmetric 4-bit only, group
size 64 / 128 / -1, no act-order (linear.rs:319-325).  ``marlin_permute_scales`` restates
linear.rs:341-379; its permutation tables are pinned against the reference's own Python
(examples/convert_awq_marlin.py:8-17) through tests/golden/marlin_perms.json.  The Marlin kernel itself
(attention-rs) is not in the tree: w = (q - 8) * scale is the published GPTQ-sym/Marlin convention.
PARITY UNPINNED for the GEMM values (no reference fixtures).
"""
from __future__ import annotations

import numpy as np


def get_scale_perms():
    """linear.rs:341-352 / examples/convert_awq_marlin.py:8-17."""
    scale_perm = []
    for i in range(8):
        scale_perm.extend([i + 8 * j for j in range(8)])
    scale_perm_single = []
    for i in range(4):
        scale_perm_single.extend([2 * i + j for j in [0, 1, 8, 9, 16, 17, 24, 25]])
    return scale_perm, scale_perm_single


def pack_gptq(q: np.ndarray) -> np.ndarray:
    """q u8 [K, N] in 0..15 -> qweight u32 [K/8, N]; nibble i of word (kp, n) = q[8 kp + i, n]."""
    K, N = q.shape
    q = q.astype(np.uint32).reshape(K // 8, 8, N)
    out = np.zeros((K // 8, N), np.uint32)
    for i in range(8):
        out |= q[:, i, :] << np.uint32(4 * i)
    return out


def unpack_gptq(qweight: np.ndarray) -> np.ndarray:
    Kp, N = qweight.shape
    q = np.empty((Kp, 8, N), np.uint8)
    for i in range(8):
        q[:, i, :] = (qweight >> np.uint32(4 * i)) & 0xF
    return q.reshape(Kp * 8, N)


def marlin_permute_scales(s: np.ndarray, size_k: int, size_n: int, group_size: int) -> np.ndarray:
    """linear.rs:354-379."""
    scale_perm, scale_perm_single = get_scale_perms()
    if group_size != -1 and group_size < size_k:
        s = s.reshape(-1, len(scale_perm))[:, scale_perm]
    else:
        s = s.reshape(-1, len(scale_perm_single))[:, scale_perm_single]
    return np.ascontiguousarray(s.reshape(-1, size_n))


def dequant_gptq(qweight: np.ndarray, scales: np.ndarray, group_size: int) -> np.ndarray:
    """-> W f64 [N, K] with W[n, k] = (q[k, n] - 8) * scales[k // g, n]  (scales in ORIGINAL order)."""
    q = unpack_gptq(qweight).astype(np.float64)
    K, N = q.shape
    g = K if group_size == -1 else group_size
    s = np.repeat(np.asarray(scales, np.float64), g, axis=0)[:K]
    return ((q - 8.0) * s).T


def gptq_matmul(x: np.ndarray, qweight: np.ndarray, scales: np.ndarray, group_size: int) -> np.ndarray:
    return (np.asarray(x, np.float64) @ dequant_gptq(qweight, scales, group_size).T).astype(np.float32)


# ---- AWQ (zero-point int4) and conventional GPTQ (act-order / asymmetric) ------------------------------------------------
AWQ_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]        # nibble i of an AWQ word holds column 8 j + AWQ_ORDER[i] (AutoAWQ pack order)


def pack_cols(q: np.ndarray) -> np.ndarray:
    """u8 [R, N] -> u32 [R, N/8], nibble i of word j = column 8 j + i (examples/convert_awq_marlin.py:19-42)."""
    R, N = q.shape
    q = q.astype(np.uint32).reshape(R, N // 8, 8)
    out = np.zeros((R, N // 8), np.uint32)
    for i in range(8):
        out |= q[:, :, i] << np.uint32(4 * i)
    return out


def unpack_cols(packed: np.ndarray) -> np.ndarray:
    """examples/convert_awq_marlin.py:44-73"""
    R, Np = packed.shape
    q = np.empty((R, Np, 8), np.uint8)
    for i in range(8):
        q[:, :, i] = (packed >> np.uint32(4 * i)) & 0xF
    return q.reshape(R, Np * 8)


def pack_awq(q: np.ndarray) -> np.ndarray:
    """u8 [R, N] natural columns -> AWQ words [R, N/8]"""
    R, N = q.shape
    return pack_cols(q.reshape(R, N // 8, 8)[:, :, AWQ_ORDER].reshape(R, N))


def unpack_awq(packed: np.ndarray) -> np.ndarray:
    q = unpack_cols(packed)
    R, N = q.shape
    return q.reshape(R, N // 8, 8)[:, :, np.argsort(AWQ_ORDER)].reshape(R, N)


def marlin_zero_points(zp: np.ndarray) -> np.ndarray:
    """natural zero points u8 [G, N] -> the packed layout marlin_awq_4bit_* takes (examples/convert_awq_marlin.py:75-96):
    scale_perm within 64-column blocks, [0,2,4,6,1,3,5,7] interleave within 8, packed along columns."""
    G, N = zp.shape
    scale_perm, _ = get_scale_perms()
    z = zp.reshape(-1, 64)[:, scale_perm]
    z = z.reshape(-1, 8)[:, AWQ_ORDER].reshape(G, N)
    return pack_cols(z)


def awq_to_marlin_zero_points(qzeros_awq: np.ndarray) -> np.ndarray:
    """examples/convert_awq_marlin.py:99-113: what the reference's offline converter writes into ``qzeros``"""
    return marlin_zero_points(unpack_awq(qzeros_awq))


def dequant_awq(qweight_awq: np.ndarray, qzeros_awq: np.ndarray, scales: np.ndarray, group_size: int) -> np.ndarray:
    """-> W f64 [N, K] = (q[k, n] - z[k // g, n]) * s[k // g, n]; qweight [K, N/8], qzeros [K/g, N/8] AWQ words"""
    q = unpack_awq(qweight_awq).astype(np.float64)
    K, N = q.shape
    g = K if group_size == -1 else group_size
    z = np.repeat(unpack_awq(qzeros_awq).astype(np.float64), g, axis=0)[:K]
    s = np.repeat(np.asarray(scales, np.float64), g, axis=0)[:K]
    return ((q - z) * s).T


def dequant_gptq_alt(qweight: np.ndarray, qzeros: np.ndarray, scales: np.ndarray, g_idx: np.ndarray) -> np.ndarray:
    """conventional GPTQ (act-order / asymmetric; gemm_half_q_half_alt, gptq.rs:182-197): qweight [K/8, N] packed along K,
    qzeros [G, N/8] packed along N in natural order and stored MINUS ONE (GPTQ v1), group of row k = g_idx[k].
    -> W f64 [N, K] = (q[k, n] - (z[g_idx[k], n] + 1)) * s[g_idx[k], n]"""
    q = unpack_gptq(qweight).astype(np.float64)
    z = unpack_cols(qzeros).astype(np.float64)[np.asarray(g_idx, np.int64)] + 1.0
    s = np.asarray(scales, np.float64)[np.asarray(g_idx, np.int64)]
    return ((q - z) * s).T


# ---- Marlin checkpoint format (checkpoint_format == "marlin": tensors `B` u32 [K/16, 2N] and `s`, /root/reference/src/openai/models/linear.rs:219-251) ----
# The tile order is defined by the Marlin project (IST-DASLab/marlin, marlin/__init__.py: `_get_perms` and `Layer.pack`; third-party, not in
# /root/reference -- only its scale permutation is, linear.rs:341-352, and it is the same `_get_perms` function).  Restated from the published
# algorithm: w [K, N] -> 16 x 16 tiles -> rows of N * 16 values -> a fixed permutation inside every 1024 values -> 8 nibbles per word, strided.
def marlin_weight_perm() -> np.ndarray:
    perm = []
    for i in range(32):
        perm1 = []
        col = i // 4
        for block in (0, 1):
            for row in (2 * (i % 4), 2 * (i % 4) + 1, 2 * (i % 4 + 4), 2 * (i % 4 + 4) + 1):
                perm1.append(16 * row + col + 8 * block)
        for j in range(4):
            perm.extend(p + 256 * j for p in perm1)
    perm = np.asarray(perm, np.int64)
    interleave = np.asarray([0, 2, 4, 6, 1, 3, 5, 7])
    return perm.reshape(-1, 8)[:, interleave].ravel()


def pack_marlin(q: np.ndarray) -> np.ndarray:
    """q u8 [K, N] in 0..15 -> B u32 [K/16, 2N] (Marlin `Layer.pack`)."""
    K, N = q.shape
    perm = marlin_weight_perm()
    w = q.reshape(K // 16, 16, N // 16, 16).transpose(0, 2, 1, 3).reshape(K // 16, N * 16)
    res = w.reshape(-1, perm.size)[:, perm].reshape(w.shape).astype(np.uint32)
    out = np.zeros((res.shape[0], res.shape[1] // 8), np.uint32)
    for i in range(8):
        out |= res[:, i::8] << np.uint32(4 * i)
    return out


def unpack_marlin(B: np.ndarray, K: int, N: int) -> np.ndarray:
    """inverse of pack_marlin -> q u8 [K, N]."""
    perm = marlin_weight_perm()
    res = np.empty((K // 16, N * 16), np.uint8)
    for i in range(8):
        res[:, i::8] = (B >> np.uint32(4 * i)) & 0xF
    w = np.empty_like(res).reshape(-1, perm.size)
    w[:, perm] = res.reshape(-1, perm.size)
    return w.reshape(K // 16, N // 16, 16, 16).transpose(0, 2, 1, 3).reshape(K, N)
