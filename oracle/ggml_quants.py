"""GGML k-quant block codecs and the quantised mat-mul, restated in numpy.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.

The reference consumes GGUF tensor bytes verbatim
(``src/openai/models/layers/quantized_var_builder.rs:118,193`` ->
``QMatMul::from_arc`` ``src/openai/models/quantized_llama.rs:332-334``) and runs
``QMatMul::forward`` on an f32 activation (``src/openai/models/linear.rs:765-806``).
The arithmetic itself lives in candle-core (fork @cafd231, v0.8.3, NOT in
/root/reference): ``k_quants.rs`` = a Rust transliteration of llama.cpp's
``ggml-quants.c``.  The block formats are the published GGUF spec (SURVEY.md
Appendix A).  Pinned against ``gguf.quants.dequantize`` (``tests/golden``).

Two matmul semantics are provided:
  * ``qmatmul_dequant``  -- y = x @ dequant(W)^T in fp64/fp32: the mathematical
    target ("fp32 oracle" of SURVEY.md §8c).
  * ``qmatmul_q8k``      -- what the reference's CPU path does [UPSTREAM]: each
    activation row is quantised to Q8_K (256-wide, int8 + f32 scale) and the dot
    product is integer; reported as the reference's own noise floor.
"""
from __future__ import annotations

import numpy as np

QK_K = 256
Q4_K_BLOCK_BYTES = 144   # f16 d | f16 dmin | u8 scales[12] | u8 qs[128]
Q6_K_BLOCK_BYTES = 210   # u8 ql[128] | u8 qh[64] | i8 scales[16] | f16 d
Q8_0_BLOCK_BYTES = 34    # f16 d | i8 qs[32]
QK8_0 = 32

GGML_TYPE_Q8_0 = 8
GGML_TYPE_Q4_K = 12
GGML_TYPE_Q6_K = 14

BLOCK_BYTES = {GGML_TYPE_Q8_0: Q8_0_BLOCK_BYTES, GGML_TYPE_Q4_K: Q4_K_BLOCK_BYTES,
               GGML_TYPE_Q6_K: Q6_K_BLOCK_BYTES}
BLOCK_ELEMS = {GGML_TYPE_Q8_0: QK8_0, GGML_TYPE_Q4_K: QK_K, GGML_TYPE_Q6_K: QK_K}


# --------------------------------------------------------------------------------------
# Q4_K
# --------------------------------------------------------------------------------------
def q4k_unpack_scales(scales: np.ndarray):
    """scales u8[..., 12] -> (sc u8[..., 8], m u8[..., 8])  (6-bit each).

    ggml ``get_scale_min_k4``: j<4: sc=q[j]&63, m=q[j+4]&63;
    j>=4: sc=(q[j+4]&0xF)|((q[j-4]>>6)<<4), m=(q[j+4]>>4)|((q[j]>>6)<<4).
    """
    q = scales.astype(np.uint8)
    sc = np.empty(q.shape[:-1] + (8,), np.uint8)
    m = np.empty_like(sc)
    for j in range(4):
        sc[..., j] = q[..., j] & 63
        m[..., j] = q[..., j + 4] & 63
    for j in range(4, 8):
        sc[..., j] = (q[..., j + 4] & 0x0F) | ((q[..., j - 4] >> 6) << 4)
        m[..., j] = (q[..., j + 4] >> 4) | ((q[..., j] >> 6) << 4)
    return sc, m


def q4k_fields(blocks: np.ndarray):
    """blocks u8[nb, 144] -> d f32[nb], dmin f32[nb], sc u8[nb,8], m u8[nb,8], q u8[nb,256]."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, Q4_K_BLOCK_BYTES)
    d = blocks[:, 0:2].copy().view(np.float16).astype(np.float32)[:, 0]
    dmin = blocks[:, 2:4].copy().view(np.float16).astype(np.float32)[:, 0]
    sc, m = q4k_unpack_scales(blocks[:, 4:16])
    qs = blocks[:, 16:144].reshape(-1, 4, 32)
    q = np.empty((blocks.shape[0], 8, 32), np.uint8)
    q[:, 0::2, :] = qs & 0x0F          # chunk c low nibbles  -> sub-block 2c
    q[:, 1::2, :] = qs >> 4            # chunk c high nibbles -> sub-block 2c+1
    return d, dmin, sc, m, q.reshape(-1, 256)


def dequantize_q4k(blocks: np.ndarray) -> np.ndarray:
    """u8[nb,144] -> f32[nb,256];  w = d*sc[j]*q - dmin*m[j]  (fp32 arithmetic, ggml order)."""
    d, dmin, sc, m, q = q4k_fields(blocks)
    d1 = (d[:, None] * sc.astype(np.float32)).astype(np.float32)      # [nb, 8]
    m1 = (dmin[:, None] * m.astype(np.float32)).astype(np.float32)
    q = q.reshape(-1, 8, 32).astype(np.float32)
    w = d1[:, :, None] * q - m1[:, :, None]
    return w.reshape(-1, 256).astype(np.float32)


# --------------------------------------------------------------------------------------
# Q6_K
# --------------------------------------------------------------------------------------
def q6k_fields(blocks: np.ndarray):
    """blocks u8[nb,210] -> d f32[nb], scales i8[nb,16], q i8[nb,256] (already minus 32)."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, Q6_K_BLOCK_BYTES)
    nb = blocks.shape[0]
    ql = blocks[:, 0:128].reshape(nb, 2, 64)
    qh = blocks[:, 128:192].reshape(nb, 2, 32)
    scales = blocks[:, 192:208].copy().view(np.int8)
    d = blocks[:, 208:210].copy().view(np.float16).astype(np.float32)[:, 0]
    q = np.empty((nb, 2, 4, 32), np.int16)
    for half in range(2):
        lo = ql[:, half, :]
        h = qh[:, half, :]
        q[:, half, 0] = (lo[:, 0:32] & 0x0F) | (((h >> 0) & 3) << 4)
        q[:, half, 1] = (lo[:, 32:64] & 0x0F) | (((h >> 2) & 3) << 4)
        q[:, half, 2] = (lo[:, 0:32] >> 4) | (((h >> 4) & 3) << 4)
        q[:, half, 3] = (lo[:, 32:64] >> 4) | (((h >> 6) & 3) << 4)
    q = (q - 32).astype(np.int8).reshape(nb, 256)
    return d, scales, q


def dequantize_q6k(blocks: np.ndarray) -> np.ndarray:
    d, scales, q = q6k_fields(blocks)
    s = (d[:, None] * scales.astype(np.float32)).astype(np.float32)    # [nb,16]
    w = s[:, :, None] * q.reshape(-1, 16, 16).astype(np.float32)
    return w.reshape(-1, 256).astype(np.float32)


# --------------------------------------------------------------------------------------
# Q8_0
# --------------------------------------------------------------------------------------
def dequantize_q8_0(blocks: np.ndarray) -> np.ndarray:
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, Q8_0_BLOCK_BYTES)
    d = blocks[:, 0:2].copy().view(np.float16).astype(np.float32)[:, 0]
    q = blocks[:, 2:34].copy().view(np.int8).astype(np.float32)
    return (d[:, None] * q).astype(np.float32)


def quantize_q8_0(x: np.ndarray) -> np.ndarray:
    """f32[..., 32k] -> u8 blocks; d = max|x|/127 (f16), q = round(x/d)  (ggml quantize_row_q8_0_ref)."""
    x = np.asarray(x, np.float32).reshape(-1, QK8_0)
    amax = np.abs(x).max(axis=1)
    d = (amax / 127.0).astype(np.float32)
    idd = np.where(d != 0, 1.0 / np.where(d != 0, d, 1), 0).astype(np.float32)
    v = x * idd[:, None]
    q = (np.sign(v) * np.floor(np.abs(v) + 0.5)).astype(np.int8)   # roundf: half away from zero
    out = np.empty((x.shape[0], Q8_0_BLOCK_BYTES), np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:] = q.view(np.uint8)
    return out


def dequantize(blocks: np.ndarray, ggml_type: int) -> np.ndarray:
    if ggml_type == GGML_TYPE_Q4_K:
        return dequantize_q4k(blocks)
    if ggml_type == GGML_TYPE_Q6_K:
        return dequantize_q6k(blocks)
    if ggml_type == GGML_TYPE_Q8_0:
        return dequantize_q8_0(blocks)
    raise ValueError(f"unsupported ggml type {ggml_type}")


def dequantize_weight(wbytes: np.ndarray, ggml_type: int, n: int, k: int) -> np.ndarray:
    """Row-major quantised weight [N, K/blk] blocks -> f32 [N, K]."""
    bb, be = BLOCK_BYTES[ggml_type], BLOCK_ELEMS[ggml_type]
    assert k % be == 0
    blocks = np.ascontiguousarray(wbytes, np.uint8).reshape(n * (k // be), bb)
    return dequantize(blocks, ggml_type).reshape(n, k)


# --------------------------------------------------------------------------------------
# Q8_K activation quantisation (CPU reference path, [UPSTREAM] candle k_quants.rs
# BlockQ8K::from_float, a transliteration of ggml quantize_row_q8_K_ref)
# --------------------------------------------------------------------------------------
def quantize_q8k(x: np.ndarray):
    """f32[..., 256k] -> (d f32[nb], qs i8[nb,256], bsums i16[nb,16]).

    iscale = -128/max (max = signed value of largest magnitude), q = min(127, round(iscale*x)),
    d = 1/iscale.  candle keeps llama.cpp's original -128 constant; unverifiable here
    (dependency not vendored) -- only the CPU-baseline noise floor depends on it.
    """
    x = np.asarray(x, np.float32).reshape(-1, QK_K)
    idx = np.abs(x).argmax(axis=1)
    mx = x[np.arange(x.shape[0]), idx]
    nz = mx != 0
    iscale = np.where(nz, -128.0 / np.where(nz, mx, 1), 0).astype(np.float32)
    v = iscale[:, None] * x
    v = np.sign(v) * np.floor(np.abs(v) + 0.5)               # f32::round (half away from zero)
    q = np.minimum(v, 127).astype(np.int8)
    d = np.where(nz, 1.0 / np.where(nz, iscale, 1), 0).astype(np.float32)
    bsums = q.reshape(-1, 16, 16).astype(np.int32).sum(axis=2).astype(np.int16)
    return d, q, bsums


def vec_dot_q4k_q8k(wblocks: np.ndarray, xd, xq, xbsums) -> np.ndarray:
    """Integer dot of nbk Q4_K blocks with nbk Q8_K blocks -> scalar f32 (ggml_vec_dot_q4_K_q8_K).

    wblocks u8[..., nbk, 144]; xd f32[nbk]; xq i8[nbk,256]; xbsums i16[nbk,16].  Returns f32[...].
    """
    lead = wblocks.shape[:-2]
    nbk = wblocks.shape[-2]
    d, dmin, sc, m, q = q4k_fields(wblocks.reshape(-1, Q4_K_BLOCK_BYTES))
    d = d.reshape(-1, nbk); dmin = dmin.reshape(-1, nbk)
    sc = sc.reshape(-1, nbk, 8).astype(np.int32); m = m.reshape(-1, nbk, 8).astype(np.int32)
    q = q.reshape(-1, nbk, 8, 32).astype(np.int32)
    xq32 = xq.reshape(nbk, 8, 32).astype(np.int32)
    sub = (q * xq32[None]).sum(axis=3)                       # [R, nbk, 8]
    isum = (sub * sc).sum(axis=2)                            # [R, nbk]
    bs = xbsums.reshape(nbk, 8, 2).astype(np.int32).sum(axis=2)   # [nbk, 8]
    msum = (m * bs[None]).sum(axis=2)
    acc = (xd[None] * d) * isum.astype(np.float32) - (xd[None] * dmin) * msum.astype(np.float32)
    return acc.astype(np.float32).sum(axis=1, dtype=np.float32).reshape(lead)


def vec_dot_q6k_q8k(wblocks: np.ndarray, xd, xq) -> np.ndarray:
    lead = wblocks.shape[:-2]
    nbk = wblocks.shape[-2]
    d, scales, q = q6k_fields(wblocks.reshape(-1, Q6_K_BLOCK_BYTES))
    d = d.reshape(-1, nbk)
    scales = scales.reshape(-1, nbk, 16).astype(np.int32)
    q = q.reshape(-1, nbk, 16, 16).astype(np.int32)
    xq32 = xq.reshape(nbk, 16, 16).astype(np.int32)
    isum = ((q * xq32[None]).sum(axis=3) * scales).sum(axis=2)
    acc = (xd[None] * d) * isum.astype(np.float32)
    return acc.astype(np.float32).sum(axis=1, dtype=np.float32).reshape(lead)


# --------------------------------------------------------------------------------------
# mat-mul
# --------------------------------------------------------------------------------------
def qmatmul_dequant(x: np.ndarray, wbytes: np.ndarray, ggml_type: int, n: int, k: int,
                    acc_dtype=np.float64) -> np.ndarray:
    """y[M,N] = x[M,K] @ dequant(W[N,K])^T, accumulated in ``acc_dtype`` -> f32."""
    w = dequantize_weight(wbytes, ggml_type, n, k)
    return (np.asarray(x, acc_dtype) @ w.T.astype(acc_dtype)).astype(np.float32)


def qmatmul_q8k(x: np.ndarray, wbytes: np.ndarray, ggml_type: int, n: int, k: int) -> np.ndarray:
    """Reference CPU semantics [UPSTREAM]: activations -> Q8_K, integer dot per block."""
    x = np.asarray(x, np.float32).reshape(-1, k)
    nbk = k // QK_K
    bb = BLOCK_BYTES[ggml_type]
    wb = np.ascontiguousarray(wbytes, np.uint8).reshape(n, nbk, bb)
    out = np.empty((x.shape[0], n), np.float32)
    for r in range(x.shape[0]):
        xd, xq, xbs = quantize_q8k(x[r])
        if ggml_type == GGML_TYPE_Q4_K:
            out[r] = vec_dot_q4k_q8k(wb, xd, xq, xbs)
        elif ggml_type == GGML_TYPE_Q6_K:
            out[r] = vec_dot_q6k_q8k(wb, xd, xq)
        else:
            raise ValueError("q8k path: Q4_K / Q6_K only")
    return out


# --------------------------------------------------------------------------------------
# synthetic weights: random valid blocks
# --------------------------------------------------------------------------------------
def random_q4k(rng: np.random.Generator, n: int, k: int, d_scale: float = 2.0 ** -8) -> np.ndarray:
    """u8[n, k/256*144]: d,dmin ~ f16 U(0.5,2)*d_scale, 12 scale bytes + 128 nibble bytes uniform."""
    nb = n * (k // QK_K)
    blocks = np.empty((nb, Q4_K_BLOCK_BYTES), np.uint8)
    d = (rng.uniform(0.5, 2.0, nb) * d_scale).astype(np.float16)
    dmin = (rng.uniform(0.5, 2.0, nb) * d_scale).astype(np.float16)
    blocks[:, 0:2] = d.view(np.uint8).reshape(-1, 2)
    blocks[:, 2:4] = dmin.view(np.uint8).reshape(-1, 2)
    blocks[:, 4:] = rng.integers(0, 256, (nb, 140), dtype=np.uint8)
    return blocks.reshape(n, -1)


def random_q6k(rng: np.random.Generator, n: int, k: int, d_scale: float = 2.0 ** -10) -> np.ndarray:
    nb = n * (k // QK_K)
    blocks = np.empty((nb, Q6_K_BLOCK_BYTES), np.uint8)
    blocks[:, 0:208] = rng.integers(0, 256, (nb, 208), dtype=np.uint8)
    d = (rng.uniform(0.5, 2.0, nb) * d_scale).astype(np.float16)
    blocks[:, 208:210] = d.view(np.uint8).reshape(-1, 2)
    return blocks.reshape(n, -1)


def random_q8_0(rng: np.random.Generator, n: int, k: int, d_scale: float = 2.0 ** -10) -> np.ndarray:
    nb = n * (k // QK8_0)
    blocks = np.empty((nb, Q8_0_BLOCK_BYTES), np.uint8)
    d = (rng.uniform(0.5, 2.0, nb) * d_scale).astype(np.float16)
    blocks[:, 0:2] = d.view(np.uint8).reshape(-1, 2)
    blocks[:, 2:] = rng.integers(0, 256, (nb, 32), dtype=np.uint8)
    return blocks.reshape(n, -1)


def random_weight(rng, ggml_type: int, n: int, k: int) -> np.ndarray:
    return {GGML_TYPE_Q4_K: random_q4k, GGML_TYPE_Q6_K: random_q6k,
            GGML_TYPE_Q8_0: random_q8_0}[ggml_type](rng, n, k)
