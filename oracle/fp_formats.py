"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the weight-only low-precision float linears of the reference:
block-scaled FP8 (LnFp8), NVFP4 (LnNvfp4) and MXFP4 (LnMxfp4).

The kernels themselves live in attention-rs (not in /root/reference), and the reference holds no tests or fixtures for
them: PARITY UNPINNED.  This file follows the tensor layouts visible at the call sites
  /root/reference/src/openai/models/linear.rs:944-973   (fp8: weight e4m3 [N,K], weight_scale f32 [ceil(N/by), ceil(K/bx)])
  /root/reference/src/openai/models/linear.rs:1812-1853 (nvfp4: blocks u8 [N,K/2], scales e4m3 [N,K/16], global scale)
  /root/reference/src/openai/models/linear.rs:1686-1700 (mxfp4: blocks u8 [N,K/2], scales e8m0 [N,K/32])
and the published element formats (OCP MX v1.0: e2m1, e8m0; e4m3fn as in torch.float8_e4m3fn).  The element decoders are
pinned against torch's dtypes in tests/test_oracle.py.
"""
import numpy as np

from .cache_ops import e4m3_to_f32          # exact e4m3fn decode (pinned to torch.float8_e4m3fn)

E2M1 = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], np.float32)


def e2m1_to_f32(nib: np.ndarray) -> np.ndarray:
    nib = np.asarray(nib, np.uint8)
    v = E2M1[nib & 7]
    return np.where(nib & 8, -v, v).astype(np.float32)


def e8m0_to_f32(b: np.ndarray) -> np.ndarray:
    b = np.asarray(b, np.uint8).astype(np.int32)
    out = np.ldexp(np.float64(1.0), b - 127)
    out = np.where(b == 255, np.nan, out)
    return out.astype(np.float64)


def unpack_fp4(blocks: np.ndarray) -> np.ndarray:
    """u8 [N, K/2] -> f32 [N, K]; low nibble = even k."""
    blocks = np.asarray(blocks, np.uint8)
    lo, hi = blocks & 0xF, blocks >> 4
    out = np.empty(blocks.shape[:-1] + (blocks.shape[-1] * 2,), np.float32)
    out[..., 0::2] = e2m1_to_f32(lo)
    out[..., 1::2] = e2m1_to_f32(hi)
    return out


def dequant_fp8_block(weight: np.ndarray, scale: np.ndarray, by: int, bx: int) -> np.ndarray:
    n, k = weight.shape
    w = e4m3_to_f32(weight).astype(np.float64)
    s = np.repeat(np.repeat(np.asarray(scale, np.float64), by, axis=0)[:n], bx, axis=1)[:, :k]
    return w * s


def dequant_nvfp4(blocks: np.ndarray, scales: np.ndarray, global_scale: float) -> np.ndarray:
    w = unpack_fp4(blocks).astype(np.float64)
    s = e4m3_to_f32(scales).astype(np.float64) * np.float64(np.float32(global_scale))
    return w * np.repeat(s, 16, axis=1)


def dequant_mxfp4(blocks: np.ndarray, scales: np.ndarray) -> np.ndarray:
    w = unpack_fp4(blocks).astype(np.float64)
    return w * np.repeat(e8m0_to_f32(scales), 32, axis=1)


def linear(x: np.ndarray, w: np.ndarray, bias=None) -> np.ndarray:
    """fp64 x . w^T (+ bias): the mathematical target the kernels are compared with."""
    y = np.asarray(x, np.float64) @ np.asarray(w, np.float64).T
    if bias is not None:
        y = y + np.asarray(bias, np.float64)[None, :]
    return y


# ---- seeded generators of checkpoint-like tensors ------------------------------------------------------------------
def random_fp8(rng, n, k, by=128, bx=128, scale_mag=2e-4):
    """e4m3 bytes without NaN codes + per-tile scales like a DeepSeek-style checkpoint (weights ~ N(0, 0.02))."""
    weight = rng.integers(0, 256, (n, k), dtype=np.uint8)
    weight[(weight & 0x7F) == 0x7F] = 0x7E                     # no NaN
    scale = (rng.uniform(0.5, 2.0, ((n + by - 1) // by, (k + bx - 1) // bx)) * scale_mag).astype(np.float32)
    return weight, scale


def random_fp4(rng, n, k):
    return rng.integers(0, 256, (n, k // 2), dtype=np.uint8)


def random_nvfp4_scales(rng, n, k):
    s = rng.integers(0x28, 0x58, (n, k // 16), dtype=np.uint8)     # positive e4m3 in [2^-2, 2^4)
    return s


def random_mxfp4_scales(rng, n, k):
    return rng.integers(127 - 10, 127 - 2, (n, k // 32), dtype=np.uint8)
