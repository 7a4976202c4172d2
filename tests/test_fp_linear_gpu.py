"""GPU parity: weight-only low-precision float linears (K10 block-FP8, K11 NVFP4, K12 MXFP4) through the C ABI vs the fp64
oracle (oracle/fp_formats.py; parity unpinned upstream, element formats pinned to torch / OCP tables in test_oracle.py).
Tolerances: the SIMT kernels decode weights exactly and accumulate in fp32 -> the result differs from the fp64 target only
by fp32 summation noise and ONE final rounding to the activation dtype (rel-Frobenius <= 2.5e-3 bf16, 4e-4 f16).  The tcgen05
block-FP8 path (any m, 64 rows per pass) rounds scaled weights and activations to fp16 first: rel-Frobenius <= 1e-3 on f16 I/O."""
import os

import numpy as np
import pytest
import torch

import candle_vllm_b200 as pkg
from oracle import fp_formats as F
from tests.gpu_util import DEV, rel_fro

pytestmark = pytest.mark.gpu
TOL = {torch.bfloat16: 2.5e-3, torch.float16: 4e-4}


def _x(rng, m, k, dtype):
    x = torch.from_numpy(rng.standard_normal((m, k)).astype(np.float32)).to(DEV).to(dtype)
    return x, x.float().cpu().numpy()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("m,n,k,by,bx", [(32, 512, 1024, 128, 128), (1, 256, 256, 128, 128), (64, 384, 2048, 128, 128),
                                         (17, 200, 768, 128, 128), (32, 7168, 2048, 128, 128), (5, 128, 512, 64, 64),
                                         (200, 256, 512, 128, 128), (3, 96, 144, 32, 16)])
def test_fp8_block_matmul(dtype, m, n, k, by, bx):
    rng = np.random.default_rng(m * 7 + n + k)
    w, s = F.random_fp8(rng, n, k, by, bx)
    x, xf = _x(rng, m, k, dtype)
    bias = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(DEV).to(dtype) if m % 2 else None
    lin = pkg.LnFp8(torch.from_numpy(w).to(DEV), torch.from_numpy(s).to(DEV), bias, (by, bx))
    y = lin.forward(x).float().cpu().numpy()
    ref = F.linear(xf, F.dequant_fp8_block(w, s, by, bx), None if bias is None else bias.float().cpu().numpy())
    assert y.shape == (m, n) and np.isfinite(y).all()
    tol = max(TOL[dtype], 1e-3) if (k % 256 == 0 and bx % 64 == 0 and n % 4 == 0) else TOL[dtype]    # tcgen05 path (64 rows per pass): fp16 operands
    assert rel_fro(y, ref) < tol, rel_fro(y, ref)


def test_fp8_tensor_core_and_simt_paths_agree_and_scale_magnitude_is_irrelevant():
    """m <= 64 runs on the tcgen05 pipeline with scaled weights in fp16; a power-of-two range shift keeps them in fp16's
    normal range for ANY checkpoint scale magnitude (1e-7 .. 1e+2 here)."""
    rng = np.random.default_rng(3)
    m, n, k = 32, 1024, 4096
    x, xf = _x(rng, m, k, torch.float16)
    for mag in (1e-7, 2e-4, 1.0, 1e2):
        w, s = F.random_fp8(rng, n, k, 128, 128, scale_mag=mag)
        lin = pkg.LnFp8(torch.from_numpy(w).to(DEV), torch.from_numpy(s).to(DEV))
        y = lin.forward(x).float().cpu().numpy()
        ref = F.linear(xf, F.dequant_fp8_block(w, s, 128, 128))
        if not np.isfinite(ref).all() or np.abs(ref).max() > 6e4:
            continue                       # f16 output would overflow: not a kernel property
        assert rel_fro(y, ref) < 1e-3, (mag, rel_fro(y, ref))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("m,n,k", [(32, 512, 1024), (1, 64, 32), (9, 200, 2880), (70, 128, 256)])
def test_nvfp4_and_mxfp4_matmul(dtype, m, n, k):
    rng = np.random.default_rng(m + n + k)
    blocks = F.random_fp4(rng, n, k)
    x, xf = _x(rng, m, k, dtype)
    bias = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(DEV).to(dtype)
    # NVFP4
    sc = F.random_nvfp4_scales(rng, n, k)
    g = 1.0 / 448.0
    y = pkg.LnNvfp4(torch.from_numpy(blocks).to(DEV), torch.from_numpy(sc).to(DEV), g, 0.37, bias).forward(x).float().cpu().numpy()
    ref = F.linear(xf, F.dequant_nvfp4(blocks, sc, g), bias.float().cpu().numpy())
    assert rel_fro(y, ref) < TOL[dtype], rel_fro(y, ref)
    # MXFP4
    se = F.random_mxfp4_scales(rng, n, k)
    y = pkg.LnMxfp4(torch.from_numpy(blocks).to(DEV), torch.from_numpy(se).to(DEV)).forward(x).float().cpu().numpy()
    ref = F.linear(xf, F.dequant_mxfp4(blocks, se))
    assert rel_fro(y, ref) < TOL[dtype], rel_fro(y, ref)


@pytest.mark.parametrize("m,n,k", [(32, 1024, 2048), (64, 384, 512), (5, 4096, 256), (33, 132, 768)])
def test_fp4_on_the_tensor_core_pipeline_full_scale_range(m, n, k):
    """Decode shapes (m <= 64 per pass, k % 256 == 0) of NVFP4 / MXFP4 run on the tcgen05 dequant-into-TMEM pipeline: e2m1 nibbles are
    placed as f16 bit patterns (value * 2^-14), the block scale is applied as one HMUL2.  Scales here span the whole e4m3 code range --
    zero, subnormals, 448 -- and e8m0 exponents from 2^-20 up (the kernel represents 2^-20 .. 2^9 exactly)."""
    rng = np.random.default_rng(m * 7 + n + k)
    dtype = torch.float16
    blocks = F.random_fp4(rng, n, k)
    x, xf = _x(rng, m, k, dtype)
    sc = rng.integers(0, 0x7f, (n, k // 16), dtype=np.uint8)          # every non-negative finite e4m3 code, incl. 0 and subnormals
    g = 0.37
    y = pkg.LnNvfp4(torch.from_numpy(blocks).to(DEV), torch.from_numpy(sc).to(DEV), g, 1.0, None).forward(x).float().cpu().numpy()
    ref = F.linear(xf, F.dequant_nvfp4(blocks, sc, g))
    assert rel_fro(y, ref) < TOL[dtype], rel_fro(y, ref)
    se = rng.integers(107, 129, (n, k // 32), dtype=np.uint8)          # 2^-20 .. 2^1 (larger scales overflow the f16 OUTPUT)
    y = pkg.LnMxfp4(torch.from_numpy(blocks).to(DEV), torch.from_numpy(se).to(DEV)).forward(x).float().cpu().numpy()
    ref = F.linear(xf, F.dequant_mxfp4(blocks, se))
    assert np.isfinite(y).all() and rel_fro(y, ref) < 2 * TOL[dtype], rel_fro(y, ref)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_prefill_chunks_dequantise_once_and_use_the_dense_tensor_core_gemm(dtype):
    """m >= 512: weights -> 16 bit once (library scratch) + dense tcgen05 GEMM with the bias in its epilogue, for all three formats.
    Tolerance: operands rounded to the activation dtype (weights once, exactly decoded first), fp32 accumulate, one output rounding:
    rel-Frobenius 1e-3 (f16) / 6e-3 (bf16: 8-bit mantissa on weights, activations and output)."""
    rng = np.random.default_rng(11)
    m, n, k = 640, 512, 1024
    tol = 1e-3 if dtype == torch.float16 else 6e-3
    x, xf = _x(rng, m, k, dtype)
    bias = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(DEV).to(dtype)
    w, s = F.random_fp8(rng, n, k, 128, 128)
    y = pkg.LnFp8(torch.from_numpy(w).to(DEV), torch.from_numpy(s).to(DEV), bias, (128, 128)).forward(x).float().cpu().numpy()
    assert rel_fro(y, F.linear(xf, F.dequant_fp8_block(w, s, 128, 128), bias.float().cpu().numpy())) < tol
    blocks = F.random_fp4(rng, n, k)
    sc = F.random_nvfp4_scales(rng, n, k)
    y = pkg.LnNvfp4(torch.from_numpy(blocks).to(DEV), torch.from_numpy(sc).to(DEV), 1.0 / 448.0, 1.0, bias).forward(x).float().cpu().numpy()
    assert rel_fro(y, F.linear(xf, F.dequant_nvfp4(blocks, sc, 1.0 / 448.0), bias.float().cpu().numpy())) < tol
    se = F.random_mxfp4_scales(rng, n, k)
    y = pkg.LnMxfp4(torch.from_numpy(blocks).to(DEV), torch.from_numpy(se).to(DEV)).forward(x).float().cpu().numpy()
    assert rel_fro(y, F.linear(xf, F.dequant_mxfp4(blocks, se))) < tol


def test_fp_linear_argument_errors():
    w = torch.zeros((128, 256), dtype=torch.uint8, device=DEV)
    with pytest.raises(pkg.BackendError, match="weight_scale must be f32"):
        pkg.LnFp8(w, torch.zeros((2, 2), device=DEV))
    lin = pkg.LnFp8(w, torch.ones((1, 2), device=DEV))
    with pytest.raises(pkg.BackendError, match="f16 or bf16"):
        lin.forward(torch.zeros((2, 256), device=DEV))
    with pytest.raises(pkg.BackendError, match="shape mismatch"):
        lin.forward(torch.zeros((2, 128), dtype=torch.float16, device=DEV))
    with pytest.raises(pkg.BackendError, match="scales must be"):
        pkg.LnMxfp4(torch.zeros((8, 64), dtype=torch.uint8, device=DEV), torch.zeros((8, 3), dtype=torch.uint8, device=DEV))
