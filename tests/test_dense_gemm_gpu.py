"""GPU parity of the dense tcgen05 GEMM (Linear::forward on 16-bit weights, linear.rs:124-172) and of the large-m QMatMul path that
dequantises once and reuses it (prefill chunks).  Oracle: fp64 matmul of the same 16-bit inputs; tolerance = the output rounding
(2^-9 bf16 / 2^-12 f16 per element, rel-Frobenius)."""
import numpy as np
import pytest
import torch

import candle_vllm_b200 as pkg
from oracle import ggml_quants as G
from tests.gpu_util import DEV, rel_fro

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,k,bias", [
    (1, 256, 128, False),          # swap-AB, single tile, k = 2 blocks
    (32, 4096, 4096, True),        # decode shape (wo), swap-AB N = 32
    (64, 1024, 512, False),        # swap-AB N = 64
    (33, 200, 72, True),           # ragged everything (k % 64 != 0, n % 128 != 0)
    (128, 256, 64, False),         # one 128 x 256 tile, one k block
    (300, 520, 200, True),         # ragged m / n / k on the 128 x 256 path
    (1024, 2048, 1024, False),     # several tiles per CTA: the double-buffered accumulator, the stage ring wrapping
])
def test_linear_matches_oracle(dtype, m, n, k, bias):
    rng = np.random.default_rng(m * 31 + n + k)
    x = torch.from_numpy(rng.standard_normal((m, k)).astype(np.float32)).to(DEV).to(dtype)
    w = torch.from_numpy((rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)).to(DEV).to(dtype)
    b = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(DEV).to(dtype) if bias else None
    y = pkg.Linear(w, b).forward(x)
    assert y.dtype == dtype and y.shape == (m, n)
    ref = x.double().cpu().numpy() @ w.double().cpu().numpy().T
    if bias:
        ref = ref + b.double().cpu().numpy()
    assert rel_fro(y.float().cpu().numpy(), ref) < (4e-3 if dtype == torch.bfloat16 else 6e-4)
    assert torch.equal(y, pkg.Linear(w, b).forward(x))           # no atomics anywhere: bitwise repeatable


def test_linear_leading_dims_and_errors():
    w = torch.zeros((64, 128), dtype=torch.float16, device=DEV)
    assert pkg.Linear(w).forward(torch.zeros((2, 3, 128), dtype=torch.float16, device=DEV)).shape == (2, 3, 64)
    with pytest.raises(pkg.BackendError, match="dtype"):
        pkg.Linear(w).forward(torch.zeros((2, 128), dtype=torch.bfloat16, device=DEV))
    with pytest.raises(pkg.BackendError, match="shape mismatch"):
        pkg.Linear(w).forward(torch.zeros((2, 64), dtype=torch.float16, device=DEV))


@pytest.mark.parametrize("ggml_type,m,n,k", [(12, 512, 256, 512), (12, 777, 1024, 2048), (14, 600, 384, 256), (8, 512, 128, 64)])
def test_qmatmul_prefill_chunk_dequantises_once(ggml_type, m, n, k):
    """m >= 512: W -> fp16 once + dense tcgen05 GEMM.  Same contract as the decode kernel (operands rounded once to fp16, fp32
    accumulate): rel-Frobenius < 1e-3 against the fp64 dequant-matmul oracle."""
    rng = np.random.default_rng(ggml_type + m)
    wb = G.random_weight(rng, ggml_type, n, k)
    x = rng.standard_normal((m, k)).astype(np.float32)
    q = pkg.QTensor(torch.from_numpy(wb).to(DEV), ggml_type, (n, k))
    y = pkg.QMatMul.from_arc(q).forward(torch.from_numpy(x).to(DEV))
    ref = G.qmatmul_dequant(x, wb, ggml_type, n, k)
    assert y.dtype == torch.float32 and rel_fro(y.cpu().numpy(), ref) < 1e-3
