"""GPU parity: QMatMul::forward (K6) for GGML Q4_K / Q6_K / Q8_0 through the C ABI vs the fp64
dequant-matmul oracle.  Tolerance: rel-Frobenius <= 1e-3; the reference's own
8-bit-activation path (Q8_K oracle) is reported alongside as its noise floor (~4e-3..7e-3)."""
import numpy as np
import pytest
import torch

import candle_vllm_b200 as pkg
from oracle import ggml_quants as G
from tests.gpu_util import DEV, rel_fro

pytestmark = pytest.mark.gpu
TYPES = [(pkg.GgmlType.Q4_K, "q4k"), (pkg.GgmlType.Q6_K, "q6k"), (pkg.GgmlType.Q8_0, "q8_0")]


@pytest.mark.parametrize("t,_n", TYPES)
def test_dequantize_bit_exact(t, _n):
    rng = np.random.default_rng(0)
    w = G.random_weight(rng, t, 24, 1024)
    q = pkg.QTensor.from_numpy(w, t, (24, 1024))
    got = q.dequantize().cpu().numpy()
    assert np.array_equal(got, G.dequantize_weight(w, t, 24, 1024))


@pytest.mark.parametrize("t,_n", TYPES)
@pytest.mark.parametrize("m,n,k", [(1, 128, 256), (32, 256, 1024), (32, 1024, 4096), (7, 136, 512), (33, 128, 768),
                                    (64, 384, 2048), (200, 128, 512)])
def test_qmatmul_matches_oracle(t, _n, m, n, k):
    rng = np.random.default_rng(m * 131 + n)
    w = G.random_weight(rng, t, n, k)
    x = rng.standard_normal((m, k)).astype(np.float32)
    mm = pkg.QMatMul.from_arc(pkg.QTensor.from_numpy(w, t, (n, k)))
    y = mm.forward(torch.from_numpy(x).to(DEV)).cpu().numpy()
    ref = G.qmatmul_dequant(x, w, t, n, k)
    assert y.shape == (m, n) and np.isfinite(y).all()
    assert rel_fro(y, ref) < 1e-3, rel_fro(y, ref)
    if t != pkg.GgmlType.Q8_0 and m <= 8:
        # the reference's CPU path (Q8_K activations) is further from the fp target than we are
        assert rel_fro(y, ref) < rel_fro(G.qmatmul_q8k(x, w, t, n, k), ref)


def test_qmatmul_llama_shapes_and_linearity():
    # metric shapes (M = 32): wq 4096x4096, wk 1024x4096, w1 14336x4096 (sampled rows), w2 4096x14336
    rng = np.random.default_rng(5)
    for n, k in ((4096, 4096), (1024, 4096), (4096, 14336)):
        w = G.random_weight(rng, pkg.GgmlType.Q4_K, n, k)
        x = rng.standard_normal((32, k)).astype(np.float32)
        mm = pkg.QMatMul(pkg.QTensor.from_numpy(w, pkg.GgmlType.Q4_K, (n, k)))
        xt = torch.from_numpy(x).to(DEV)
        y = mm.forward(xt)
        rows = rng.choice(n, 64, replace=False)
        wb = w.reshape(n, -1)[rows]
        ref = G.qmatmul_dequant(x, wb, pkg.GgmlType.Q4_K, 64, k)
        assert rel_fro(y[:, rows].cpu().numpy(), ref) < 1e-3
        # size-independent property: linearity in x (fp16 activation rounding bounds the residual)
        y2 = mm.forward(xt * 2.0)                         # scaling by 2 is exact in fp16
        assert torch.allclose(y2, 2 * y, rtol=1e-4, atol=1e-4 * float(y.abs().max()))


@pytest.mark.parametrize("t", [pkg.GgmlType.Q4_K, pkg.GgmlType.Q6_K])
@pytest.mark.parametrize("m,n,k", [(32, 4096, 4096), (32, 6144, 4096), (17, 1024, 2048), (32, 4096, 14336), (64, 2048, 4096), (5, 200, 512)])
def test_qmatmul_slabs_sum_to_product_and_are_deterministic(t, m, n, k):
    """slab mode (what the decode engine runs): no atomics, nothing pre-zeroed, every slab element written; the slabs add
    up to the product and two runs agree bit for bit."""
    if t == pkg.GgmlType.Q6_K and k == 14336:
        pytest.skip("Q6_K tensor-core path needs k % 2048 == 0; covered by the generic path elsewhere")
    rng = np.random.default_rng(n + k + m)
    w = G.random_weight(rng, t, n, k)
    x = rng.standard_normal((m, k)).astype(np.float32)
    mm = pkg.QMatMul(pkg.QTensor.from_numpy(w, t, (n, k)))
    xt = torch.from_numpy(x).to(DEV)
    a = mm.forward_slabs(xt)
    b = mm.forward_slabs(xt)
    assert 1 <= a.shape[0] <= 9 and torch.isfinite(a).all()
    assert torch.equal(a, b)
    rows = rng.choice(n, min(n, 96), replace=False)
    ref = G.qmatmul_dequant(x, w.reshape(n, -1)[rows], t, len(rows), k)
    assert rel_fro(a.sum(0)[:, rows].cpu().numpy(), ref) < 1e-3
    if (n // 128) * (k // 256) >= 2 * 148 and n // 128 < 4 * 148:
        assert a.shape[0] > 1            # the metric shapes really are split over K


def test_qmatmul_batched_leading_dims_and_f16_input():
    rng = np.random.default_rng(6)
    w = G.random_weight(rng, pkg.GgmlType.Q4_K, 256, 512)
    mm = pkg.QMatMul(pkg.QTensor.from_numpy(w, pkg.GgmlType.Q4_K, (256, 512)))
    x = torch.from_numpy(rng.standard_normal((2, 3, 512)).astype(np.float32)).to(DEV)
    y = mm.forward(x)
    assert y.shape == (2, 3, 256)
    y16 = mm.forward(x.half())
    assert torch.allclose(y16, y, rtol=1e-5, atol=1e-6 * float(y.abs().max()))   # same fp16-activation contract (fp32 atomics may reorder sums)
    with pytest.raises(pkg.BackendError, match="shape mismatch"):
        mm.forward(x[..., :256])


def test_small_ops_match_oracle():
    from oracle import attention as OA
    rng = np.random.default_rng(7)
    x = rng.standard_normal((32, 4096)).astype(np.float32) * 3
    w = rng.uniform(0.5, 1.5, 4096).astype(np.float32)
    y = pkg.rms_norm(torch.from_numpy(x).to(DEV), torch.from_numpy(w).to(DEV), 1e-5).cpu().numpy()
    assert np.allclose(y, OA.rms_norm(x, w, 1e-5), rtol=2e-5, atol=1e-6)
    cos, sin = OA.rope_tables(128, 512, 500000.0)
    q = rng.standard_normal((5, 32, 128)).astype(np.float32); k = rng.standard_normal((5, 8, 128)).astype(np.float32)
    pos = np.array([0, 1, 17, 300, 511])
    for inter in (True, False):
        qt, kt = torch.from_numpy(q).to(DEV), torch.from_numpy(k).to(DEV)
        pkg.fused_rope(qt, kt, torch.from_numpy(cos).to(DEV), torch.from_numpy(sin).to(DEV), torch.from_numpy(pos).to(DEV), inter)
        assert np.allclose(qt.cpu().numpy(), OA.apply_rope(q, cos, sin, pos, inter), atol=2e-6)
        assert np.allclose(kt.cpu().numpy(), OA.apply_rope(k, cos, sin, pos, inter), atol=2e-6)
    g = rng.standard_normal((32, 1000)).astype(np.float32) * 4; u = rng.standard_normal((32, 1000)).astype(np.float32)
    s = pkg.silu_mul(torch.from_numpy(g).to(DEV), torch.from_numpy(u).to(DEV)).cpu().numpy()
    assert np.allclose(s, OA.silu_mul(g, u), rtol=1e-5, atol=1e-6)
    lg = rng.standard_normal((32, 128256)).astype(np.float32)
    lg[3, 77] = lg[3, 99999] = 50.0                        # tie -> first index
    am = pkg.argmax(torch.from_numpy(lg).to(DEV)).cpu().numpy()
    assert np.array_equal(am, lg.argmax(axis=1))


def test_qlinear_dispatch_like_reference():
    """QLinear (linear.rs:845-916): GGUF tensor -> QMatMul on f32-cast x (f32 out, [b,1,d] squeezed and restored, bias added);
    transposed GGUF tensor -> dequantise + dense matmul; GPTQ tuple -> gptq_matmul."""
    from oracle import gptq as OG
    rng = np.random.default_rng(9)
    n, k = 256, 512
    w = G.random_weight(rng, pkg.GgmlType.Q4_K, n, k)
    qt = pkg.QTensor.from_numpy(w, pkg.GgmlType.Q4_K, (n, k))
    bias = torch.from_numpy(rng.standard_normal(n).astype(np.float32)).to(DEV)
    lin = pkg.QLinear.from_qtensor(qt, bias)
    x = torch.from_numpy(rng.standard_normal((6, 1, k)).astype(np.float32)).to(DEV).to(torch.bfloat16)
    y = lin.forward(x)
    assert y.shape == (6, 1, n) and y.dtype == torch.float32
    ref = G.qmatmul_dequant(x.float().cpu().numpy().reshape(6, k), w, pkg.GgmlType.Q4_K, n, k) + bias.cpu().numpy()
    assert rel_fro(y.cpu().numpy().reshape(6, n), ref) < 1e-3
    y2 = lin.forward(x.reshape(2, 3, k))                                  # seq_len > 1: shape kept
    assert y2.shape == (2, 3, n) and torch.allclose(y2.reshape(6, n), y.reshape(6, n), rtol=1e-5, atol=1e-5)
    # transposed weight: the tensor holds W^T [in = 256, out = 512]
    lin_t = pkg.QLinear.from_qtensor(qt, None, transposed_weight=True)
    xt = torch.from_numpy(rng.standard_normal((4, n)).astype(np.float32)).to(DEV).half()
    yt = lin_t.forward(xt)
    wd = G.dequantize_weight(w, pkg.GgmlType.Q4_K, n, k)
    assert yt.shape == (4, k) and rel_fro(yt.float().cpu().numpy(), xt.float().cpu().numpy() @ wd) < 2e-3
    # GPTQ tuple
    q = rng.integers(0, 16, (k, n), dtype=np.uint8)
    sc = torch.from_numpy(rng.uniform(0.005, 0.02, (k // 128, n)).astype(np.float32)).to(DEV).half()
    w_m = pkg.marlin_weight_repack(torch.from_numpy(OG.pack_gptq(q).view(np.int32)).to(DEV), 4, False)
    gl = pkg.QLinear.from_gptq(w_m, pkg.marlin_permute_scales(sc, k, n, 128), None, None, torch.zeros(n, dtype=torch.int32, device=DEV), 128, 4,
                               bias=bias.half())
    xg = torch.from_numpy(rng.standard_normal((2, 3, k)).astype(np.float32)).to(DEV).half()
    yg = gl.forward(xg)
    refg = OG.gptq_matmul(xg.float().cpu().numpy().reshape(6, k), OG.pack_gptq(q), sc.float().cpu().numpy(), 128) + bias.half().float().cpu().numpy()
    assert yg.shape == (2, 3, n) and rel_fro(yg.float().cpu().numpy().reshape(6, n), refg) < 2e-3
