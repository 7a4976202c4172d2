"""GPTQ / Marlin int4: oracle pinned to the reference's permutation tables (CPU) and GPU parity of
gptq_repack + marlin_4bit_{f16,bf16} through the C ABI against the fp64 oracle."""
import json
import os

import numpy as np
import pytest
import torch

import candle_vllm_b200 as pkg
from oracle import gptq as OG

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_scale_perms_match_reference_golden():
    g = json.load(open(os.path.join(GOLD, "marlin_perms.json")))     # produced by executing the reference's python
    sp, sps = OG.get_scale_perms()
    assert sp == g["scale_perm"] and sps == g["scale_perm_single"]
    from candle_vllm_b200 import gptq
    assert list(gptq.get_scale_perms()[0]) == sp and list(gptq.get_scale_perms()[1]) == sps


def test_pack_unpack_and_permute_roundtrip():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 16, (256, 64), dtype=np.uint8)
    qw = OG.pack_gptq(q)
    assert qw.shape == (32, 64) and np.array_equal(OG.unpack_gptq(qw), q)
    assert (qw[0, 0] & 0xF) == q[0, 0] and ((qw[0, 0] >> 28) & 0xF) == q[7, 0]       # LSB-first along K
    s = rng.standard_normal((2, 128)).astype(np.float32)
    p = OG.marlin_permute_scales(s, 256, 128, 128)
    sp, _ = OG.get_scale_perms()
    assert np.array_equal(p.reshape(-1, 64), s.reshape(-1, 64)[:, sp])
    t = torch.from_numpy(s)
    assert np.array_equal(pkg.marlin_permute_scales(t, 256, 128, 128).numpy(), p)
    assert np.array_equal(pkg.marlin_permute_scales(t[:1], 256, 128, -1).numpy(), OG.marlin_permute_scales(s[:1], 256, 128, -1))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("m,k,n,g", [(32, 512, 256, 128), (7, 1024, 384, 64), (64, 256, 128, -1), (32, 4096, 1024, 128), (1, 2048, 512, 128),
                                     (150, 512, 256, 128)])         # > 64 rows: 64 per tensor-core pass
def test_marlin_matmul_matches_oracle(dtype, m, k, n, g):
    rng = np.random.default_rng(m * 7 + n)
    q = rng.integers(0, 16, (k, n), dtype=np.uint8)
    ng = 1 if g == -1 else k // g
    scales = (rng.uniform(0.005, 0.02, (ng, n))).astype(np.float32)
    x = rng.standard_normal((m, k)).astype(np.float32)
    st = torch.from_numpy(scales).cuda().to(dtype)
    xt = torch.from_numpy(x).cuda().to(dtype)
    qw = torch.from_numpy(OG.pack_gptq(q).view(np.int32)).cuda()
    w_m = pkg.marlin_weight_repack(qw, 4, False)
    assert w_m.shape == (k // 16, 2 * n)
    s_m = pkg.marlin_permute_scales(st, k, n, g)
    ws = torch.zeros(n, dtype=torch.int32, device="cuda")
    y = pkg.gptq_matmul(xt, w_m, s_m, None, None, ws, 4, g)
    assert y.dtype == dtype and y.shape == (m, n)
    ref = OG.gptq_matmul(xt.float().cpu().numpy(), OG.pack_gptq(q), st.float().cpu().numpy(), g)
    rel = np.linalg.norm(y.float().cpu().numpy() - ref) / np.linalg.norm(ref)
    # 16-bit output rounding dominates: 2^-9 (bf16) / 2^-12 (f16) relative per element
    assert rel < (4e-3 if dtype == torch.bfloat16 else 1e-3), rel


@pytest.mark.gpu
def test_marlin_matmul_with_checkpoint_g_idx_and_qzeros_like_the_reference_call_site():
    """The reference hands the checkpoint's g_idx (trivial k // group_size when desc_act = false, linear.rs:298-337) and qzeros to
    marlin_4bit_* for every quant_method == "gptq" layer (gptq.rs:27-35,139-152): the result must be the same as without them."""
    rng = np.random.default_rng(11)
    m, k, n, g = 9, 1024, 256, 128
    q = rng.integers(0, 16, (k, n), dtype=np.uint8)
    st = torch.from_numpy(rng.uniform(0.005, 0.02, (k // g, n)).astype(np.float32)).cuda().half()
    xt = torch.from_numpy(rng.standard_normal((m, k)).astype(np.float32)).cuda().half()
    w_m = pkg.marlin_weight_repack(torch.from_numpy(OG.pack_gptq(q).view(np.int32)).cuda(), 4, False)
    s_m = pkg.marlin_permute_scales(st, k, n, g)
    ws = torch.zeros(n, dtype=torch.int32, device="cuda")
    g_idx = torch.arange(k, dtype=torch.int32, device="cuda") // g
    qzeros = torch.full((k // g, n // 8), 0x77777777, dtype=torch.int32, device="cuda")      # GPTQ v1 symmetric: zero - 1 = 7
    y0 = pkg.gptq_matmul(xt, w_m, s_m, None, None, ws, 4, g)
    y1 = pkg.gptq_matmul(xt, w_m, s_m, qzeros, g_idx, ws, 4, g)
    assert torch.equal(y0, y1)
    ref = OG.gptq_matmul(xt.float().cpu().numpy(), OG.pack_gptq(q), st.float().cpu().numpy(), g)
    assert np.linalg.norm(y1.float().cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-3


def test_marlin_pack_roundtrip_on_the_oracle():
    rng = np.random.default_rng(2)
    q = rng.integers(0, 16, (128, 192), dtype=np.uint8)
    B = OG.pack_marlin(q)
    assert B.shape == (128 // 16, 2 * 192) and np.array_equal(OG.unpack_marlin(B, 128, 192), q)
    assert sorted(OG.marlin_weight_perm().tolist()) == list(range(1024))


@pytest.mark.gpu
@pytest.mark.parametrize("k,n,g", [(256, 128, 128), (1024, 384, 64), (512, 64, -1)])
def test_marlin_format_checkpoint_is_accepted_after_one_repack(k, n, g):
    """checkpoint_format == "marlin" (linear.rs:219-251): `B` in the Marlin project's tile order + `s` already permuted.  One load-time
    marlin_checkpoint_repack turns B into exactly what marlin_weight_repack makes from the GPTQ tensors; the GEMM then matches the oracle."""
    from candle_vllm_b200 import gptq
    rng = np.random.default_rng(k + n)
    q = rng.integers(0, 16, (k, n), dtype=np.uint8)
    b = torch.from_numpy(OG.pack_marlin(q).view(np.int32)).cuda()
    w_from_marlin = gptq.marlin_checkpoint_repack(b, k, n)
    w_from_gptq = pkg.marlin_weight_repack(torch.from_numpy(OG.pack_gptq(q).view(np.int32)).cuda(), 4, False)
    assert torch.equal(w_from_marlin, w_from_gptq)
    ng = 1 if g == -1 else k // g
    scales = rng.uniform(0.005, 0.02, (ng, n)).astype(np.float32)
    st = torch.from_numpy(scales).cuda().half()
    x = torch.from_numpy(rng.standard_normal((7, k)).astype(np.float32)).cuda().half()
    y = pkg.gptq_matmul(x, w_from_marlin, pkg.marlin_permute_scales(st, k, n, g), None, None, torch.zeros(n, dtype=torch.int32, device="cuda"), 4, g)
    ref = OG.gptq_matmul(x.float().cpu().numpy(), OG.pack_gptq(q), st.float().cpu().numpy(), g)
    assert np.linalg.norm(y.float().cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-3


def test_awq_zero_point_layout_matches_reference_converter():
    """oracle restatement of examples/convert_awq_marlin.py:75-113 against vectors produced by EXECUTING that script
    (tests/golden/make_golden.py: awq_zero_points), plus pack / unpack round trips."""
    g = json.load(open(os.path.join(GOLD, "awq_zero_points.json")))
    zp = np.asarray(g["zp"], np.uint8)
    packed = np.asarray(g["awq_packed"], np.uint32)
    assert np.array_equal(OG.pack_awq(zp), packed) and np.array_equal(OG.unpack_awq(packed), zp)
    assert np.array_equal(OG.awq_to_marlin_zero_points(packed), np.asarray(g["marlin_zp"], np.uint32))
    assert np.array_equal(OG.unpack_cols(OG.pack_cols(zp)), zp)
    assert (packed[0, 0] >> 4) & 0xF == zp[0, 2]                   # nibble 1 of an AWQ word holds column 2


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("m,k,n,g", [(32, 512, 256, 128), (5, 1024, 128, 64), (64, 256, 192, -1), (16, 4096, 512, 128)])
def test_awq_marlin_matmul_matches_oracle(dtype, m, k, n, g):
    """awq_repack + marlin_awq_4bit_{f16,bf16}: AWQ checkpoint tensors in, zero points converted like the reference's offline
    tool, against the fp64 oracle (q - z) * s."""
    rng = np.random.default_rng(m + k + n)
    q = rng.integers(0, 16, (k, n), dtype=np.uint8)
    ng = 1 if g == -1 else k // g
    z = rng.integers(0, 16, (ng, n), dtype=np.uint8)
    scales = rng.uniform(0.005, 0.02, (ng, n)).astype(np.float32)
    x = rng.standard_normal((m, k)).astype(np.float32)
    st = torch.from_numpy(scales).cuda().to(dtype)
    xt = torch.from_numpy(x).cuda().to(dtype)
    qw_awq, qz_awq = OG.pack_awq(q), OG.pack_awq(z)
    w_m = pkg.marlin_weight_repack(torch.from_numpy(qw_awq.view(np.int32)).cuda(), 4, True)
    assert w_m.shape == (k // 16, 2 * n)
    s_m = pkg.marlin_permute_scales(st, k, n, g)
    z_m = torch.from_numpy(OG.awq_to_marlin_zero_points(qz_awq).view(np.int32)).cuda()
    ws = torch.zeros(n, dtype=torch.int32, device="cuda")
    y = pkg.gptq_matmul(xt, w_m, s_m, z_m, None, ws, 4, g, is_awq=True)
    ref = xt.float().cpu().numpy().astype(np.float64) @ OG.dequant_awq(qw_awq, qz_awq, st.float().cpu().numpy(), g).T
    rel = np.linalg.norm(y.float().cpu().numpy() - ref) / np.linalg.norm(ref)
    assert y.dtype == dtype and y.shape == (m, n)
    assert rel < (4e-3 if dtype == torch.bfloat16 else 1e-3), rel


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("m,k,n,g", [(3, 256, 128, 64), (17, 512, 136, 128), (1, 128, 64, 32)])
def test_conventional_gptq_act_order_matmul_matches_oracle(bits, m, k, n, g):
    """gemm_half_q_half_alt (no marlin workspace): asymmetric zero points stored minus one, act-order g_idx (a random permutation of
    the rows' groups), 4 and 8 bit."""
    rng = np.random.default_rng(bits + m + k)
    pack = 32 // bits
    q = rng.integers(0, 1 << bits, (k, n), dtype=np.uint32)
    ng = k // g
    z = rng.integers(0, (1 << bits) - 1, (ng, n), dtype=np.uint32)              # stored value; effective zero = z + 1
    scales = rng.uniform(0.005, 0.02, (ng, n)).astype(np.float16)
    g_idx = rng.permutation(np.repeat(np.arange(ng), g)).astype(np.int32)       # desc_act: rows of a group are scattered
    qw = np.zeros((k // pack, n), np.uint32)
    for i in range(pack):
        qw |= q.reshape(k // pack, pack, n)[:, i, :] << np.uint32(bits * i)
    qz = np.zeros((ng, n // pack), np.uint32)
    for i in range(pack):
        qz |= z.reshape(ng, n // pack, pack)[:, :, i] << np.uint32(bits * i)
    x = rng.standard_normal((m, k)).astype(np.float16)
    y = pkg.gptq_matmul(torch.from_numpy(x).cuda(), torch.from_numpy(qw.view(np.int32)).cuda(), torch.from_numpy(scales).cuda(),
                        torch.from_numpy(qz.view(np.int32)).cuda(), torch.from_numpy(g_idx).cuda(), None, bits, g)
    w = (q.astype(np.float64) - (z.astype(np.float64)[g_idx] + 1.0)) * scales.astype(np.float64)[g_idx]         # [K, N]
    ref = x.astype(np.float64) @ w
    if bits == 4:
        assert np.allclose(OG.dequant_gptq_alt(qw, qz, scales, g_idx).T, w)
    rel = np.linalg.norm(y.float().cpu().numpy() - ref) / np.linalg.norm(ref)
    assert y.dtype == torch.float16 and rel < 1e-3, rel


@pytest.mark.gpu
def test_marlin_unsupported_paths_error_like_reference():
    x = torch.zeros((2, 256), dtype=torch.float16, device="cuda")
    qw = torch.zeros((16, 128), dtype=torch.int32, device="cuda")
    s = torch.ones((2, 64), dtype=torch.float16, device="cuda")
    with pytest.raises(pkg.BackendError, match="workspace is required"):
        pkg.gptq_matmul(x, qw, s, None, None, None, 4, 128)                       # neither marlin nor conventional inputs
    with pytest.raises(pkg.BackendError, match="only supported for f16 non-marlin"):
        pkg.gptq_matmul(x.bfloat16(), qw, s.bfloat16(), qw, qw, None, 4, 128)
    with pytest.raises(pkg.BackendError):
        pkg.gptq_matmul(x.float(), qw, s, None, None, torch.zeros(64, dtype=torch.int32, device="cuda"), 4, 128)
    with pytest.raises(pkg.BackendError, match="group size"):
        pkg.gptq_matmul(x, qw, s, None, None, torch.zeros(64, dtype=torch.int32, device="cuda"), 4, 32)
    with pytest.raises(pkg.BackendError, match="needs qzeros"):
        pkg.gptq_matmul(x, qw, s, None, None, torch.zeros(64, dtype=torch.int32, device="cuda"), 4, 128, is_awq=True)
