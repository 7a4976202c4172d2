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
@pytest.mark.parametrize("m,k,n,g", [(32, 512, 256, 128), (7, 1024, 384, 64), (64, 256, 128, -1), (32, 4096, 1024, 128), (1, 2048, 512, 128)])
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
def test_marlin_unsupported_paths_error_like_reference():
    x = torch.zeros((2, 256), dtype=torch.float16, device="cuda")
    qw = torch.zeros((16, 128), dtype=torch.int32, device="cuda")
    s = torch.ones((2, 64), dtype=torch.float16, device="cuda")
    with pytest.raises(pkg.BackendError, match="workspace is required"):
        pkg.gptq_matmul(x, qw, s, None, None, None, 4, 128)
    with pytest.raises(pkg.BackendError):
        pkg.gptq_matmul(x.float(), qw, s, None, None, torch.zeros(64, dtype=torch.int32, device="cuda"), 4, 128)
    with pytest.raises(pkg.BackendError, match="group size"):
        pkg.gptq_matmul(x, qw, s, None, None, torch.zeros(64, dtype=torch.int32, device="cuda"), 4, 32)
