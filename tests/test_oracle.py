"""CPU tests: the oracle against the golden vectors and against itself (numpy vs C restatement)."""
import json
import os

import numpy as np
import pytest

from oracle import attention as A
from oracle import cache_ops as C
from oracle import cpu_ref, ggml_quants as G, llama as LL

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "ggml_dequant.npz"))


@pytest.mark.parametrize("name,fn", [("q4k", G.dequantize_q4k), ("q6k", G.dequantize_q6k), ("q8_0", G.dequantize_q8_0)])
def test_dequant_matches_gguf_golden(gold, name, fn):
    # bit-exact: same fp32 operation order as ggml (d*sc first, then *q, then -dmin*m)
    got = fn(gold[f"{name}_blocks"])
    assert np.array_equal(got, gold[f"{name}_deq"])


def test_q8_0_quantize_matches_gguf_golden(gold):
    got = G.quantize_q8_0(gold["q8_0_x"])
    assert np.array_equal(got.reshape(-1), gold["q8_0_quant"].reshape(-1))


def test_dequant_matches_installed_gguf_package():
    gguf = pytest.importorskip("gguf")
    rng = np.random.default_rng(3)
    for t, gt in ((G.GGML_TYPE_Q4_K, gguf.GGMLQuantizationType.Q4_K), (G.GGML_TYPE_Q6_K, gguf.GGMLQuantizationType.Q6_K)):
        w = G.random_weight(rng, t, 16, 1024)
        assert np.array_equal(G.dequantize_weight(w, t, 16, 1024), gguf.quants.dequantize(w, gt))


def test_slot_mapping_known_answers():
    cases = json.load(open(os.path.join(GOLD, "slot_mapping.json")))
    for c in cases:
        assert C.decode_slot(c["table"], c["position"], c["block_size"]) == c["slot"]
        assert C.used_blocks_for_len(c["seq_len"], c["block_size"], len(c["table"])) == c["used"]
    assert C.used_blocks_for_len(0, 64, 5) == 0
    with pytest.raises(ValueError):
        C.decode_slot([1, 2], 128, 64)            # "Block table is too small"


def test_prepare_decode_pads_and_trims():
    p = C.prepare_decode([65, 3], [[7, 3, 99], [4]], 64)
    assert p["block_tables"].tolist() == [[7, 3], [4, 0]]     # trimmed to used blocks, 0 padded
    assert p["slot_mapping"].tolist() == [3 * 64 + 0, 4 * 64 + 2]
    assert p["context_lens"].tolist() == [65, 3] and p["max_context_len"] == 65


def test_e4m3_matches_torch():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(50000).astype(np.float32) * s for s in (1e-3, 0.02, 1, 30, 100)])
    x = x[np.abs(x) <= 448]
    x = np.concatenate([x, np.array([0, -0.0, 448, -448, 2 ** -9, 2 ** -10, 1.5 * 2 ** -10, 2 ** -6], np.float32)])
    ref = torch.from_numpy(x).to(torch.float8_e4m3fn).view(torch.uint8).numpy()
    assert np.array_equal(C.f32_to_e4m3(x), ref)
    # saturating beyond the finite range
    assert C.f32_to_e4m3(np.array([1e9, -1e9, 465], np.float32)).tolist() == [0x7E, 0xFE, 0x7E]
    allb = np.arange(256, dtype=np.uint8)
    d = C.e4m3_to_f32(allb)
    r = torch.from_numpy(allb).view(torch.float8_e4m3fn).float().numpy()
    ok = ~np.isnan(r)
    assert np.array_equal(d[ok], r[ok]) and np.isnan(d[~ok]).all()


def test_copy_and_swap_blocks_semantics():
    rng = np.random.default_rng(0)
    kc = [rng.standard_normal((6, 4, 2, 8)).astype(np.float32) for _ in range(2)]
    vc = [rng.standard_normal((6, 4, 2, 8)).astype(np.float32) for _ in range(2)]
    k0 = [k.copy() for k in kc]
    C.copy_blocks(kc, vc, {1: [3, 4], 0: [5]})
    for l in range(2):
        assert np.array_equal(kc[l][3], k0[l][1]) and np.array_equal(kc[l][4], k0[l][1]) and np.array_equal(kc[l][5], k0[l][0])
        assert np.array_equal(kc[l][2], k0[l][2])
    dst = np.zeros_like(kc[0])
    C.swap_blocks(kc[0], dst, {0: 2, 3: 1})
    assert np.array_equal(dst[2], kc[0][0]) and np.array_equal(dst[1], kc[0][3]) and not dst[0].any()


def test_reshape_and_cache_layouts_agree():
    rng = np.random.default_rng(0)
    T, kvh, hd, bs, nb = 5, 2, 32, 4, 6
    k = rng.standard_normal((T, kvh, hd)).astype(np.float32)
    v = rng.standard_normal((T, kvh, hd)).astype(np.float32)
    slots = np.array([9, 0, -1, 23, 5])
    kf = np.zeros(C.flash_kv_shape(nb, bs, kvh, hd), np.float32); vf = np.zeros_like(kf)
    C.reshape_and_cache_flash(k, v, kf, vf, slots)
    kp = np.zeros(C.paged_k_shape(nb, bs, kvh, hd, 2), np.float32); vp = np.zeros(C.paged_v_shape(nb, bs, kvh, hd), np.float32)
    C.reshape_and_cache_paged(k, v, kp, vp, slots)
    q = rng.standard_normal((2, 4, hd)).astype(np.float32)
    bt = np.array([[2, 0], [5, 1]]); ctx = np.array([2, 4])
    a = A.paged_attention_decode(q, kf, vf, bt, ctx, 0.17)
    b = A.paged_attention_decode(q, kp, vp, bt, ctx, 0.17, layout="paged")
    assert np.allclose(a, b, atol=1e-6)
    assert not kf.reshape(nb * bs, kvh, hd)[2 * 1 + 0].any()    # pad slot skipped


def test_paged_decode_equals_contiguous_attention():
    rng = np.random.default_rng(5)
    B, H, kvh, hd, bs, nb = 3, 8, 2, 64, 16, 32
    ctx = np.array([1, 33, 100])
    bt = rng.permutation(nb)[:21].reshape(3, 7)
    kc = rng.standard_normal((nb, bs, kvh, hd)).astype(np.float32); vc = rng.standard_normal((nb, bs, kvh, hd)).astype(np.float32)
    q = rng.standard_normal((B, H, hd)).astype(np.float32)
    out = A.paged_attention_decode(q, kc, vc, bt, ctx, hd ** -0.5)
    for b in range(B):
        k = kc[bt[b]].reshape(-1, kvh, hd)[:ctx[b]]; v = vc[bt[b]].reshape(-1, kvh, hd)[:ctx[b]]
        for h in range(H):
            s = (k[:, h // 4] @ q[b, h]) * hd ** -0.5
            p = np.exp(s - s.max()); p /= p.sum()
            assert np.allclose(out[b, h], p @ v[:, h // 4], atol=1e-5)


def test_prefill_last_row_equals_decode():
    rng = np.random.default_rng(6)
    H, kvh, hd, bs, nb = 4, 2, 32, 8, 16
    kc = rng.standard_normal((nb, bs, kvh, hd)).astype(np.float32); vc = rng.standard_normal((nb, bs, kvh, hd)).astype(np.float32)
    bt = rng.permutation(nb)[:8].reshape(2, 4)
    q = rng.standard_normal((5 + 7, H, hd)).astype(np.float32)
    cu_q, cu_k = np.array([0, 5, 12]), np.array([0, 20, 27])      # seq0: 15 cached + 5 new; seq1: 7 new
    out = A.paged_attention_prefill(q, kc, vc, bt, cu_q, cu_k, 0.2)
    dec = A.paged_attention_decode(q[[4, 11]], kc, vc, bt, np.array([20, 7]), 0.2)
    assert np.allclose(out[[4, 11]], dec, atol=1e-6)
    first = A.paged_attention_decode(q[[5]], kc, vc, bt[1:], np.array([1]), 0.2)     # causal: row 0 of seq1 sees 1 key
    assert np.allclose(out[5], first[0], atol=1e-6)


def test_rope_interleaved_vs_neox_and_norm():
    rng = np.random.default_rng(7)
    cos, sin = A.rope_tables(16, 64, 10000.0)
    x = rng.standard_normal((3, 2, 16)).astype(np.float32)
    pos = np.array([0, 5, 63])
    xi = A.apply_rope(x, cos, sin, pos, True)
    xn = A.apply_rope(np.concatenate([x[..., 0::2], x[..., 1::2]], -1), cos, sin, pos, False)
    assert np.allclose(xi[..., 0::2], xn[..., :8], atol=1e-6) and np.allclose(xi[..., 1::2], xn[..., 8:], atol=1e-6)
    assert np.allclose(xi[0], x[0])                                 # position 0 = identity
    assert np.allclose(np.linalg.norm(xi, axis=-1), np.linalg.norm(x, axis=-1), rtol=1e-5)
    w = rng.uniform(0.5, 1.5, 16).astype(np.float32)
    y = A.rms_norm(x, w, 1e-5)
    assert np.allclose(y, x / np.sqrt((x ** 2).mean(-1, keepdims=True) + 1e-5) * w, rtol=1e-5)


def test_c_oracle_matches_numpy_oracle():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 1024)).astype(np.float32)
    for t in (G.GGML_TYPE_Q4_K, G.GGML_TYPE_Q6_K):
        w = G.random_weight(rng, t, 48, 1024)
        a = cpu_ref.qmatmul_q8k(x, w, t, 48, 1024)
        b = G.qmatmul_q8k(x, w, t, 48, 1024)
        assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()
        # and the integer path stays within the 8-bit activation noise floor of the fp32 target
        d = G.qmatmul_dequant(x, w, t, 48, 1024)
        assert np.linalg.norm(b - d) / np.linalg.norm(d) < 2e-2
    B, H, kvh, hd, bs, nb = 3, 8, 2, 128, 16, 40
    kc = LL.bf16_round(rng.standard_normal((nb, bs, kvh, hd)).astype(np.float32))
    vc = LL.bf16_round(rng.standard_normal((nb, bs, kvh, hd)).astype(np.float32))
    q = LL.bf16_round(rng.standard_normal((B, H, hd)).astype(np.float32))
    ctx = np.array([37, 1, 160], np.uint32); bt = rng.permutation(nb)[:30].reshape(3, 10).astype(np.uint32)
    o1 = LL.bf16_round(A.paged_attention_decode(q, kc, vc, bt, ctx, hd ** -0.5))
    o2 = cpu_ref.paged_attention_decode_bf16(q, LL.f32_to_bf16_bits(kc), LL.f32_to_bf16_bits(vc), bt, ctx, hd ** -0.5)
    assert np.abs(o1 - o2).max() < 2e-2      # one bf16 ulp at |x|~2


def test_bf16_round_matches_torch():
    torch = pytest.importorskip("torch")
    x = np.random.default_rng(2).standard_normal(10000).astype(np.float32) * 100
    assert np.array_equal(LL.bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())
    assert np.array_equal(LL.bf16_bits_to_f32(LL.f32_to_bf16_bits(x)), LL.bf16_round(x))


def test_fp_format_decoders_pinned():
    """e4m3fn / e8m0 decoders against torch's dtypes (all 256 codes), e2m1 against the OCP MX v1.0 value table; nibble order of
    the packed fp4 tensors (low nibble = even k, the layout of the reference's `blocks` / `weight_packed` tensors)."""
    import torch
    from oracle import fp_formats as F
    b = np.arange(256, dtype=np.uint8)
    t = torch.from_numpy(b.copy()).view(torch.float8_e4m3fn).float().numpy()
    assert np.array_equal(np.nan_to_num(F.e4m3_to_f32(b), nan=777.0), np.nan_to_num(t, nan=777.0))
    if hasattr(torch, "float8_e8m0fnu"):
        t8 = torch.from_numpy(b.copy()).view(torch.float8_e8m0fnu).float().numpy().astype(np.float64)
        with np.errstate(invalid="ignore"):
            assert np.array_equal(np.nan_to_num(F.e8m0_to_f32(b), nan=-1.0), np.nan_to_num(t8, nan=-1.0))
    assert F.e2m1_to_f32(np.arange(16, dtype=np.uint8)).tolist() == [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0,
                                                                      -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0]
    assert F.unpack_fp4(np.array([[0x21, 0xF7]], np.uint8)).tolist() == [[0.5, 1.0, 6.0, -6.0]]
    # block scale broadcast: scale[i, j] covers rows [i*by, (i+1)*by) x cols [j*bx, (j+1)*bx)
    w = np.full((3, 4), 0x38, np.uint8)                   # 1.0 in e4m3
    s = np.array([[2.0, 3.0], [5.0, 7.0]], np.float32)
    assert F.dequant_fp8_block(w, s, 2, 2).tolist() == [[2, 2, 3, 3], [2, 2, 3, 3], [5, 5, 7, 7]]
