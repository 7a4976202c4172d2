"""GPU parity: paged attention decode (K1) and chunked prefill (K2) through the C ABI vs the fp64
oracle.  Tolerance: <= 2e-3 abs on bf16 I/O for O(1) values = half a bf16 ulp at 1.0
plus fp32-accumulation noise; we assert max-abs <= 2e-2 * max|out| elementwise and a tight
rel-Frobenius bound (the output itself is rounded to bf16: rel step 2^-8 = 3.9e-3)."""
import numpy as np
import pytest
import torch

import candle_vllm_b200 as pkg
from oracle import attention as OA
from oracle import cache_ops as OC
from tests.gpu_util import DEV, rel_fro, to_bf16_t

pytestmark = pytest.mark.gpu


def _mk(rng, B, H, kvh, hd, bs, nb, ctx, dtype=torch.bfloat16, fp8=False, layout="flash", max_blocks=None):
    nblk = [-(-int(c) // bs) for c in ctx]
    width = max_blocks or max(nblk)
    perm = rng.permutation(nb)
    bt = np.zeros((B, width), np.int32)
    o = 0
    for b in range(B):
        bt[b, :nblk[b]] = perm[o:o + nblk[b]]; o += nblk[b]
    kf = rng.standard_normal((nb, bs, kvh, hd)).astype(np.float32)
    vf = rng.standard_normal((nb, bs, kvh, hd)).astype(np.float32)
    q = torch.from_numpy(rng.standard_normal((B, H, hd)).astype(np.float32)).to(DEV).to(dtype)
    if fp8:
        kc = torch.from_numpy(OC.f32_to_e4m3(kf)).to(DEV); vc = torch.from_numpy(OC.f32_to_e4m3(vf)).to(DEV)
        kn, vn = kc.cpu().numpy(), vc.cpu().numpy()
    else:
        kc = torch.from_numpy(kf).to(DEV).to(dtype); vc = torch.from_numpy(vf).to(DEV).to(dtype)
        kn, vn = kc.float().cpu().numpy(), vc.float().cpu().numpy()
    if layout == "paged":
        x = 16 // kc.element_size()
        kc = kc.view(nb, bs, kvh, hd // x, x).permute(0, 2, 3, 1, 4).contiguous()
        vc = vc.permute(0, 2, 3, 1).contiguous()
    return q, kc, vc, kn, vn, bt


def _check(out, ref, tol_fro=4e-3):
    out = out.float().cpu().numpy()
    assert np.isfinite(out).all()
    assert rel_fro(out, ref) < tol_fro, rel_fro(out, ref)
    assert np.abs(out - ref).max() <= 2e-2 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("B,H,kvh,hd,bs,ctx", [
    (4, 32, 8, 128, 64, [1, 64, 65, 300]),                   # Llama-3-8B head config, ragged
    (3, 8, 8, 128, 64, [129, 5, 1000]),                      # MHA (group 1)
    (2, 16, 2, 128, 64, [700, 64]),                          # group 8
    (5, 4, 1, 64, 16, [1, 2, 17, 33, 100]),                  # small head_dim / block
    (2, 6, 2, 96, 32, [50, 77]),                             # head_dim not a power of two
    (32, 32, 8, 128, 64, [4096 + 7 * i for i in range(32)]), # metric shape: batch 32, ctx ~4k
])
def test_decode_matches_oracle(B, H, kvh, hd, bs, ctx):
    rng = np.random.default_rng(hash((B, H, kvh, hd)) & 0xffff)
    nb = sum(-(-c // bs) for c in ctx) + 3
    q, kc, vc, kn, vn, bt = _mk(rng, B, H, kvh, hd, bs, nb, ctx)
    attn = pkg.PagedAttention(H, hd, hd ** -0.5, kvh)
    meta = pkg.InputMetadata(is_prefill=False, slot_mapping=torch.zeros(0, dtype=torch.int64, device=DEV),
                             block_tables=torch.from_numpy(bt).to(DEV), context_lens=torch.tensor(ctx, dtype=torch.int32, device=DEV))
    out = attn.forward(q, None, None, None, kc, vc, meta)
    ref = OA.paged_attention_decode(q.float().cpu().numpy(), kn, vn, bt, ctx, hd ** -0.5)
    _check(out, ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,H,kvh,ctx", [
    (4, 32, 8, [1, 64, 65, 300]),                             # ragged, single-tile items (header ring of the 3-stage pipeline)
    (3, 8, 8, [129, 5, 1000]),                                # group 1
    (2, 16, 2, [700, 64]),                                    # group 8
    (6, 8, 2, [31, 32, 33, 95, 96, 97]),                      # tile boundaries: masked V rows in the f16 staging tile
    (32, 32, 8, [4096 + 7 * i for i in range(32)]),           # metric shape (config 3: FP8 KV), batch 32, ctx ~4k
])
def test_decode_fp8_kv_on_the_tma_path_matches_oracle(dtype, B, H, kvh, ctx):
    """FP8 (e4m3, scale 1.0) KV cache, flash layout, block 64, head 128: the TMA split-KV kernel with the e4m3 -> f16 staging tile.
    Same tolerance as the 16-bit cache (the expansion is exact; q is bf16 -> f16, P is f16)."""
    rng = np.random.default_rng(hash((B, H, kvh, len(ctx))) & 0xffff)
    nb = sum(-(-c // 64) for c in ctx) + 3
    q, kc, vc, kn, vn, bt = _mk(rng, B, H, kvh, 128, 64, nb, ctx, dtype=dtype, fp8=True)
    attn = pkg.PagedAttention(H, 128, 128 ** -0.5, kvh, fp8_kvcache=True)
    meta = pkg.InputMetadata(is_prefill=False, slot_mapping=torch.zeros(0, dtype=torch.int64, device=DEV),
                             block_tables=torch.from_numpy(bt).to(DEV), context_lens=torch.tensor(ctx, dtype=torch.int32, device=DEV))
    out = attn.forward(q, None, None, None, kc, vc, meta)
    ref = OA.paged_attention_decode(q.float().cpu().numpy(), kn, vn, bt, ctx, 128 ** -0.5, fp8=True)
    _check(out, ref)
    out2 = attn.forward(q, None, None, None, kc, vc, meta)        # work queue re-zeroed by the merge kernel: repeatable
    assert torch.equal(out, out2)


def test_decode_padded_tables_like_graph_replay():
    # graph.rs:732-738: block tables padded to max_num_blocks; kernel must only read context_lens
    rng = np.random.default_rng(1)
    ctx = [100, 700, 64]
    q, kc, vc, kn, vn, bt = _mk(rng, 3, 32, 8, 128, 64, 40, ctx, max_blocks=96)
    attn = pkg.PagedAttention(32, 128, 128 ** -0.5, 8)
    meta = pkg.InputMetadata(False, torch.zeros(0, dtype=torch.int64, device=DEV), torch.from_numpy(bt).to(DEV),
                             torch.tensor(ctx, dtype=torch.int32, device=DEV))
    out = attn.forward(q, None, None, None, kc, vc, meta)
    ref = OA.paged_attention_decode(q.float().cpu().numpy(), kn, vn, bt, ctx, 128 ** -0.5)
    _check(out, ref)
    # f16 hand-off output (bf16-rounded values stored as fp16) used by the fused decode layer
    out16 = attn.forward(q, None, None, None, kc, vc, meta, out_dtype=torch.float16)
    assert out16.dtype == torch.float16
    # bf16 -> fp16 is exact except below the fp16 normal range (2^-14): half a subnormal quantum
    assert torch.allclose(out16.float(), out.float(), rtol=0, atol=3.1e-8)


@pytest.mark.parametrize("kw", [dict(fp8=True), dict(layout="paged"), dict(fp8=True, layout="paged"),
                                dict(dtype=torch.float16), dict(softcap=30.0), dict(window=40)])
def test_decode_variants(kw):
    rng = np.random.default_rng(2)
    ctx = [33, 150, 64]
    fp8, layout = kw.get("fp8", False), kw.get("layout", "flash")
    dtype = kw.get("dtype", torch.bfloat16)
    q, kc, vc, kn, vn, bt = _mk(rng, 3, 8, 2, 128, 16, 30, ctx, dtype=dtype, fp8=fp8, layout=layout)
    attn = pkg.PagedAttention(8, 128, 128 ** -0.5, 2, sliding_window=kw.get("window"), fp8_kvcache=fp8)
    meta = pkg.InputMetadata(False, torch.zeros(0, dtype=torch.int64, device=DEV), torch.from_numpy(bt).to(DEV),
                             torch.tensor(ctx, dtype=torch.int32, device=DEV))
    out = attn.forward(q, None, None, None, kc, vc, meta, softcapping=kw.get("softcap"))
    ref = OA.paged_attention_decode(q.float().cpu().numpy(), kn, vn, bt, ctx, 128 ** -0.5, softcap=kw.get("softcap"),
                                    sliding_window=kw.get("window"), fp8=fp8)
    _check(out, ref)


def test_forward_writes_cache_then_attends():
    # PagedAttention.forward = reshape_and_cache (K3) then attention (K1): the new token must be visible
    rng = np.random.default_rng(3)
    B, H, kvh, hd, bs, nb = 4, 8, 2, 128, 64, 12
    lens = [1, 64, 65, 130]
    tables = [[3], [5], [7, 1], [9, 0, 2]]
    prep = pkg.prepare_decode(lens, [0] * B, tables, bs)
    kc = to_bf16_t(rng.standard_normal((nb, bs, kvh, hd))); vc = to_bf16_t(rng.standard_normal((nb, bs, kvh, hd)))
    kn, vn = kc.float().cpu().numpy(), vc.float().cpu().numpy()
    q = to_bf16_t(rng.standard_normal((B, H, hd))); k = to_bf16_t(rng.standard_normal((B, kvh, hd))); v = to_bf16_t(rng.standard_normal((B, kvh, hd)))
    _, _, meta = pkg.inputs.to_device(prep)
    out = pkg.PagedAttention(H, hd, hd ** -0.5, kvh).forward(q, k, v, None, kc, vc, meta)
    OC.reshape_and_cache_flash(k.float().cpu().numpy(), v.float().cpu().numpy(), kn, vn, prep["slot_mapping"])
    assert np.array_equal(kc.float().cpu().numpy(), kn)                     # cache write bit-exact
    ref = OA.paged_attention_decode(q.float().cpu().numpy(), kn, vn, prep["block_tables"], prep["context_lens"], hd ** -0.5)
    _check(out, ref)


def test_prefill_chunked_matches_oracle():
    rng = np.random.default_rng(4)
    H, kvh, hd, bs, nb = 8, 2, 128, 16, 40
    prompts = [list(range(70)), list(range(33)), list(range(200))]
    tables = [[1, 2, 3, 4, 5], [9, 8, 7], list(range(10, 23))]
    cached = [32, 0, 128]                                    # seq 0 and 2 continue from earlier chunks
    prep = pkg.prepare_prompt(prompts, tables, bs, cached, chunk_size=64)
    T = len(prep["tokens"])
    kc = to_bf16_t(rng.standard_normal((nb, bs, kvh, hd))); vc = to_bf16_t(rng.standard_normal((nb, bs, kvh, hd)))
    q = to_bf16_t(rng.standard_normal((T, H, hd))); k = to_bf16_t(rng.standard_normal((T, kvh, hd))); v = to_bf16_t(rng.standard_normal((T, kvh, hd)))
    kn, vn = kc.float().cpu().numpy(), vc.float().cpu().numpy()
    _, _, meta = pkg.inputs.to_device(prep)
    out = pkg.PagedAttention(H, hd, hd ** -0.5, kvh).forward(q, k, v, None, kc, vc, meta)
    OC.reshape_and_cache_flash(k.float().cpu().numpy(), v.float().cpu().numpy(), kn, vn, prep["slot_mapping"])
    ref = OA.paged_attention_prefill(q.float().cpu().numpy(), kn, vn, prep["block_tables"], prep["cu_seqlens_q"],
                                     prep["cu_seqlens_k"], hd ** -0.5)
    _check(out, ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("H,kvh,qlens,cached,window", [
    (8, 2, [70, 33, 200], [32, 0, 128], None),       # ragged chunks continuing cached prefixes (bottom-right causal mask)
    (4, 4, [64, 1, 129], [0, 500, 63], None),         # exact tile, single row deep in a context, one row past two tiles
    (8, 1, [300], [100], 96),                         # sliding window: pages entirely before the window are skipped
    (32, 8, [1024], [1024], None),                    # Llama-3-8B head config, 1 K chunk on a 1 K cached prefix
])
def test_prefill_on_tensor_cores_matches_oracle(dtype, H, kvh, qlens, cached, window):
    """Chunked prefill over the paged cache on the tensor-core kernel (block 64, head 128; csrc/attention_prefill.cu): every key and
    value -- cached prefix and this chunk -- is read from the cache by TMA, one 64-token page per 64-row query tile.  Tolerance as
    for decode (16-bit P and output).  The generic kernel (B200_PREFILL_GENERIC=1) must agree to the same tolerance."""
    rng = np.random.default_rng(len(qlens) * 13 + H)
    bs, hd = 64, 128
    klens = [q + c for q, c in zip(qlens, cached)]
    nblk = [-(-k // bs) for k in klens]
    nb = sum(nblk) + 2
    perm = rng.permutation(nb)
    tables, o = [], 0
    for n in nblk:
        tables.append([int(x) for x in perm[o:o + n]]); o += n
    prompts = [list(range(k)) for k in klens]
    prep = pkg.prepare_prompt(prompts, tables, bs, cached, chunk_size=max(qlens))
    T = len(prep["tokens"])
    assert T == sum(qlens)
    mk = lambda *shape: torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(DEV).to(dtype)
    kc, vc = mk(nb, bs, kvh, hd), mk(nb, bs, kvh, hd)
    kc[perm[-1]] = float("nan"); vc[perm[-1]] = float("nan")        # an unowned block: must never be touched
    q, k, v = mk(T, H, hd), mk(T, kvh, hd), mk(T, kvh, hd)
    kn, vn = kc.float().cpu().numpy(), vc.float().cpu().numpy()
    _, _, meta = pkg.inputs.to_device(prep)
    attn = pkg.PagedAttention(H, hd, hd ** -0.5, kvh, sliding_window=window)
    out = attn.forward(q, k, v, None, kc, vc, meta)
    OC.reshape_and_cache_flash(k.float().cpu().numpy(), v.float().cpu().numpy(), kn, vn, prep["slot_mapping"])
    ref = OA.paged_attention_prefill(q.float().cpu().numpy(), kn, vn, prep["block_tables"], prep["cu_seqlens_q"], prep["cu_seqlens_k"],
                                     hd ** -0.5, sliding_window=window)
    _check(out, ref)


def test_decode_from_flashinfer_csr_metadata_equals_block_tables():
    """The reference's default build describes the pages as CSR (indptr / indices / last_len, inputs.rs:477-506): expanded on the device by
    flashinfer_csr_to_paged, decode must give the same bits as with padded block tables."""
    rng = np.random.default_rng(21)
    B, H, kvh, hd, bs = 5, 8, 2, 128, 64
    ctx = [1, 64, 65, 300, 128]
    nb = sum(-(-c // bs) for c in ctx) + 3
    q, kc, vc, kn, vn, bt = _mk(rng, B, H, kvh, hd, bs, nb, ctx)
    attn = pkg.PagedAttention(H, hd, hd ** -0.5, kvh)
    zero = torch.zeros(0, dtype=torch.int64, device=DEV)
    ref = attn.forward(q, None, None, None, kc, vc, pkg.InputMetadata(False, zero, torch.from_numpy(bt).to(DEV), torch.tensor(ctx, dtype=torch.int32, device=DEV)))
    csr = pkg.flashinfer_csr(ctx, [list(r) for r in bt], bs)
    t = lambda a: torch.from_numpy(a.astype(np.int32)).to(DEV)
    fm = pkg.FlashInferMetadata(t(csr["indptr"]), t(csr["indices"]), t(csr["last_len"]), max_blocks_per_seq=bt.shape[1])
    tables, lens = fm.to_paged(bs)
    assert lens.cpu().numpy().tolist() == ctx
    for b in range(B):
        n = -(-ctx[b] // bs)
        assert tables[b, :n].cpu().numpy().tolist() == bt[b, :n].tolist() and (tables[b, n:] == 0).all()
    out = attn.forward(q, None, None, None, kc, vc, pkg.InputMetadata(False, zero, flashinfer_metadata=fm))
    assert torch.equal(out, ref)


def test_attention_argument_errors():
    attn = pkg.PagedAttention(8, 128, 0.1, 2)
    q = torch.zeros(2, 8, 128, dtype=torch.float32, device=DEV)
    kc = torch.zeros(4, 16, 2, 128, dtype=torch.bfloat16, device=DEV)
    meta = pkg.InputMetadata(False, torch.zeros(0, dtype=torch.int64, device=DEV), torch.zeros(2, 2, dtype=torch.int32, device=DEV),
                             torch.ones(2, dtype=torch.int32, device=DEV))
    with pytest.raises(pkg.BackendError):
        attn.forward(q, None, None, None, kc, kc, meta)                    # f32 query unsupported
    with pytest.raises(pkg.BackendError):
        attn.forward(q.bfloat16()[:, :4], None, None, None, kc, kc, meta)   # head count mismatch
    with pytest.raises(pkg.BackendError):
        pkg.PagedAttention(6, 128, 0.1, 4)


def test_decode_repeatable_on_one_workspace():
    """The work-queue head lives in the caller's workspace and is re-armed by the merge kernel: 200 back-to-back launches
    on ONE workspace, with changing context lengths (1 .. many chunks, one empty context), must all match the first-launch
    result bit for bit (the split merge folds chunks in a fixed order) and the oracle."""
    rng = np.random.default_rng(21)
    B, H, kvh, hd, bs = 16, 32, 8, 128, 64
    ctxs = [[1 + (37 * b * (r + 1)) % 2900 for b in range(B)] for r in range(4)]
    ctxs[1][3] = 0                                            # an empty context attends to nothing -> zeros
    nb = 16 * 46 + 3
    q, kc, vc, kn, vn, bt = _mk(rng, B, H, kvh, hd, bs, nb, [2900] * B)
    attn = pkg.PagedAttention(H, hd, hd ** -0.5, kvh)
    first = {}
    for it in range(200):
        r = it % 4
        meta = pkg.InputMetadata(False, torch.zeros(0, dtype=torch.int64, device=DEV), torch.from_numpy(bt).to(DEV),
                                 torch.tensor(ctxs[r], dtype=torch.int32, device=DEV))
        out = attn.forward(q, None, None, None, kc, vc, meta)
        if r not in first:
            first[r] = out.clone()
            ref = OA.paged_attention_decode(q.float().cpu().numpy(), kn, vn, bt, ctxs[r], hd ** -0.5)
            _check(out, ref)
            if r == 1:
                assert float(out[3].abs().max()) == 0.0
        else:
            assert torch.equal(out, first[r]), f"launch {it} differs from launch {r}"
