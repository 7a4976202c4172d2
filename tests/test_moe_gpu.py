"""GPU parity of the fused MoE path on GGUF experts: topk_softmax, sort_expert_assignments, the grouped dequant-GEMM
(tcgen05, device-side item list) and the whole FusedMoe block against oracle/moe.py.  Tolerance: the QMatMul contract (fp16 operands,
fp32 accumulate) -> rel-Frobenius 1e-3 per GEMM, 2e-3 for the block (three chained GEMMs)."""
import numpy as np
import pytest
import torch

import candle_vllm_b200 as pkg
from oracle import ggml_quants as G, moe as OM
from tests.gpu_util import DEV, rel_fro

pytestmark = pytest.mark.gpu


def test_topk_softmax_matches_oracle():
    rng = np.random.default_rng(0)
    for T, E, k in [(16, 128, 8), (1, 8, 2), (33, 60, 4), (5, 256, 8)]:
        logits = rng.standard_normal((T, E)).astype(np.float32) * 2
        w, ids = pkg.topk_softmax(torch.from_numpy(logits).to(DEV), k)
        rw, rids = OM.topk_softmax(logits, k)
        assert np.array_equal(ids.cpu().numpy().astype(np.uint32), rids)
        assert np.allclose(w.cpu().numpy(), rw, rtol=1e-5, atol=1e-7)
    # ties: equal logits -> the smaller expert id first, like a stable sort
    _, ids = pkg.topk_softmax(torch.zeros((2, 16), device=DEV), 4)
    assert ids.cpu().numpy().tolist() == [[0, 1, 2, 3]] * 2


def test_sort_expert_assignments_is_a_stable_ascending_sort():
    rng = np.random.default_rng(1)
    for P, E in [(128, 128), (7, 4), (4096, 64)]:
        ids = rng.integers(0, E, P).astype(np.int32)
        e, s = pkg.sort_expert_assignments(torch.from_numpy(ids).to(DEV), E)
        e, s = e.cpu().numpy(), s.cpu().numpy()
        order = np.argsort(ids, kind="stable")
        assert np.array_equal(e, ids[order]) and np.array_equal(s, order)


@pytest.mark.parametrize("ggml_type,E,N,K,T,k", [
    (12, 8, 256, 512, 5, 2),        # tcgen05 grouped path, tiny
    (12, 128, 768, 2048, 16, 8),    # Qwen3-30B-A3B expert shapes (gate / up), decode batch 16, top-8
    (12, 128, 2048, 768, 16, 8),    # ... and the down projection (k = 768: 3 super-blocks)
    (14, 16, 384, 2048, 9, 4),      # Q6_K experts
    (12, 4, 200, 256, 40, 2),       # ragged N, > 32 rows per expert (several chunks)
    (8, 8, 64, 96, 6, 2),           # Q8_0: shape-generic kernel
])
def test_moe_gemm_gguf_matches_oracle(ggml_type, E, N, K, T, k):
    rng = np.random.default_rng(E + N + T)
    stacked = np.concatenate([G.random_weight(rng, ggml_type, N, K).reshape(-1) for _ in range(E)])
    ids = np.stack([rng.permutation(E)[:k] for _ in range(T)]).astype(np.int32)
    tw = rng.uniform(0.05, 0.5, (T, k)).astype(np.float32)
    x = rng.standard_normal((T, K)).astype(np.float32)
    xp = rng.standard_normal((T * k, K)).astype(np.float32)
    wt = torch.from_numpy(stacked).to(DEV)
    e, s = pkg.sort_expert_assignments(torch.from_numpy(ids).to(DEV), E)
    # one row per token, no routing weight (gate / up)
    y = pkg.moe_gemm_gguf(torch.from_numpy(x).to(DEV), wt, ggml_type, (E, N, K), None, s, e, k).cpu().numpy()
    ref = OM.moe_gemm(x, stacked, ggml_type, E, N, K, ids.reshape(-1), k)
    assert y.shape == (T * k, N) and rel_fro(y, ref) < 1e-3, rel_fro(y, ref)
    # one row per pair, routing weights folded in (down)
    y = pkg.moe_gemm_gguf(torch.from_numpy(xp).to(DEV), wt, ggml_type, (E, N, K), torch.from_numpy(tw).to(DEV), s, e, k).cpu().numpy()
    ref = OM.moe_gemm(xp, stacked, ggml_type, E, N, K, ids.reshape(-1), k, tw.reshape(-1))
    assert rel_fro(y, ref) < 1e-3, rel_fro(y, ref)


def test_fused_moe_block_matches_oracle():
    """Qwen3-30B-A3B block shapes scaled down in expert count: hidden 2048, inter 768, 32 experts, top-8, 16 tokens."""
    rng = np.random.default_rng(5)
    E, H, I, k, T = 32, 2048, 768, 8, 16
    gate = (rng.standard_normal((E, H)) * 0.05).astype(np.float32)
    ge = np.concatenate([G.random_weight(rng, 12, I, H).reshape(-1) for _ in range(E)])
    ue = np.concatenate([G.random_weight(rng, 12, I, H).reshape(-1) for _ in range(E)])
    de = np.concatenate([G.random_weight(rng, 12, H, I).reshape(-1) for _ in range(E)])
    x = rng.standard_normal((T, H)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(DEV)
    blk = pkg.FusedMoe(t(gate), t(ge), t(ue), t(de), (12, 12, 12), E, H, I, k)
    y = blk.forward(t(x)).cpu().numpy()
    ref, (rw, rids) = OM.fused_moe(x, gate, ge, ue, de, (12, 12, 12), E, H, I, k)
    assert y.shape == (T, H) and rel_fro(y, ref) < 2e-3, rel_fro(y, ref)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("E,N,K,T,k,by,bx", [
    (16, 512, 2048, 16, 4, 128, 128),     # decode batch on block-FP8 experts (DeepSeek-V3 / Qwen3-FP8 block size)
    (8, 256, 768, 70, 2, 128, 128),       # > 32 rows per expert (several chunks), k = 3 super-blocks
    (4, 100, 96, 5, 2, 128, 128),         # ragged shapes: shape-generic kernel
])
def test_moe_gemm_fp8_matches_oracle(dtype, E, N, K, T, k, by, bx):
    """moe_gemm_fp8 (moe.rs:1447-1473): block-scaled e4m3 experts, 16-bit activations and outputs.  Grouped tcgen05 path for k % 256 == 0,
    n % by == 0; tolerance as fp8_matmul (fp16 operands, one rounding of the output)."""
    from oracle import fp_formats as F
    rng = np.random.default_rng(E + N + K + T)
    w = rng.integers(0, 0x7f, (E, N, K), dtype=np.uint8) | (rng.integers(0, 2, (E, N, K), dtype=np.uint8) << 7)
    w[(w & 0x7f) == 0x7f] = 0x7e                                                  # no NaN codes
    sc = (rng.uniform(0.5, 2.0, (E, -(-N // by), -(-K // bx))) * 1e-3).astype(np.float32)     # checkpoint-sized scales (~1e-3)
    ids = np.stack([rng.permutation(E)[:k] for _ in range(T)]).astype(np.int32)
    tw = rng.uniform(0.05, 0.5, (T, k)).astype(np.float32)
    x = torch.from_numpy(rng.standard_normal((T, K)).astype(np.float32)).to(DEV).to(dtype)
    xp = torch.from_numpy(rng.standard_normal((T * k, K)).astype(np.float32)).to(DEV).to(dtype)
    wt, st = torch.from_numpy(w).to(DEV), torch.from_numpy(sc).to(DEV)
    e, s = pkg.sort_expert_assignments(torch.from_numpy(ids).to(DEV), E)
    tol = 1e-3 if dtype == torch.float16 else 4e-3
    y = pkg.moe_gemm_fp8(x, wt, st, None, s, e, k, by, bx).float().cpu().numpy()
    ref = OM.moe_gemm_fp8(x.float().cpu().numpy(), w, sc, by, bx, ids.reshape(-1), k)
    assert y.shape == (T * k, N) and np.isfinite(y).all() and rel_fro(y, ref) < tol, rel_fro(y, ref)
    y = pkg.moe_gemm_fp8(xp, wt, st, torch.from_numpy(tw).to(DEV), s, e, k, by, bx).float().cpu().numpy()
    ref = OM.moe_gemm_fp8(xp.float().cpu().numpy(), w, sc, by, bx, ids.reshape(-1), k, tw.reshape(-1))
    assert rel_fro(y, ref) < tol, rel_fro(y, ref)
