"""CPU tests for the host-side logic above the C ABI (metadata builder, shapes, TP sharding)."""
import json
import os

import numpy as np
import pytest
import torch

import candle_vllm_b200 as pkg
from candle_vllm_b200 import cache_engine, inputs, synthetic
from oracle import cache_ops as OC
from oracle import ggml_quants as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_prepare_decode_matches_oracle_and_golden():
    cases = json.load(open(os.path.join(GOLD, "slot_mapping.json")))
    for c in cases:
        p = inputs.prepare_decode([c["seq_len"]], [11], [c["table"]], c["block_size"])
        assert p["slot_mapping"].tolist() == [c["slot"]] and p["positions"].tolist() == [c["position"]]
        assert p["block_tables"].shape[1] == c["used"]
    rng = np.random.default_rng(0)
    lens = [int(x) for x in rng.integers(1, 700, 9)]
    tables = [rng.permutation(200)[: -(-L // 64) + int(rng.integers(0, 3))].tolist() for L in lens]
    a = inputs.prepare_decode(lens, [1] * 9, tables, 64)
    b = OC.prepare_decode(lens, tables, 64)
    for k in ("positions", "slot_mapping", "context_lens", "block_tables"):
        assert np.array_equal(np.asarray(a[k]).astype(np.int64), np.asarray(b[k]).astype(np.int64)), k
    assert a["max_context_len"] == b["max_context_len"]


def test_prepare_decode_table_too_small_errors_like_reference():
    with pytest.raises(pkg.BackendError, match="Block table is too small"):
        inputs.prepare_decode([129], [1], [[1, 2]], 64)


def test_prepare_prompt_chunked():
    prompt = list(range(100, 100 + 150))
    table = [9, 4, 7]
    p = inputs.prepare_prompt([prompt], [table], 64, num_cached_tokens=[64], chunk_size=64)
    assert p["tokens"].tolist() == prompt[64:128]
    assert p["positions"].tolist() == list(range(64, 128))
    assert p["slot_mapping"].tolist() == OC.prefill_slots(table, 64, 128, 64).tolist()
    assert p["cu_seqlens_q"].tolist() == [0, 64] and p["cu_seqlens_k"].tolist() == [0, 128]
    assert p["block_tables"].tolist() == [[9, 4]]


def test_kv_head_shard_matches_reference_rules():
    assert cache_engine.kv_head_shard(8, 3, 4) == (2, 6)
    assert cache_engine.kv_head_shard(8, 5, 8) == (1, 5)
    assert cache_engine.kv_head_shard(2, 5, 8) == (1, 1)        # replicate: ranks 4..7 -> head 1
    assert cache_engine.kv_head_shard(8, 0, 1) == (8, 0)
    with pytest.raises(pkg.BackendError):
        cache_engine.kv_head_shard(6, 0, 4)


def test_cache_engine_shapes_cpu():
    eng = cache_engine.CacheEngine.__new__(cache_engine.CacheEngine)
    eng.block_size, eng.num_kv_heads, eng.head_dim, eng.dtype = 64, 8, 128, torch.bfloat16
    eng.layout = pkg.KvLayout.FLASH
    assert eng.key_block_shape() == (64, 8, 128) == eng.value_block_shape()
    eng.layout = pkg.KvLayout.PAGED
    assert eng.key_block_shape() == OC.paged_k_shape(1, 64, 8, 128, 2)[1:]
    assert eng.value_block_shape() == OC.paged_v_shape(1, 64, 8, 128)[1:]
    eng.dtype = torch.uint8
    assert eng.key_block_shape() == (8, 8, 64, 16)


def test_qtensor_validation_needs_cuda_and_sizes():
    with pytest.raises(pkg.BackendError):
        pkg.QTensor(torch.zeros(144, dtype=torch.uint8), pkg.GgmlType.Q4_K, (1, 100))       # k % 256
    with pytest.raises(pkg.BackendError):
        pkg.QTensor(torch.zeros(100, dtype=torch.uint8), pkg.GgmlType.Q4_K, (1, 256))       # byte count


def test_synthetic_blocks_are_valid_and_sharding_is_consistent():
    gen = torch.Generator(device="cpu"); gen.manual_seed(0)
    w = synthetic.random_q4k(gen, 8, 1024, "cpu").numpy()
    deq = G.dequantize_weight(w, G.GGML_TYPE_Q4_K, 8, 1024)
    assert np.isfinite(deq).all() and 0.005 < deq.std() < 0.05 and abs(deq.mean()) < 0.01
    w6 = synthetic.random_q6k(gen, 8, 1024, "cpu").numpy()
    d6 = G.dequantize_weight(w6, G.GGML_TYPE_Q6_K, 8, 1024)
    assert np.isfinite(d6).all() and 0.005 < d6.std() < 0.05
    # raw-byte shards: dequant(shard) == slice(dequant(full))
    full = pkg.QTensor(torch.from_numpy(w), G.GGML_TYPE_Q4_K, (8, 1024), allow_cpu=True)
    for r in range(2):
        rs = synthetic.shard_rows(full, r, 2)
        assert np.array_equal(G.dequantize_weight(rs.data.numpy(), 12, 4, 1024), deq[4 * r:4 * r + 4])
        cs = synthetic.shard_cols(full, r, 2)
        assert np.array_equal(G.dequantize_weight(cs.data.numpy(), 12, 8, 512), deq[:, 512 * r:512 * r + 512])


def test_prepare_decode_rectangular_tables_match_ragged_path():
    """ndarray block tables take a vectorised path; results (values and dtypes) equal the per-sequence loop."""
    from candle_vllm_b200 import inputs
    rng = np.random.default_rng(11)
    B, bs, width = 9, 16, 12
    tabs = rng.integers(0, 500, (B, width)).astype(np.int32)
    for base in (1, 15, 16, 17, 100, bs * width - B):
        lens = [base + i for i in range(B)]
        toks = [int(t) for t in rng.integers(0, 1000, B)]
        a = inputs.prepare_decode(lens, toks, tabs.tolist(), bs)
        b = inputs.prepare_decode(np.asarray(lens), np.asarray(toks), tabs, bs)
        assert a.keys() == b.keys()
        for k in a:
            assert np.array_equal(a[k], b[k]), k
            if hasattr(a[k], "dtype"):
                assert a[k].dtype == b[k].dtype, k
    with pytest.raises(inputs.BackendError, match="Block table is too small"):
        inputs.prepare_decode(np.full(B, bs * width + 1), np.zeros(B, np.int64), tabs, bs)


def test_flashinfer_csr_matches_the_reference_construction():
    """inputs.rs:477-531: indptr / indices / last_len and the derived kv_len, incl. a sequence that exactly fills its last page and an
    empty one."""
    import candle_vllm_b200 as pkg
    csr = pkg.flashinfer_csr([1, 64, 65, 130, 0], [[3], [5], [7, 1], [9, 0, 2], [4]], 64)
    assert csr["indptr"].tolist() == [0, 1, 2, 4, 7, 7]
    assert csr["indices"].tolist() == [3, 5, 7, 1, 9, 0, 2]
    assert csr["last_len"].tolist() == [1, 64, 1, 2, 0]
    assert csr["kv_len"].tolist() == [1, 64, 65, 130, 0]


def test_mla_and_turboquant_storage_shapes_follow_the_reference():
    """cache_engine.rs:172-185 (MLA) and :401-482 (TurboQuant storage; shapes only -- the algorithm is not in the reference tree)."""
    import torch
    from candle_vllm_b200.cache_engine import allocate_mla_cache, turboquant_layer_shapes
    c = allocate_mla_cache(2, 5, 64, 512, 64, dtype=torch.bfloat16, device="cpu")
    assert len(c) == 2 and c[0][0].shape == (5, 64, 1, 512) and c[0][1].shape == (5, 64, 1, 64) and c[1][0].dtype == torch.bfloat16
    s4 = turboquant_layer_shapes("turbo4", 10, 64, 8, 128, num_shards=2)
    assert s4 == {"v_absmax": (10, 64, 4), "v_quant": (10, 64, 4, 64), "k_absmax": (10, 64, 4), "k_quant": (10, 64, 4, 64)}
    s3 = turboquant_layer_shapes("turbo3", 10, 64, 8, 128)
    assert s3["k_quant"] == (10, 64, 8, 48) and s3["v_quant"] == (10, 64, 8, 64)
    assert set(turboquant_layer_shapes("turbo8", 1, 64, 8, 128)) == {"v_absmax", "v_quant"}
    import pytest
    from candle_vllm_b200 import BackendError
    with pytest.raises(BackendError):
        turboquant_layer_shapes("turbo5", 1, 64, 8, 128)
