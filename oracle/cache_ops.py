"""KV-cache block ops restated in numpy: layouts, slot arithmetic, reshape_and_cache,
copy_blocks, swap_blocks, FP8(e4m3) KV cast.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.  PARITY UNPINNED for the
kernels (they live in attention-rs @a97f519, not in /root/reference); the semantics
follow the in-tree call sites cited on each function.
"""
from __future__ import annotations

import numpy as np

PAD_SLOT_ID = -1          # src/openai/pipelines/llm_engine.rs:94


# --------------------------------------------------------------------------------------
# layouts  (src/scheduler/cache_engine.rs:298-341)
# --------------------------------------------------------------------------------------
def flash_kv_shape(num_blocks, block_size, num_kv_heads, head_dim):
    """Flash layout, K and V alike: [num_blocks, block_size, num_kv_heads, head_dim]
    (``calculate_flash_key_value_block_shape`` cache_engine.rs:326-340)."""
    return (num_blocks, block_size, num_kv_heads, head_dim)


def paged_k_shape(num_blocks, block_size, num_kv_heads, head_dim, elem_size):
    """Legacy paged K: [num_blocks, kvh, head_dim/x, block_size, x], x = 16/elem_size
    (``calculate_key_block_shape`` cache_engine.rs:298-312)."""
    x = 16 // elem_size
    return (num_blocks, num_kv_heads, head_dim // x, block_size, x)


def paged_v_shape(num_blocks, block_size, num_kv_heads, head_dim):
    """Legacy paged V: [num_blocks, kvh, head_dim, block_size] (cache_engine.rs:314-324)."""
    return (num_blocks, num_kv_heads, head_dim, block_size)


# --------------------------------------------------------------------------------------
# slot / block-table arithmetic  (src/openai/pipelines/inputs.rs)
# --------------------------------------------------------------------------------------
def used_blocks_for_len(seq_len: int, block_size: int, table_len: int) -> int:
    """inputs.rs:12-22."""
    if seq_len == 0:
        return 0
    return min(-(-seq_len // block_size), table_len)


def decode_slot(block_table, position: int, block_size: int) -> int:
    """inputs.rs:410-423: slot = table[pos / bs] * bs + pos % bs; table too small -> error."""
    bi = position // block_size
    if bi >= len(block_table):
        raise ValueError(f"Block table is too small (completion)! start_pos={position} "
                         f"block_size={block_size} table_len={len(block_table)}")
    return int(block_table[bi]) * block_size + position % block_size


def prepare_decode(seq_lens, block_tables, block_size: int):
    """Restates ``LLMEngine::prepare_decode`` (inputs.rs:376-454,552-568) for plain lists.

    seq_lens[i] = sequence length INCLUDING the token being decoded; block_tables[i] = list of
    physical block ids.  Returns dict(positions i64[B], slot_mapping i64[B], context_lens u32[B],
    block_tables u32[B, max_used] zero padded, max_context_len).
    """
    positions, slots, ctx, tabs = [], [], [], []
    for L, table in zip(seq_lens, block_tables):
        pos = L - 1
        positions.append(pos)
        ctx.append(L)
        slots.append(decode_slot(table, pos, block_size))
        used = used_blocks_for_len(L, block_size, len(table))
        tabs.append(list(table[:used]))
    width = max(len(t) for t in tabs)
    bt = np.zeros((len(tabs), width), np.uint32)
    for i, t in enumerate(tabs):
        bt[i, :len(t)] = t
    return dict(positions=np.asarray(positions, np.int64), slot_mapping=np.asarray(slots, np.int64),
                context_lens=np.asarray(ctx, np.uint32), block_tables=bt,
                max_context_len=int(max(ctx)))


def prefill_slots(block_table, start: int, end: int, block_size: int):
    """inputs.rs:180-194: one slot per prompt position in [start, end)."""
    return np.asarray([decode_slot(block_table, p, block_size) for p in range(start, end)], np.int64)


# --------------------------------------------------------------------------------------
# FP8 e4m3 (fn variant: no inf, max 448) round-to-nearest-even, saturating
# --------------------------------------------------------------------------------------
def f32_to_e4m3(x: np.ndarray) -> np.ndarray:
    """f32 -> u8 (e4m3fn bits).  Saturating (|x|>448 -> 448; NaN -> 0x7f), RNE, subnormals kept.

    The reference stores FP8 KV as U8 (src/main.rs:263-267) and passes NO scale tensor at any
    call site (attention.rs:566-575,888-897), so scale 1.0 is assumed.
    """
    x = np.asarray(x, np.float32)
    sign = (np.signbit(x)).astype(np.uint8) << 7
    a = np.abs(x).astype(np.float64)
    out = np.zeros(x.shape, np.uint8)
    nan = np.isnan(x)
    a = np.where(nan, 0, a)
    a = np.minimum(a, 448.0)
    # normal range: exponent e in [-6, 8]; quantum = 2^(e-3); subnormal quantum = 2^-9
    e = np.floor(np.log2(np.where(a > 0, a, 1.0)))
    e = np.clip(e, -6, 8)
    quantum = np.exp2(e - 3)
    q = np.rint(a / quantum)            # RNE on an exactly representable ratio (f64)
    val = q * quantum                   # may round up to next binade: recompute fields from val
    val = np.minimum(val, 448.0)
    e2 = np.floor(np.log2(np.where(val > 0, val, 1.0)))
    e2 = np.clip(e2, -6, 8)
    is_sub = val < 2.0 ** -6
    mant = np.where(is_sub, np.rint(val / 2.0 ** -9), np.rint(val / np.exp2(e2 - 3)) - 8)
    expf = np.where(is_sub, 0, e2 + 7)
    out = (expf.astype(np.uint8) << 3) | mant.astype(np.uint8)
    out = np.where(val == 0, 0, out).astype(np.uint8)
    out = out | sign
    out = np.where(nan, np.uint8(0x7F), out)
    return out.astype(np.uint8)


def e4m3_to_f32(b: np.ndarray) -> np.ndarray:
    b = np.asarray(b, np.uint8)
    s = np.where(b & 0x80, -1.0, 1.0)
    e = ((b >> 3) & 0xF).astype(np.int32)
    m = (b & 7).astype(np.float64)
    v = np.where(e == 0, m * 2.0 ** -9, (8 + m) * np.exp2(e - 10.0))
    v = np.where((b & 0x7F) == 0x7F, np.nan, v)
    return (s * v).astype(np.float32)


# --------------------------------------------------------------------------------------
# reshape_and_cache
# --------------------------------------------------------------------------------------
def reshape_and_cache_flash(key, value, k_cache, v_cache, slot_mapping, fp8=False):
    """key/value [T, kvh, hd] -> caches [nb, bs, kvh, hd] at flat slot ``slot_mapping[t]``;
    slot < 0 (pad, -1) is skipped.  In place.  (K3; slot math inputs.rs:180-194,410-423.)"""
    nb, bs, kvh, hd = k_cache.shape
    kf = k_cache.reshape(nb * bs, kvh, hd)
    vf = v_cache.reshape(nb * bs, kvh, hd)
    for t, s in enumerate(np.asarray(slot_mapping, np.int64)):
        if s < 0:
            continue
        kf[s] = f32_to_e4m3(key[t]) if fp8 else key[t]
        vf[s] = f32_to_e4m3(value[t]) if fp8 else value[t]


def reshape_and_cache_paged(key, value, k_cache, v_cache, slot_mapping, fp8=False):
    """Legacy layout: K [nb, kvh, hd/x, bs, x], V [nb, kvh, hd, bs] (cache_engine.rs:298-324)."""
    nb, kvh, hdx, bs, x = k_cache.shape
    for t, s in enumerate(np.asarray(slot_mapping, np.int64)):
        if s < 0:
            continue
        b, o = divmod(int(s), bs)
        kk = f32_to_e4m3(key[t]) if fp8 else key[t]
        vv = f32_to_e4m3(value[t]) if fp8 else value[t]
        k_cache[b, :, :, o, :] = kk.reshape(kvh, hdx, x)
        v_cache[b, :, :, o] = vv


# --------------------------------------------------------------------------------------
# copy_blocks / swap_blocks
# --------------------------------------------------------------------------------------
def copy_blocks(key_caches, value_caches, block_mapping):
    """``backend::copy_blocks`` (src/backend/cache.rs:15-165): for every layer, K and V,
    dst_block <- src_block for every (src, dst) pair; one src may fan out to several dst
    (``HashMap<usize, Vec<usize>>`` cache.rs:103-109).  ``block_mapping``: dict src -> [dst...]
    or a flat list of (src, dst) pairs.  In place."""
    pairs = mapping_pairs(block_mapping)
    for kc, vc in zip(key_caches, value_caches):
        for s, d in pairs:
            kc[d] = kc[s]
            vc[d] = vc[s]


def mapping_pairs(block_mapping):
    if isinstance(block_mapping, dict):
        pairs = []
        for s, ds in block_mapping.items():
            if isinstance(ds, (list, tuple)):
                pairs += [(int(s), int(d)) for d in ds]
            else:
                pairs.append((int(s), int(ds)))
        return pairs
    return [(int(s), int(d)) for s, d in block_mapping]


def swap_blocks(src, dst, mapping):
    """``attention_rs::cache::swap_blocks(src, dst, &HashMap<usize,usize>)``
    (call site src/scheduler/cache_engine.rs:527-535): dst[d] <- src[s] block-wise, src and dst
    may live on different devices (GPU cache <-> CPU cache).  In place on ``dst``."""
    for s, d in mapping_pairs(mapping):
        dst[d] = src[s]
