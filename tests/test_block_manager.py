"""Host-side block manager / prefix cache: the reference's own known-answer tests restated
(/root/reference/src/scheduler/block_engine.rs:1537-1751, /root/reference/src/scheduler/prefix_cache.rs:401-599 -- the only
golden tests the reference holds for the inputs of the hot path) plus copy-on-write, swap and conservation checks."""
import numpy as np
import pytest

from candle_vllm_b200.block_manager import (AllocStatus, BlockManager, PrefixCache, PrefixCacheConfig, Seq, SeqGroup,
                                             cpu_index, is_gpu)
from candle_vllm_b200 import inputs


def _group(gid, sid, bs, tokens):
    s = Seq(sid, tokens, bs)
    return SeqGroup.of(gid, s), s


# ---- block_engine.rs:1537-1751 ------------------------------------------------------------------------------------------
def test_allocate_with_prefix_cache_reuses_blocks():
    bs = 4
    eng = BlockManager(bs, 8, 8, PrefixCacheConfig(True, 4))
    g1, s1 = _group(1, 1, bs, [1, 2, 3, 4, 5, 6, 7, 8])
    free_before = eng.num_free_gpu_blocks()
    eng.allocate(g1)
    free_after_alloc = eng.num_free_gpu_blocks()
    assert free_after_alloc < free_before
    cached_ids = eng.tables[s1.id][:2]
    eng.cache_sequence(s1)
    eng.free_sequence(s1)
    assert eng.num_free_gpu_blocks() == free_after_alloc + 1       # 3 blocks released, 2 of them kept alive by the cache
    g2, s2 = _group(2, 2, bs, list(range(1, 13)))
    eng.allocate(g2)
    assert s2.num_cached_tokens == 8
    assert eng.tables[s2.id][:2] == cached_ids
    eng.check_invariants()


def test_prefix_cache_eviction_does_not_free_active_sequence_blocks():
    bs = 4
    eng = BlockManager(bs, 8, 8, PrefixCacheConfig(True, 4))
    g1, s1 = _group(1, 1, bs, [1, 2, 3, 4, 5, 6, 7, 8])
    eng.allocate(g1); eng.cache_sequence(s1); eng.free_sequence(s1)
    g2, s2 = _group(2, 2, bs, list(range(1, 13)))
    eng.allocate(g2)
    active = eng.tables[s2.id][:2]
    assert eng.evict_prefix_cache_blocks(2) == 2
    for b in active:
        assert b not in eng.free_gpu_block_ids()
    eng.check_invariants()


def test_append_token_slot_repairs_table_after_skipped_boundary_allocation():
    bs = 4
    eng = BlockManager(bs, 4, 4)
    g, s = _group(1, 1, bs, [1, 2, 3, 4])
    eng.allocate(g)
    eng._release(eng.tables[s.id].pop())           # the table falls one block behind
    s.add_token(5)
    assert s.logical_blocks() == 2 and len(eng.tables[s.id]) == 1
    assert eng.can_append_token(g)
    assert eng.append_token_slot(s) is None
    assert len(eng.tables[s.id]) == 2
    eng.check_invariants()


def test_allocate_for_prefill_reserves_and_extends_by_chunk():
    bs = 4
    eng = BlockManager(bs, 4, 4)
    g, s = _group(1, 1, bs, list(range(1, 11)))
    assert eng.can_allocate(g, bs) == AllocStatus.OK
    eng.allocate(g, bs)
    assert len(eng.tables[s.id]) == 1
    s.num_cached_tokens = 4
    assert eng.prefill_chunk_blocks_required(g, bs) == 1
    assert eng.can_append_prefill_chunk(g, bs)
    eng.append_prefill_chunk_slots(g, bs)
    assert len(eng.tables[s.id]) == 2
    s.num_cached_tokens = 8
    eng.append_prefill_chunk_slots(g, bs)
    assert len(eng.tables[s.id]) == 3
    eng.check_invariants()


def test_rebuild_sequence_with_cached_prefix_shrinks_cached_tokens():
    bs = 4
    eng = BlockManager(bs, 8, 8, PrefixCacheConfig(True, 8))
    g, s = _group(1, 1, bs, list(range(1, 13)))
    eng.allocate(g)
    s.num_cached_tokens = 8
    first, n = eng.tables[s.id][0], len(eng.tables[s.id])
    assert eng.rebuild_with_cached_prefix(s, 4)
    assert len(eng.tables[s.id]) == n and eng.tables[s.id][0] == first
    assert s.num_cached_tokens == 4 and s.prefix_hash is not None
    eng.check_invariants()


# ---- prefix_cache.rs:401-599 ---------------------------------------------------------------------------------------------
def _pc(max_blocks):
    return PrefixCache(4, PrefixCacheConfig(True, max_blocks))


def test_prefix_cache_matches_full_blocks():
    c = _pc(8)
    assert c.insert_prefix([1, 2, 3, 4, 5, 6, 7, 8], [0, 1]) == []
    matched, last = c.match_prefix(list(range(1, 13)))
    assert matched == 2
    assert c.blocks_for_match(last) == [0, 1]


def test_prefix_cache_evicts_leaf_blocks():
    c = _pc(1)
    toks = [1, 2, 3, 4, 5, 6, 7, 8]
    assert c.insert_prefix(toks, [5, 6]) == []            # just-inserted blocks are protected
    assert c.cached_blocks() == 2
    assert c.evict_blocks(1) == [6]                         # the leaf goes first
    assert c.match_prefix(toks)[0] == 1


def test_prefix_cache_insert_trims_older_leaves_before_new_prefix():
    c = _pc(2)
    old, new = [1, 2, 3, 4, 5, 6, 7, 8], [9, 10, 11, 12, 13, 14, 15, 16]
    assert c.insert_prefix(old, [1, 2]) == []
    assert c.insert_prefix(new, [3, 4]) == [2, 1]          # leaf, then its parent once that became a leaf
    assert c.match_prefix(old)[0] == 0
    assert c.match_prefix(new)[0] == 2


def test_lru_stays_bounded_after_repeated_touches():
    c = _pc(64)
    c.insert_prefix([1, 2, 3, 4], [0])
    for _ in range(500):
        c.match_prefix([1, 2, 3, 4])
    assert c.lru_entries() < 500


def test_insert_does_not_evict_just_inserted_blocks():
    c = _pc(3)
    c.insert_prefix([10, 20, 30, 40], [10])
    new = list(range(1, 13))
    evicted = c.insert_prefix(new, [0, 1, 2])
    assert not {0, 1, 2} & set(evicted)
    assert c.match_prefix(new)[0] == 3


def test_evict_blocks_respects_protected_set():
    c = _pc(100)
    a, b = [1, 2, 3, 4], [5, 6, 7, 8]
    c.insert_prefix(a, [0]); c.insert_prefix(b, [1])
    assert c.cached_blocks() == 2
    ha = c.match_prefix(a)[1]
    assert c.evict_blocks(2, protected={ha}) == [1]
    assert c.match_prefix(a)[0] == 1


def test_seed_block_affects_only_target_block_hash():
    c = _pc(100)
    toks = list(range(1, 13))
    h = c.hash_for_blocks(toks, 3)
    h0, h1, h2 = (c.hash_for_blocks(toks, 3, 42, i) for i in range(3))
    assert h != h0 and h0 != h1 and h1 != h2
    assert c.hash_for_blocks(toks, 3, 42, 1) == h1
    assert c.hash_for_blocks(toks, 3, 99, 1) != h1
    # the seed only enters at its block: earlier blocks hash as without a seed
    assert c.hash_for_blocks(toks, 1, 42, 1) == c.hash_for_blocks(toks, 1)


# ---- beyond the reference's tests: copy-on-write, swap, admission, conservation ------------------------------------------
def test_fork_then_append_copies_on_write_and_feeds_prepare_decode():
    bs = 4
    eng = BlockManager(bs, 8, 4)
    g, s = _group(1, 1, bs, [1, 2, 3, 4, 5, 6])
    eng.allocate(g)
    child = Seq(2, s.tokens, bs)
    eng.fork(s, child)
    assert eng.tables[child.id] == eng.tables[s.id] and eng.refcount(eng.tables[s.id][-1]) == 2
    child.add_token(7)
    src, dst = eng.append_token_slot(child)                 # last block is shared -> the pair copy_blocks consumes
    assert src == eng.tables[s.id][-1] and dst == eng.tables[child.id][-1] and src != dst
    assert eng.refcount(src) == 1 and eng.tables[child.id][0] == eng.tables[s.id][0]
    s.add_token(8)
    assert eng.append_token_slot(s) is None                 # now exclusively owned
    prep = inputs.prepare_decode([len(s), len(child)], [8, 7], [eng.block_table(s.id), eng.block_table(child.id)], bs)
    assert prep["slot_mapping"].tolist() == [eng.tables[s.id][1] * bs + 2, dst * bs + 2]
    eng.check_invariants()


def test_swap_out_and_in_round_trip_and_rollback():
    bs = 4
    eng = BlockManager(bs, 6, 6)
    g, s = _group(1, 1, bs, list(range(10)))
    eng.allocate(g)
    before = list(eng.tables[s.id])
    assert eng.can_swap_out(g)
    m = eng.swap_out(g)
    assert sorted(m) == sorted(before) and eng.num_free_gpu_blocks() == 6
    assert all(not is_gpu(b) for b in eng.tables[s.id])
    with pytest.raises(Exception, match="swapped out"):
        eng.block_table(s.id)
    eng.rollback_swap_out(g.id)                              # the copy failed: everything back where it was
    assert eng.tables[s.id] == before
    eng.check_invariants()
    m = eng.swap_out(g); eng.finalize_swap_out(g.id)
    assert eng.swap_in_required_blocks(g) == 3 and eng.can_swap_in(g)
    back = eng.swap_in(g)
    assert sorted(back) == sorted(m.values())
    eng.finalize_swap_in(g.id)
    assert all(is_gpu(b) for b in eng.tables[s.id]) and eng.num_free_cpu_blocks() == 6
    eng.check_invariants()


def test_shared_prefix_blocks_stay_on_the_gpu_when_a_group_swaps_out():
    bs = 4
    eng = BlockManager(bs, 8, 8, PrefixCacheConfig(True, 8))
    g1, s1 = _group(1, 1, bs, list(range(1, 9)))
    eng.allocate(g1); eng.cache_sequence(s1); eng.free_sequence(s1)
    g2, s2 = _group(2, 2, bs, list(range(1, 14)))
    eng.allocate(g2)
    assert s2.num_cached_tokens == 8
    shared = eng.tables[s2.id][:2]
    m = eng.swap_out(g2)
    assert not set(shared) & set(m) and eng.tables[s2.id][:2] == shared
    eng.finalize_swap_out(g2.id)
    eng.check_invariants()


def test_admission_status():
    bs = 4
    eng = BlockManager(bs, 4, 0)
    g_big, _ = _group(1, 1, bs, list(range(40)))
    assert eng.can_allocate(g_big) == AllocStatus.IMPOSSIBLE
    g, s = _group(2, 2, bs, list(range(9)))
    assert eng.can_allocate(g) == AllocStatus.OK
    eng.allocate(g)
    g3, _ = _group(3, 3, bs, list(range(9)))
    assert eng.can_allocate(g3) == AllocStatus.LATER
    with pytest.raises(Exception, match="double free"):
        eng._release(eng._gfree[0])


def test_random_traffic_conserves_blocks():
    """seeded random admission / decode / fork / cache / free / swap traffic: reference counts always equal the references held
    by tables and cache, free lists hold exactly the unreferenced blocks, and no live sequence ever loses a block it reads"""
    rng = np.random.default_rng(0)
    bs = 4
    eng = BlockManager(bs, 48, 24, PrefixCacheConfig(True, 12))
    live, next_id = {}, 1
    prompts = [list(rng.integers(1, 50, n)) for n in (5, 8, 8, 12, 13, 16)]
    for step in range(600):
        op = rng.integers(0, 6)
        if op == 0 and len(live) < 8:
            base = prompts[rng.integers(len(prompts))]
            toks = base[:rng.integers(1, len(base) + 1)] + list(rng.integers(1, 50, rng.integers(0, 4)))
            g, s = _group(next_id, next_id, bs, toks)
            if eng.can_allocate(g) == AllocStatus.OK:
                eng.allocate(g); live[s.id] = (g, s); next_id += 1
        elif op == 1 and live:
            g, s = live[list(live)[rng.integers(len(live))]]
            if all(is_gpu(b) for b in eng.tables[s.id]):
                s.add_token(int(rng.integers(1, 50)))
                if eng.can_append_token(g):
                    eng.append_token_slot(s)
                else:
                    s.tokens.pop()
        elif op == 2 and live and len(live) < 8:
            g, s = live[list(live)[rng.integers(len(live))]]
            if all(is_gpu(b) for b in eng.tables[s.id]):
                c = Seq(next_id, s.tokens, bs); eng.fork(s, c)
                live[c.id] = (SeqGroup.of(next_id, c), c); next_id += 1
        elif op == 3 and live:
            sid = list(live)[rng.integers(len(live))]
            g, s = live.pop(sid)
            if all(is_gpu(b) for b in eng.tables[s.id]):
                eng.cache_sequence(s)
            eng.free_sequence(s)
        elif op == 4 and live:
            g, s = live[list(live)[rng.integers(len(live))]]
            if all(is_gpu(b) for b in eng.tables[s.id]) and eng.can_swap_out(g):
                eng.swap_out(g)
                eng.finalize_swap_out(g.id) if rng.integers(2) else eng.rollback_swap_out(g.id)
        elif op == 5 and live:
            g, s = live[list(live)[rng.integers(len(live))]]
            if any(not is_gpu(b) for b in eng.tables[s.id]) and eng.can_swap_in(g):
                eng.swap_in(g)
                eng.finalize_swap_in(g.id) if rng.integers(4) else eng.rollback_swap_in(g.id)
        eng.check_invariants()
        for g, s in live.values():
            assert len(eng.tables[s.id]) >= -(-len(s) // bs) or len(eng.tables[s.id]) == s.logical_blocks() - 1
    for g, s in list(live.values()):
        eng.free_sequence(s)
    eng.evict_prefix_cache_blocks(1000)
    eng.check_invariants()
    assert eng.num_free_gpu_blocks() == 48 and eng.num_free_cpu_blocks() == 24


@pytest.mark.gpu
def test_manager_traffic_drives_copy_and_swap_kernels_bit_exactly():
    """The block manager's copy-on-write pairs and swap mappings go straight into ``CacheEngine.copy`` / ``swap_out`` /
    ``swap_in`` (copy_blocks / swap_blocks through the C ABI): a sequence's KV, read back through its CURRENT block table, must
    be bit-identical before and after a fork + divergent append, a swap-out with the GPU blocks clobbered, and a swap-in that
    lands on different GPU blocks."""
    import torch
    import candle_vllm_b200 as pkg
    bs, layers, kvh, hd = 16, 3, 2, 128
    dev = torch.device("cuda", 0)
    eng = BlockManager(bs, 12, 12)
    ce = pkg.CacheEngine(layers, kvh, hd, pkg.CacheConfig(bs, 12, 12), device=dev)
    g = torch.Generator(device=dev); g.manual_seed(0)
    for k, v in ce.gpu_cache:
        k.copy_(torch.randn(k.shape, device=dev, generator=g).to(k.dtype)); v.copy_(torch.randn(v.shape, device=dev, generator=g).to(v.dtype))

    def gather(seq):
        ids = torch.tensor(eng.block_table(seq.id), device=dev)
        n = len(seq)
        return [(k[ids].reshape(-1, kvh, hd)[:n].clone(), v[ids].reshape(-1, kvh, hd)[:n].clone()) for k, v in ce.gpu_cache]

    def same(a, b, n):
        return all(torch.equal(x[0][:n], y[0][:n]) and torch.equal(x[1][:n], y[1][:n]) for x, y in zip(a, b))

    grp, s = _group(1, 1, bs, list(range(40)))               # 40 tokens: 2 full blocks + 8 tokens in the third
    eng.allocate(grp)
    before = gather(s)
    child = Seq(2, s.tokens, bs)
    eng.fork(s, child)
    child.add_token(99)
    cow = eng.append_token_slot(child)
    assert cow is not None
    ce.copy({cow[0]: [cow[1]]})                               # copy_blocks_bf16
    torch.cuda.synchronize()
    assert same(gather(child), before, 40) and same(gather(s), before, 40)
    assert eng.block_table(child.id)[-1] != eng.block_table(s.id)[-1]

    cgrp = SeqGroup.of(2, child)
    assert eng.can_swap_out(cgrp)
    out_map = eng.swap_out(cgrp)                              # only the block the child owns alone leaves the GPU
    assert list(out_map) == [cow[1]]
    ce.swap_out(out_map)
    torch.cuda.synchronize()
    eng.finalize_swap_out(cgrp.id)
    for k, v in ce.gpu_cache:                                 # whoever gets the freed block next overwrites it
        k[cow[1]].zero_(); v[cow[1]].zero_()
    filler_g, filler = _group(3, 3, bs, list(range(5)))
    eng.allocate(filler_g)                                    # takes a block off the free list so the swap-in lands elsewhere
    in_map = eng.swap_in(cgrp)
    ce.swap_in(in_map)
    torch.cuda.synchronize()
    eng.finalize_swap_in(cgrp.id)
    assert same(gather(child), before, 40) and same(gather(s), before, 40)
    eng.check_invariants()
