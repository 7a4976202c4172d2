"""Host-side KV block manager: block tables, copy-on-write, CPU swap and the prefix cache -- the producer of the
``block_tables`` / ``copy_blocks`` / ``swap_blocks`` inputs of the hot path.

Behavioural mirror of the reference's ``BlockEngine`` and ``PrefixCache``
(/root/reference/src/scheduler/block_engine.rs:201-1484, /root/reference/src/scheduler/prefix_cache.rs:36-383), written from
their observable contract -- the reference's own known-answer tests (block_engine.rs:1537-1751, prefix_cache.rs:401-599) are
restated in tests/test_block_manager.py and are the parity pin for this file.  Not mirrored: Mamba prefix snapshots and
image-seeded hashing call sites (the seed hook itself exists), which are outside the decode path.

Representation (deliberately not the reference's Arc<Mutex<block>> graph): a physical block is an ``int`` -- GPU block ``i`` is
``i >= 0``, CPU block ``j`` is ``-(j + 1)`` -- and reference counts live in two flat integer arrays; free lists are FIFO deques
(allocate pops the front, a block whose count drops to zero is appended, block_engine.rs:112-131), so a table is a plain list
of ints that ``inputs.prepare_decode`` consumes directly.
"""
from __future__ import annotations

import hashlib
import struct
from collections import OrderedDict, deque
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

from ._lib import BackendError


# ---------------------------------------------------------------------------------------------------------------------
# sequences (the slice of scheduler/sequence.rs the block manager reads)
# ---------------------------------------------------------------------------------------------------------------------
class Seq:
    """Token ids + the counters the manager needs (sequence.rs:83-300)."""

    def __init__(self, seq_id: int, prompt: Sequence[int], block_size: int):
        self.id, self.block_size = int(seq_id), int(block_size)
        self.tokens: List[int] = [int(t) for t in prompt]
        self.prompt_len = len(self.tokens)
        self.num_cached_tokens = 0
        self.prefix_hash: Optional[int] = None          # hash of the reused prefix (the reference keeps it for Mamba state)

    def __len__(self) -> int:
        return len(self.tokens)

    def logical_blocks(self) -> int:
        """sequence.rs:208-227: a block that fills up immediately opens the next (empty) one, so n tokens occupy
        n // bs + 1 logical blocks (0 for an empty sequence)."""
        return len(self.tokens) // self.block_size + 1 if self.tokens else 0

    def blocks_to_add_new_tok(self) -> int:
        """sequence.rs:112-120: 1 when the last logical block is empty (or there is none), i.e. the next token opens a block."""
        return 1 if len(self.tokens) % self.block_size == 0 else 0

    def add_token(self, token: int) -> None:
        self.tokens.append(int(token))

    def prefill_chunk_tokens(self, chunk_size: int) -> int:
        """sequence.rs:279-300 without the Mamba warm-up boundary."""
        remaining = max(0, self.prompt_len - self.num_cached_tokens)
        return remaining if chunk_size == 0 else min(remaining, chunk_size)


@dataclass
class SeqGroup:
    id: int
    seqs: "OrderedDict[int, Seq]" = field(default_factory=OrderedDict)

    @classmethod
    def of(cls, group_id: int, *seqs: Seq) -> "SeqGroup":
        return cls(group_id, OrderedDict((s.id, s) for s in seqs))


# ---------------------------------------------------------------------------------------------------------------------
# prefix cache
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class PrefixCacheConfig:
    enabled: bool = False
    max_cached_blocks: int = 0


def _h64(parent: int, payload: bytes) -> int:
    return int.from_bytes(hashlib.blake2b(struct.pack("<Q", parent) + payload, digest_size=8).digest(), "little")


class PrefixCache:
    """Chained-hash trie over FULL token blocks with least-recently-used eviction of leaves (prefix_cache.rs:36-383).
    An entry's hash covers its parent's hash and its own ``block_size`` tokens, so a hit on block i implies hits on 0..i-1.
    Only leaves may be evicted; evicting a leaf may turn its parent into a leaf (queued as most recent, like the reference).
    Blocks are opaque ints here; ``retain`` is called once for every block the cache starts to hold, and evicted blocks are
    RETURNED for the owner to release -- the cache never frees anything itself."""

    def __init__(self, block_size: int, config: PrefixCacheConfig):
        self.block_size, self.config = int(block_size), config
        self._e: Dict[int, list] = {}                 # hash -> [parent hash | None, block, children]
        self._leaves: "OrderedDict[int, None]" = OrderedDict()   # oldest first

    def enabled(self) -> bool:
        return bool(self.config.enabled) and self.config.max_cached_blocks > 0

    def cached_blocks(self) -> int:
        return len(self._e)

    def lru_entries(self) -> int:
        """size of the recency structure (bounded by the number of leaves; prefix_cache.rs:294-309 compacts a lazy queue)."""
        return len(self._leaves)

    # -- hashing ------------------------------------------------------------------------------------------------
    def _chain(self, tokens: Sequence[int], full_blocks: int, seed: Optional[int], seed_block: Optional[int]) -> Iterable[int]:
        parent = 0
        bs = self.block_size
        for i in range(full_blocks):
            if seed is not None and seed_block is not None and i == seed_block:
                parent = _h64(parent, b"seed" + struct.pack("<Q", seed & 0xFFFFFFFFFFFFFFFF))
            parent = _h64(parent, struct.pack(f"<{bs}I", *[t & 0xFFFFFFFF for t in tokens[i * bs:(i + 1) * bs]]))
            yield parent

    def hash_for_blocks(self, tokens: Sequence[int], full_blocks: int, seed: Optional[int] = None,
                        seed_block: Optional[int] = None) -> Optional[int]:
        if not self.enabled() or full_blocks == 0:
            return None
        last = None
        for last in self._chain(tokens, min(full_blocks, len(tokens) // self.block_size), seed, seed_block):
            pass
        return last

    # -- lookup -------------------------------------------------------------------------------------------------
    def match_prefix(self, tokens: Sequence[int], seed: Optional[int] = None, seed_block: Optional[int] = None) -> Tuple[int, Optional[int]]:
        """(matched full blocks, hash of the last matched block); matched entries become most recently used."""
        if not self.enabled():
            return 0, None
        matched, last = 0, None
        for h in self._chain(tokens, len(tokens) // self.block_size, seed, seed_block):
            if h not in self._e:
                break
            matched, last = matched + 1, h
            if h in self._leaves:
                self._leaves.move_to_end(h)
        return matched, last

    def hashes_for_match(self, last_hash: int) -> List[int]:
        out, cur = [], last_hash
        while cur is not None and cur in self._e:
            out.append(cur)
            cur = self._e[cur][0]
        return out[::-1]

    def blocks_for_match(self, last_hash: int) -> List[int]:
        return [self._e[h][1] for h in self.hashes_for_match(last_hash)]

    # -- update -------------------------------------------------------------------------------------------------
    def insert_prefix(self, tokens: Sequence[int], blocks: Sequence[int], retain=lambda b: None, seed: Optional[int] = None,
                      seed_block: Optional[int] = None) -> List[int]:
        """Cache the first min(full blocks of tokens, len(blocks)) blocks; returns the blocks evicted to respect
        ``max_cached_blocks`` (never the ones just inserted or refreshed: prefix_cache.rs:203-258)."""
        if not self.enabled():
            return []
        n = min(len(tokens) // self.block_size, len(blocks))
        parent, protected = None, set()
        for h, blk in zip(self._chain(tokens, n, seed, seed_block), blocks):
            protected.add(h)
            if h in self._e:
                if h in self._leaves:
                    self._leaves.move_to_end(h)
            else:
                if parent is not None:
                    pe = self._e[parent]
                    if pe[2] == 0:
                        self._leaves.pop(parent, None)
                    pe[2] += 1
                retain(blk)
                self._e[h] = [parent, blk, 0]
                self._leaves[h] = None
            parent = h
        excess = len(self._e) - self.config.max_cached_blocks
        return self.evict_blocks(excess, protected) if excess > 0 else []

    def evict_blocks(self, num_blocks: int, protected: Iterable[int] = ()) -> List[int]:
        protected = set(protected)
        out: List[int] = []
        while len(out) < num_blocks:
            victim = next((h for h in self._leaves if h not in protected), None)
            if victim is None:
                break
            parent, blk, _ = self._e.pop(victim)
            del self._leaves[victim]
            if parent is not None and parent in self._e:
                pe = self._e[parent]
                pe[2] = max(0, pe[2] - 1)
                if pe[2] == 0:
                    self._leaves[parent] = None          # a new leaf queues as most recent (prefix_cache.rs:356-365)
            out.append(blk)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# block manager
# ---------------------------------------------------------------------------------------------------------------------
def is_gpu(block: int) -> bool:
    return block >= 0


def cpu_block(index: int) -> int:
    return -(index + 1)


def cpu_index(block: int) -> int:
    return -block - 1


class AllocStatus:
    OK, LATER, IMPOSSIBLE = "ok", "later", "impossible"          # block_engine.rs:183-199


class BlockManager:
    def __init__(self, block_size: int, num_gpu_blocks: int, num_cpu_blocks: int,
                 prefix_cache: Optional[PrefixCacheConfig] = None):
        self.block_size, self.num_gpu_blocks, self.num_cpu_blocks = int(block_size), int(num_gpu_blocks), int(num_cpu_blocks)
        self._gref = [0] * self.num_gpu_blocks
        self._cref = [0] * self.num_cpu_blocks
        self._gfree: "deque[int]" = deque(range(self.num_gpu_blocks))
        self._cfree: "deque[int]" = deque(cpu_block(i) for i in range(self.num_cpu_blocks))
        self.tables: Dict[int, List[int]] = {}
        pc = prefix_cache or PrefixCacheConfig()
        self.prefix_cache = PrefixCache(self.block_size, pc) if pc.enabled and pc.max_cached_blocks > 0 else None
        self._pending_out: Dict[int, Tuple[dict, dict]] = {}
        self._pending_in: Dict[int, Tuple[dict, dict]] = {}

    # -- counters -----------------------------------------------------------------------------------------------
    def num_free_gpu_blocks(self) -> int:
        return len(self._gfree)

    def num_free_cpu_blocks(self) -> int:
        return len(self._cfree)

    def refcount(self, block: int) -> int:
        return self._gref[block] if is_gpu(block) else self._cref[cpu_index(block)]

    def free_gpu_block_ids(self) -> List[int]:
        return list(self._gfree)

    def block_table(self, seq_id: int) -> List[int]:
        """device block ids of a sequence -- the row ``inputs.prepare_decode`` takes"""
        t = self.tables[seq_id]
        if any(not is_gpu(b) for b in t):
            raise BackendError(f"sequence {seq_id} is swapped out")
        return list(t)

    # -- raw allocation -----------------------------------------------------------------------------------------
    def _alloc(self, gpu: bool = True) -> int:
        free = self._gfree if gpu else self._cfree
        if not free:
            raise BackendError("out of %s blocks" % ("GPU" if gpu else "CPU"))
        b = free.popleft()
        if gpu:
            self._gref[b] = 1
        else:
            self._cref[cpu_index(b)] = 1
        return b

    def _retain(self, b: int) -> None:
        """+1; a block sitting on the free list (count 0) is taken off it (block_engine.rs:904-922)"""
        ref, idx, free = (self._gref, b, self._gfree) if is_gpu(b) else (self._cref, cpu_index(b), self._cfree)
        if ref[idx] == 0:
            free.remove(b)
        ref[idx] += 1

    def _release(self, b: int) -> None:
        ref, idx, free = (self._gref, b, self._gfree) if is_gpu(b) else (self._cref, cpu_index(b), self._cfree)
        if ref[idx] == 0:
            raise BackendError(f"physical block {b} experienced a double free")     # block_engine.rs:120-125
        ref[idx] -= 1
        if ref[idx] == 0:
            free.append(b)

    # -- admission ----------------------------------------------------------------------------------------------
    def _blocks_without_prefix(self, group: SeqGroup, chunk: int) -> int:
        """block_engine.rs:455-475"""
        if chunk == 0:
            return sum(s.logical_blocks() for s in group.seqs.values())
        return sum(-(-self._chunk_end(s.prompt_len, 0, chunk) // self.block_size) for s in group.seqs.values())

    @staticmethod
    def _chunk_end(prompt_len: int, cached: int, chunk: int) -> int:
        """block_engine.rs:417-439 without the Mamba warm-up boundary"""
        cached = min(cached, prompt_len)
        return prompt_len if chunk == 0 else cached + min(prompt_len - cached, chunk)

    def _match_for_allocate(self, s: Seq) -> Tuple[int, Optional[int]]:
        """prefix-cache hit for a new sequence; a prompt that is cached to its last token still recomputes its final block so that
        there is something to prefill (block_engine.rs:1352-1359)"""
        matched, last = self.prefix_cache.match_prefix(s.tokens)
        if matched > 0 and matched == len(s.tokens) // self.block_size and len(s.tokens) % self.block_size == 0:
            matched -= 1
        return matched, last

    def can_allocate(self, group: SeqGroup, chunk: int = 0) -> str:
        """block_engine.rs:292-373: blocks the first prefill chunk needs (net of prefix-cache hits) against the free list -- evicting
        cached prefixes if that helps -- and the whole prompt against the pool size."""
        total = sum(s.logical_blocks() for s in group.seqs.values())
        if self.prefix_cache is not None:
            first = next(iter(group.seqs.values()))
            matched, _ = self._match_for_allocate(first)
            need = max(0, -(-self._chunk_end(len(first.tokens), matched * self.block_size, chunk) // self.block_size) - matched)
        else:
            need = sum(-(-self._chunk_end(s.prompt_len, 0, chunk) // self.block_size) for s in group.seqs.values())
        if self.num_free_gpu_blocks() < need:
            self.evict_prefix_cache_until_free(need)
        if self.num_gpu_blocks < total:
            return AllocStatus.IMPOSSIBLE
        return AllocStatus.OK if self.num_free_gpu_blocks() >= need else AllocStatus.LATER

    def allocate(self, group: SeqGroup, chunk: int = 0) -> None:
        """block_engine.rs:382-415 / :1331-1465.  All sequences of a group share the prompt's blocks (fork semantics)."""
        seqs = list(group.seqs.values())
        table: List[int] = []
        cached = 0
        first = seqs[0]
        if self.prefix_cache is not None:
            matched, last = self._match_for_allocate(first)
            first.prefix_hash = None
            if matched > 0:
                hashes = self.prefix_cache.hashes_for_match(last)
                first.prefix_hash = hashes[matched - 1]
                for b in self.prefix_cache.blocks_for_match(last)[:matched]:
                    self._retain(b)
                    table.append(b)
            cached = matched * self.block_size
            first.num_cached_tokens = cached
            need = first.logical_blocks() if chunk == 0 else -(-(first.prefill_chunk_tokens(chunk) + cached) // self.block_size)
        else:
            need = self._blocks_without_prefix(SeqGroup.of(0, first), chunk)
        while len(table) < need:
            table.append(self._alloc())
        for i, s in enumerate(seqs):
            s.num_cached_tokens = cached
            if i > 0:
                s.prefix_hash = first.prefix_hash
                for b in table:
                    self._retain(b)
            self.tables[s.id] = list(table)

    def fork(self, parent: Seq, child: Seq) -> None:
        """child shares every block of parent (beam / parallel sampling); the first divergent append copies on write"""
        t = list(self.tables[parent.id])
        for b in t:
            self._retain(b)
        self.tables[child.id] = t
        child.num_cached_tokens, child.prefix_hash = parent.num_cached_tokens, parent.prefix_hash

    # -- decode growth ------------------------------------------------------------------------------------------
    def _missing(self, s: Seq) -> int:
        return max(0, s.logical_blocks() - len(self.tables.get(s.id, ())))

    def _needed_to_append(self, s: Seq) -> int:
        m = self._missing(s)
        if m:
            return m
        if s.blocks_to_add_new_tok():
            return 0
        t = self.tables.get(s.id)
        return int(bool(t) and self.refcount(t[-1]) > 1)

    def can_append_token(self, group: SeqGroup) -> bool:
        return sum(self._needed_to_append(s) for s in group.seqs.values()) <= self.num_free_gpu_blocks()

    def append_token_slot(self, s: Seq) -> Optional[Tuple[int, int]]:
        """Make room for the token just added to ``s``; returns the copy-on-write pair (src, dst) for ``copy_blocks`` when the
        last block is shared (block_engine.rs:1181-1212).  A table that fell behind the logical blocks is repaired first."""
        t = self.tables[s.id]
        m = self._missing(s)
        if m:
            t.extend(self._alloc() for _ in range(m))
            return None
        if s.blocks_to_add_new_tok():
            return None
        last = t[-1]
        if not is_gpu(last):
            raise BackendError(f"sequence {s.id}: last block is on the CPU")
        if self.refcount(last) == 1:
            return None
        new = self._alloc()
        self._release(last)
        t[-1] = new
        return last, new

    # -- chunked prefill ----------------------------------------------------------------------------------------
    def _missing_for_chunk(self, s: Seq, chunk: int) -> int:
        end = s.num_cached_tokens + s.prefill_chunk_tokens(chunk)
        return max(0, -(-end // self.block_size) - len(self.tables.get(s.id, ())))

    def prefill_chunk_blocks_required(self, group: SeqGroup, chunk: int) -> int:
        return sum(self._missing_for_chunk(s, chunk) for s in group.seqs.values())

    def can_append_prefill_chunk(self, group: SeqGroup, chunk: int) -> bool:
        return self.prefill_chunk_blocks_required(group, chunk) <= self.num_free_gpu_blocks()

    def append_prefill_chunk_slots(self, group: SeqGroup, chunk: int) -> None:
        for s in group.seqs.values():
            self.tables[s.id].extend(self._alloc() for _ in range(self._missing_for_chunk(s, chunk)))

    # -- release / cache ----------------------------------------------------------------------------------------
    def free_sequence(self, s: Seq) -> None:
        for b in self.tables.pop(s.id):
            self._release(b)

    def cache_sequence(self, s: Seq) -> None:
        """offer the sequence's full blocks to the prefix cache (block_engine.rs:594-646)"""
        pc = self.prefix_cache
        t = self.tables.get(s.id)
        full = len(s.tokens) // self.block_size
        if pc is None or t is None or full == 0 or len(t) < full or any(not is_gpu(b) for b in t[:full]):
            return
        for b in pc.insert_prefix(s.tokens, t[:full], retain=self._retain):
            self._release(b)

    def evict_prefix_cache_blocks(self, n: int) -> int:
        if self.prefix_cache is None or n <= 0:
            return 0
        ev = self.prefix_cache.evict_blocks(n)
        for b in ev:
            self._release(b)
        return len(ev)

    def evict_prefix_cache_until_free(self, min_free: int) -> int:
        total = 0
        while self.num_free_gpu_blocks() < min_free and self.evict_prefix_cache_blocks(1):
            total += 1
        return total

    def fallback_to_full_prefill(self, s: Seq) -> bool:
        """fresh, unshared blocks for the whole sequence (block_engine.rs:933-965); False (state untouched) when they do not fit"""
        old = self.tables.pop(s.id, None)
        if old is None:
            return False
        for b in old:
            self._release(b)
        need = s.logical_blocks()
        if self.num_free_gpu_blocks() < need:
            self.evict_prefix_cache_until_free(need)
        if self.num_free_gpu_blocks() < need:
            for b in old:
                self._retain(b)
            self.tables[s.id] = old
            return False
        self.tables[s.id] = [self._alloc() for _ in range(need)]
        s.num_cached_tokens, s.prefix_hash = 0, None
        return True

    def rebuild_with_cached_prefix(self, s: Seq, cached_tokens: int) -> bool:
        """keep only the first ``cached_tokens`` worth of the table's blocks and re-allocate the rest (block_engine.rs:968-1038)"""
        full = cached_tokens // self.block_size
        logical = s.logical_blocks()
        target = self.prefix_cache.hash_for_blocks(s.tokens, full) if self.prefix_cache is not None else None
        if cached_tokens == 0 or full == 0 or full > logical or target is None:
            return self.fallback_to_full_prefill(s)
        old = self.tables.pop(s.id, None)
        if old is None:
            return False
        if len(old) < full:
            self.tables[s.id] = old
            return False
        prefix = old[:full]
        for b in old:
            self._release(b)
        for b in prefix:
            self._retain(b)
        suffix = max(0, logical - full)
        if self.num_free_gpu_blocks() < suffix:
            self.evict_prefix_cache_until_free(suffix)
        if self.num_free_gpu_blocks() < suffix:
            for b in prefix:
                self._release(b)
            for b in old:
                self._retain(b)
            self.tables[s.id] = old
            return False
        self.tables[s.id] = prefix + [self._alloc() for _ in range(suffix)]
        s.num_cached_tokens, s.prefix_hash = full * self.block_size, target
        return True

    # -- CPU swap -----------------------------------------------------------------------------------------------
    def _prefix_counts(self, group: SeqGroup) -> Dict[int, int]:
        """blocks at the head of each table that the group does not own alone: cached prefix + anything shared
        (block_engine.rs:1083-1106); they stay on the GPU"""
        out = {}
        for sid, s in group.seqs.items():
            t = self.tables[sid]
            n = min(s.num_cached_tokens // self.block_size if self.prefix_cache is not None else 0, len(t))
            while n < len(t) and self.refcount(t[n]) > 1:
                n += 1
            out[sid] = n
        return out

    def can_swap_out(self, group: SeqGroup) -> bool:
        pre = self._prefix_counts(group)
        need, refs = set(), {}
        for sid in group.seqs:
            for b in self.tables[sid][pre[sid]:]:
                if not is_gpu(b):
                    return False
                need.add(b)
                refs[b] = refs.get(b, 0) + 1
        if any(self.refcount(b) > c for b, c in refs.items()):       # somebody outside the group still reads it
            return False
        return bool(need) and len(need) <= self.num_free_cpu_blocks()

    def swap_out(self, group: SeqGroup) -> Dict[int, int]:
        """-> {gpu block id: cpu block index} for ``swap_blocks`` (block_engine.rs:1122-1177); GPU blocks are released at once,
        ``rollback_swap_out`` restores them if the copy fails"""
        pre = self._prefix_counts(group)
        mapping: Dict[int, int] = {}
        old, new = {}, {}
        for sid in group.seqs:
            o = self.tables[sid]
            n = list(o)
            for i in range(pre[sid], len(o)):
                g = o[i]
                if g in mapping:
                    self._retain(mapping[g])
                else:
                    mapping[g] = self._alloc(gpu=False)
                n[i] = mapping[g]
            old[sid], new[sid] = o, n
            self.tables[sid] = n
        for sid in group.seqs:
            for b in old[sid]:
                if b not in new[sid]:
                    self._release(b)
        self._pending_out[group.id] = (old, new)
        return {g: cpu_index(c) for g, c in mapping.items()}

    def swap_in_required_blocks(self, group: SeqGroup) -> int:
        return len({b for sid in group.seqs for b in self.tables[sid] if not is_gpu(b)})

    def can_swap_in(self, group: SeqGroup) -> bool:
        return self.swap_in_required_blocks(group) <= self.num_free_gpu_blocks()

    def swap_in(self, group: SeqGroup) -> Dict[int, int]:
        """-> {cpu block index: gpu block id} (block_engine.rs:1223-1264); CPU blocks are released by ``finalize_swap_in``"""
        mapping: Dict[int, int] = {}
        old, new = {}, {}
        for sid in group.seqs:
            o = self.tables[sid]
            n = list(o)
            for i, c in enumerate(o):
                if is_gpu(c):
                    continue
                if c in mapping:
                    self._retain(mapping[c])
                else:
                    mapping[c] = self._alloc()
                n[i] = mapping[c]
            old[sid], new[sid] = o, n
            self.tables[sid] = n
        self._pending_in[group.id] = (old, new)
        return {cpu_index(c): g for c, g in mapping.items()}

    def finalize_swap_out(self, group_id: int) -> None:
        self._pending_out.pop(group_id, None)

    def rollback_swap_out(self, group_id: int) -> None:
        old, new = self._pending_out.pop(group_id, ({}, {}))
        for sid, o in old.items():
            for b in o:
                if b not in new[sid]:
                    self._retain(b)
            for b in new[sid]:
                if b not in o:
                    self._release(b)
            self.tables[sid] = o

    def finalize_swap_in(self, group_id: int) -> None:
        old, new = self._pending_in.pop(group_id, ({}, {}))
        for sid, o in old.items():
            for b in o:
                if b not in new[sid]:
                    self._release(b)

    def rollback_swap_in(self, group_id: int) -> None:
        old, new = self._pending_in.pop(group_id, ({}, {}))
        for sid, o in old.items():
            for b in new[sid]:
                if b not in o:
                    self._release(b)
            self.tables[sid] = o

    # -- invariants (used by the tests) ---------------------------------------------------------------------------
    def check_invariants(self) -> None:
        """every reference is accounted for: tables + prefix cache (+ CPU copies pending release) == refcounts, and the
        free lists hold exactly the zero-count blocks"""
        want: Dict[int, int] = {}
        for t in self.tables.values():
            for b in t:
                want[b] = want.get(b, 0) + 1
        if self.prefix_cache is not None:
            for e in self.prefix_cache._e.values():
                want[e[1]] = want.get(e[1], 0) + 1
        for old, new in self._pending_in.values():
            for sid, o in old.items():
                for b in o:
                    if b not in new[sid]:
                        want[b] = want.get(b, 0) + 1
        for b in range(self.num_gpu_blocks):
            assert self._gref[b] == want.get(b, 0), f"gpu block {b}: refcount {self._gref[b]} vs {want.get(b, 0)} references"
        for i in range(self.num_cpu_blocks):
            assert self._cref[i] == want.get(cpu_block(i), 0), f"cpu block {i}: refcount {self._cref[i]}"
        assert sorted(self._gfree) == [b for b in range(self.num_gpu_blocks) if self._gref[b] == 0]
        assert sorted(cpu_index(b) for b in self._cfree) == [i for i in range(self.num_cpu_blocks) if self._cref[i] == 0]
