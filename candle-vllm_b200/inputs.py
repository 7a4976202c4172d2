"""Metadata builder: mirror of ``LLMEngine::prepare_decode`` / ``prepare_prompt``
(/root/reference/src/openai/pipelines/inputs.rs:90-374, :376-575) for plain sequences.

Pure host-side integer logic (numpy); ``to_device`` wraps the result into the ``InputMetadata`` the
kernels consume (5 small H2D copies per step, like the reference).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from ._lib import BackendError

PAD_SLOT_ID = -1                      # llm_engine.rs:94
PREFILL_CHUNK_SIZE = 8192             # llm_engine.rs:95


def used_blocks_for_len(seq_len: int, block_size: int, table_len: int) -> int:
    """inputs.rs:12-22"""
    if seq_len == 0:
        return 0
    return min((seq_len + block_size - 1) // block_size, table_len)


def _slot(table: Sequence[int], position: int, block_size: int, what: str) -> int:
    bi = position // block_size
    if bi >= len(table):
        raise BackendError(f"Block table is too small ({what})! start_pos={position} block_size={block_size} table_len={len(table)}")
    return int(table[bi]) * block_size + position % block_size


def _pad_tables(tables: List[List[int]]) -> np.ndarray:
    width = max(1, max(len(t) for t in tables))
    out = np.zeros((len(tables), width), np.int32)       # _make_tensor_with_pad(.., pad = 0)
    for i, t in enumerate(tables):
        out[i, :len(t)] = t
    return out


def prepare_decode(seq_lens: Sequence[int], last_tokens: Sequence[int], block_tables: Sequence[Sequence[int]],
                   block_size: int) -> dict:
    """inputs.rs:376-454,552-568.  seq_lens include the token being decoded.  ``block_tables`` may be a list of per-sequence
    lists (ragged, like ``Sequence::block_table``) or a rectangular integer ndarray [B, width] whose rows are valid up to
    each sequence's used-block count -- the latter takes a vectorised path with identical results."""
    if isinstance(block_tables, np.ndarray) and block_tables.ndim == 2:
        return _prepare_decode_rect(np.asarray(seq_lens, np.int64), np.asarray(last_tokens), block_tables, block_size)
    tokens, positions, slots, ctx, tabs = [], [], [], [], []
    for L, tok, table in zip(seq_lens, last_tokens, block_tables):
        pos = L - 1
        tokens.append(int(tok)); positions.append(pos); ctx.append(L)
        slots.append(_slot(table, pos, block_size, "completion"))
        tabs.append(list(table[:used_blocks_for_len(L, block_size, len(table))]))
    return dict(is_prefill=False, tokens=np.asarray(tokens, np.uint32), positions=np.asarray(positions, np.int64),
                slot_mapping=np.asarray(slots, np.int64), context_lens=np.asarray(ctx, np.int32),
                block_tables=_pad_tables(tabs), max_context_len=int(max(ctx)))


def _prepare_decode_rect(lens: np.ndarray, toks: np.ndarray, tables: np.ndarray, block_size: int) -> dict:
    B, width = tables.shape
    if lens.shape != (B,) or toks.shape != (B,):
        raise BackendError(f"prepare_decode: {B} block tables for {lens.shape[0]} sequences / {toks.shape[0]} tokens")
    pos = lens - 1
    bi = pos // block_size
    if (bi >= width).any() or (pos < 0).any():
        b = int(np.argmax((bi >= width) | (pos < 0)))
        raise BackendError(f"Block table is too small (completion)! start_pos={int(pos[b])} block_size={block_size} table_len={width}")
    rows = np.arange(B)
    slots = tables[rows, bi].astype(np.int64) * block_size + pos % block_size
    used = np.minimum((lens + block_size - 1) // block_size, width)          # used_blocks_for_len
    w = max(1, int(used.max()))
    tabs = np.where(np.arange(w)[None, :] < used[:, None], tables[:, :w], 0).astype(np.int32)
    return dict(is_prefill=False, tokens=toks.astype(np.uint32), positions=pos.astype(np.int64), slot_mapping=slots,
                context_lens=lens.astype(np.int32), block_tables=tabs, max_context_len=int(lens.max()))


def prepare_prompt(prompts: Sequence[Sequence[int]], block_tables: Sequence[Sequence[int]], block_size: int,
                   num_cached_tokens: Optional[Sequence[int]] = None, chunk_size: int = PREFILL_CHUNK_SIZE) -> dict:
    """inputs.rs:90-374 for the chunked-prefill case: sequence i contributes prompt positions
    [cached_i, min(len_i, cached_i + chunk)); keys cover [0, end_i) and are read from the paged cache
    (``use_cached_kv`` :133-143)."""
    cached = list(num_cached_tokens) if num_cached_tokens is not None else [0] * len(prompts)
    tokens, positions, slots, tabs = [], [], [], []
    cu_q, cu_k = [0], [0]
    max_q = max_k = 0
    for prompt, table, c0 in zip(prompts, block_tables, cached):
        end = min(len(prompt), c0 + chunk_size)
        for p in range(c0, end):
            tokens.append(int(prompt[p])); positions.append(p)
            slots.append(_slot(table, p, block_size, "prompt"))
        cu_q.append(cu_q[-1] + (end - c0)); cu_k.append(cu_k[-1] + end)
        max_q, max_k = max(max_q, end - c0), max(max_k, end)
        tabs.append(list(table[:used_blocks_for_len(end, block_size, len(table))]))
    return dict(is_prefill=True, tokens=np.asarray(tokens, np.uint32), positions=np.asarray(positions, np.int64),
                slot_mapping=np.asarray(slots, np.int64), cu_seqlens_q=np.asarray(cu_q, np.int32),
                cu_seqlens_k=np.asarray(cu_k, np.int32), max_seqlen_q=max_q, max_seqlen_k=max_k,
                block_tables=_pad_tables(tabs), max_context_len=max_k)


def flashinfer_csr(lens: Sequence[int], block_tables: Sequence[Sequence[int]], block_size: int) -> dict:
    """The CSR page tables the reference builds for its flashinfer backend (inputs.rs:477-531): ``indptr`` [B + 1], ``indices`` (the used
    blocks of every sequence, back to back), ``last_len`` = (len - 1) % block_size + 1 (0 for an empty sequence) and the derived
    ``kv_len`` = (pages - 1) * block_size + last_len."""
    indptr, indices, last_len, kv_len = [0], [], [], []
    for n, table in zip(lens, block_tables):
        used = used_blocks_for_len(int(n), block_size, len(table))
        indices.extend(int(b) for b in table[:used])
        indptr.append(len(indices))
        last = 0 if n == 0 else (int(n) - 1) % block_size + 1
        last_len.append(last)
        pages = indptr[-1] - indptr[-2]
        kv_len.append(0 if pages == 0 else (pages - 1) * block_size + last)
    return dict(indptr=np.asarray(indptr, np.uint32), indices=np.asarray(indices, np.uint32), last_len=np.asarray(last_len, np.uint32),
                kv_len=np.asarray(kv_len, np.uint32))


def to_device(prep: dict, device="cuda"):
    """numpy metadata -> (tokens i64, positions i64, InputMetadata) on the device."""
    import torch
    from .backend import InputMetadata
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(device) if dt is None else torch.from_numpy(np.ascontiguousarray(a)).to(device).to(dt)
    meta = InputMetadata(is_prefill=prep["is_prefill"], slot_mapping=t(prep["slot_mapping"]), block_tables=t(prep["block_tables"]),
                         max_context_len=prep["max_context_len"])
    if prep["is_prefill"]:
        meta.cu_seqlens_q, meta.cu_seqlens_k = t(prep["cu_seqlens_q"]), t(prep["cu_seqlens_k"])
        meta.max_seqlen_q, meta.max_seqlen_k = prep["max_seqlen_q"], prep["max_seqlen_k"]
    else:
        meta.context_lens = t(prep["context_lens"])
    return t(prep["tokens"].astype(np.int64)), t(prep["positions"]), meta
