"""Multi-head latent attention over the paged latent cache: mirror of the attention-rs MLA calls made by ``MlaAttention::forward``
(/root/reference/src/openai/models/layers/mla_attention.rs:479-552) over the C ABI (``concat_and_cache_mla``, ``mla_paged_attention``)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import BackendError, check, lib, require_device
from .backend import _cuda, _dt, _ptr, _stream


def concat_and_cache_mla(ckv: torch.Tensor, k_pe: torch.Tensor, ckv_cache: torch.Tensor, kpe_cache: torch.Tensor, slot_mapping: torch.Tensor) -> None:
    """ckv [T, R], k_pe [T, P] -> caches [nb, bs, 1, R] / [nb, bs, 1, P] at slot_mapping (i64; negative = pad)."""
    for t, n in ((ckv, "ckv"), (k_pe, "k_pe"), (ckv_cache, "ckv_cache"), (kpe_cache, "kpe_cache"), (slot_mapping, "slot_mapping")):
        _cuda(t, n)
    require_device()
    if ckv.dtype != ckv_cache.dtype or k_pe.dtype != kpe_cache.dtype or ckv.dtype not in (torch.float16, torch.bfloat16):
        raise BackendError("concat_and_cache_mla: 16-bit tensors of the cache dtype expected")
    with torch.cuda.device(ckv.device):
        lib().concat_and_cache_mla(_ptr(ckv.contiguous()), _ptr(k_pe.contiguous()), _ptr(ckv_cache), _ptr(kpe_cache), _ptr(slot_mapping.contiguous()),
                                   C.c_int32(ckv.shape[0]), C.c_int32(ckv.shape[-1]), C.c_int32(k_pe.shape[-1]), C.c_int32(_dt(ckv)), _stream(ckv.device))
    check("concat_and_cache_mla")


def _mla(q_absorbed, q_pe, ckv_cache, kpe_cache, block_tables, context_lens, cu_seqlens_q, sm_scale):
    _cuda(q_absorbed, "q_absorbed"); require_device()
    rows, H, R = q_absorbed.shape
    P = q_pe.shape[-1]
    nb, bs = ckv_cache.shape[0], ckv_cache.shape[1]
    if q_absorbed.dtype != ckv_cache.dtype or q_absorbed.dtype not in (torch.float16, torch.bfloat16):
        raise BackendError("mla attention: q and the latent cache must share a 16-bit dtype")
    out = torch.empty((rows, H, R), dtype=q_absorbed.dtype, device=q_absorbed.device)
    B, W = block_tables.shape
    L = lib()
    need = int(L.mla_paged_decode_workspace_bytes(C.c_int32(B), C.c_int32(H), C.c_int32(W), C.c_int32(bs)))
    ws = torch.empty(need, dtype=torch.uint8, device=q_absorbed.device)
    with torch.cuda.device(q_absorbed.device):
        L.mla_paged_attention(_ptr(out), _ptr(q_absorbed.contiguous()), _ptr(q_pe.contiguous()), _ptr(ckv_cache), _ptr(kpe_cache),
                              _ptr(block_tables.to(torch.int32).contiguous()), _ptr(context_lens.to(torch.int32).contiguous()),
                              _ptr(None if cu_seqlens_q is None else cu_seqlens_q.to(torch.int32).contiguous()), C.c_int32(B), C.c_int32(rows), C.c_int32(H),
                              C.c_int32(R), C.c_int32(P), C.c_int32(bs), C.c_int32(W), C.c_int64(nb), C.c_float(sm_scale), C.c_int32(_dt(q_absorbed)),
                              _ptr(ws), C.c_size_t(need), _stream(q_absorbed.device))
    check("mla_paged_attention")
    return out


def mla_paged_decode(q_absorbed, q_pe, ckv_cache, kpe_cache, block_tables, context_lens, sm_scale: float) -> torch.Tensor:
    """attention_rs::mla::mla_paged_decode (mla_attention.rs:541-550): one query row per sequence -> [B, H, kv_lora_rank]."""
    return _mla(q_absorbed, q_pe, ckv_cache, kpe_cache, block_tables, context_lens, None, sm_scale)


def mla_paged_prefill(q_absorbed, q_pe, ckv_cache, kpe_cache, block_tables, context_lens, cu_seqlens_q, sm_scale: float) -> torch.Tensor:
    """attention_rs::mla::mla_paged_prefill (mla_attention.rs:527-537): causal, rows = the last positions of each sequence's context."""
    return _mla(q_absorbed, q_pe, ckv_cache, kpe_cache, block_tables, context_lens, cu_seqlens_q, sm_scale)
