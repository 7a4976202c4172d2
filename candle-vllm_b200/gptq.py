"""GPTQ / Marlin int4 linear: mirror of ``src/backend/gptq.rs`` (``gptq_matmul``, ``marlin_weight_repack``) and of the
host-side preparation in ``src/openai/models/linear.rs:300-413`` (repack + ``marlin_permute_scales`` + workspace)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import BackendError, check, lib, require_device
from .backend import _ptr, _stream


def get_scale_perms():
    """linear.rs:341-352"""
    scale_perm = [i + 8 * j for i in range(8) for j in range(8)]
    scale_perm_single = [2 * i + j for i in range(4) for j in (0, 1, 8, 9, 16, 17, 24, 25)]
    return scale_perm, scale_perm_single


def marlin_permute_scales(s: torch.Tensor, size_k: int, size_n: int, group_size: int) -> torch.Tensor:
    """linear.rs:354-379 (host-side tensor op, as in the reference)."""
    scale_perm, scale_perm_single = get_scale_perms()
    if group_size != -1 and group_size < size_k:
        s = s.reshape(-1, len(scale_perm))[:, torch.tensor(scale_perm, device=s.device)]
    else:
        s = s.reshape(-1, len(scale_perm_single))[:, torch.tensor(scale_perm_single, device=s.device)]
    return s.reshape(-1, size_n).contiguous()


def marlin_weight_repack(qweight: torch.Tensor, bits: int = 4, is_awq: bool = False) -> torch.Tensor:
    """``marlin_weight_repack`` (gptq.rs:356-359): GPTQ u32 [K/8, N] or AWQ u32 [K, N/8] -> u32 [K/16, 2N]
    (library-private layout, same word count)."""
    if qweight.dtype not in (torch.int32, torch.uint32):
        raise BackendError(f"MarlinRepack is only supported for i32/u32 weight ({qweight.dtype})")
    if not qweight.is_cuda:
        raise BackendError("no cpu support for MarlinRepack")
    if bits != 4:
        raise BackendError("marlin repack: only 4-bit weights")
    require_device()
    d0, d1 = qweight.shape
    q = qweight.contiguous()
    with torch.cuda.device(qweight.device):
        if is_awq:                                           # [K, N/8] -> out_shape [K/8/2, N/8*8*2] (gptq.rs:283-289)
            out = torch.empty((d0 // 8 // 2, d1 * 8 * 2), dtype=qweight.dtype, device=qweight.device)
            lib().awq_repack(_ptr(q), _ptr(out), C.c_int32(d0), C.c_int32(d1), C.c_int32(bits), _stream(qweight.device))
            check("awq_repack")
        else:
            out = torch.empty((d0 // 2, d1 * 2), dtype=qweight.dtype, device=qweight.device)
            lib().gptq_repack(_ptr(q), _ptr(out), C.c_int32(d0), C.c_int32(d1), _stream(qweight.device))
            check("gptq_repack")
    return out


def marlin_checkpoint_repack(b: torch.Tensor, size_k: int, size_n: int) -> torch.Tensor:
    """Weights of a checkpoint that is ALREADY in Marlin format (``B`` u32 [K/16, 2N], linear.rs:219-251) -> the layout ``gptq_matmul`` reads
    (what ``marlin_weight_repack`` would have produced from the GPTQ tensors), same shape."""
    require_device()
    if b.dtype not in (torch.int32, torch.uint32) or tuple(b.shape) != (size_k // 16, size_n * 2):
        raise BackendError(f"marlin B tensor must be u32 [{size_k // 16}, {size_n * 2}], got {tuple(b.shape)} {b.dtype}")
    out = torch.empty_like(b)
    scratch = torch.empty((size_k // 8, size_n), dtype=torch.int32, device=b.device)
    with torch.cuda.device(b.device):
        lib().marlin_checkpoint_repack(_ptr(b.contiguous()), _ptr(out), _ptr(scratch), C.c_int32(size_k), C.c_int32(size_n), _stream(b.device))
    check("marlin_checkpoint_repack")
    return out


def gptq_matmul(x: torch.Tensor, qweight: torch.Tensor, scales: torch.Tensor, qzeros: Optional[torch.Tensor],
                g_idx: Optional[torch.Tensor], workspace: Optional[torch.Tensor], bits: int, group_size: int,
                is_awq: bool = False) -> torch.Tensor:
    """``gptq_matmul`` (gptq.rs:242-262).  With a ``workspace`` (marlin format): x [.., K] f16/bf16, qweight =
    marlin_weight_repack(...), scales = marlin_permute_scales(...), and for AWQ ``qzeros`` in the layout of the reference's
    converter (examples/convert_awq_marlin.py).  Without one: conventional GPTQ (qweight [K/pack, N], qzeros, g_idx) through
    ``gemm_half_q_half_alt``, f16 only -- the same split as GPTQMatMul::cuda_fwd_t (gptq.rs:46-49, :102-197)."""
    if scales.dtype != x.dtype:
        raise BackendError("scales must have the activation dtype (linear.rs:249-251)")
    require_device()
    pack = 32 // bits
    if workspace is None:
        if x.dtype != torch.float16:
            raise BackendError("GPTQMatMul is only supported for f16 non-marlin matmul. Use '--dtype f16' parameter instead.")
        if qzeros is None or g_idx is None:
            raise BackendError("conventional GPTQ matmul needs qzeros and g_idx (workspace is required for marlin matmul!)")
        size_k, size_n = qweight.shape[0] * pack, qweight.shape[1]
        if x.shape[-1] != size_k:
            raise BackendError(f"shape mismatch: x {tuple(x.shape)} vs K = {size_k}")
        x2 = x.reshape(-1, size_k).contiguous()
        out = torch.empty((x2.shape[0], size_n), dtype=x.dtype, device=x.device)
        gi = g_idx.to(torch.int32).contiguous()
        with torch.cuda.device(x.device):
            lib().gemm_half_q_half_alt(_ptr(x2), _ptr(qweight.contiguous()), _ptr(qzeros.contiguous()), _ptr(scales.contiguous()), _ptr(gi),
                                       _ptr(out), C.c_int32(x2.shape[0]), C.c_int32(size_n), C.c_int32(size_k), C.c_int32(bits),
                                       _stream(x.device))
        check("gptq_matmul")
        return out.reshape(*x.shape[:-1], size_n)
    if x.dtype not in (torch.float16, torch.bfloat16):
        raise BackendError("GPTQMatMul is only supported for f16/bf16 marlin matmul.")
    if bits != 4:
        raise BackendError("marlin matmul: only 4-bit weights")
    if is_awq and qzeros is None:
        raise BackendError("AWQ marlin matmul needs qzeros (marlin zero-point layout)")
    size_k = qweight.shape[0] * pack * 2
    size_n = qweight.shape[1] // 2
    if x.shape[-1] != size_k:
        raise BackendError(f"shape mismatch: x {tuple(x.shape)} vs K = {size_k}")
    x2 = x.reshape(-1, size_k).contiguous()
    m = x2.shape[0]
    out = torch.empty((m, size_n), dtype=x.dtype, device=x.device)
    L = lib()
    if is_awq:
        fn = L.marlin_awq_4bit_f16 if x.dtype == torch.float16 else L.marlin_awq_4bit_bf16
    else:
        fn = L.marlin_4bit_f16 if x.dtype == torch.float16 else L.marlin_4bit_bf16
    with torch.cuda.device(x.device):
        fn(_ptr(x2), _ptr(qweight), _ptr(scales), _ptr(qzeros), _ptr(g_idx), _ptr(out), C.c_int32(m), C.c_int32(size_k),
           C.c_int32(size_n), _ptr(workspace), C.c_int32(group_size), _stream(x.device))
    check("gptq_matmul")
    return out.reshape(*x.shape[:-1], size_n)
