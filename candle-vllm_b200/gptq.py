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
    """``marlin_weight_repack`` (gptq.rs:356-359): u32 [K/8, N] -> u32 [K/16, 2N] (library-private layout)."""
    if qweight.dtype not in (torch.int32, torch.uint32):
        raise BackendError(f"MarlinRepack is only supported for i32/u32 weight ({qweight.dtype})")
    if not qweight.is_cuda:
        raise BackendError("no cpu support for MarlinRepack")
    if bits != 4 or is_awq:
        raise BackendError("marlin repack: only 4-bit GPTQ is supported in this round")
    require_device()
    kp, n = qweight.shape
    out = torch.empty((kp // 2, n * 2), dtype=qweight.dtype, device=qweight.device)
    with torch.cuda.device(qweight.device):
        lib().gptq_repack(_ptr(qweight.contiguous()), _ptr(out), C.c_int32(kp), C.c_int32(n), _stream(qweight.device))
    check("gptq_repack")
    return out


def gptq_matmul(x: torch.Tensor, qweight: torch.Tensor, scales: torch.Tensor, qzeros: Optional[torch.Tensor],
                g_idx: Optional[torch.Tensor], workspace: Optional[torch.Tensor], bits: int, group_size: int,
                is_awq: bool = False) -> torch.Tensor:
    """``gptq_matmul`` (gptq.rs:242-262) in marlin format: x [.., K] f16/bf16, qweight = marlin_weight_repack(...),
    scales = marlin_permute_scales(...).  Returns [.., N] in x.dtype."""
    if workspace is None:
        raise BackendError("workspace is required for marlin matmul!")
    if x.dtype not in (torch.float16, torch.bfloat16):
        raise BackendError("GPTQMatMul is only supported for f16/bf16 marlin matmul.")
    if scales.dtype != x.dtype:
        raise BackendError("scales must have the activation dtype (linear.rs:249-251)")
    if bits != 4 or is_awq:
        raise BackendError("only 4-bit GPTQ marlin is supported in this round")
    require_device()
    size_k = qweight.shape[0] * (32 // bits) * 2
    size_n = qweight.shape[1] // 2
    if x.shape[-1] != size_k:
        raise BackendError(f"shape mismatch: x {tuple(x.shape)} vs K = {size_k}")
    x2 = x.reshape(-1, size_k).contiguous()
    m = x2.shape[0]
    out = torch.empty((m, size_n), dtype=x.dtype, device=x.device)
    fn = lib().marlin_4bit_f16 if x.dtype == torch.float16 else lib().marlin_4bit_bf16
    with torch.cuda.device(x.device):
        fn(_ptr(x2), _ptr(qweight), _ptr(scales), _ptr(qzeros), _ptr(g_idx), _ptr(out), C.c_int32(m), C.c_int32(size_k),
           C.c_int32(size_n), _ptr(workspace), C.c_int32(group_size), _stream(x.device))
    check("gptq_matmul")
    return out.reshape(*x.shape[:-1], size_n)
