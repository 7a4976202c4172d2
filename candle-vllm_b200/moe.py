"""Fused mixture-of-experts on GGUF expert tensors: mirror of ``FusedMoe`` (GGUF flavour,
/root/reference/src/openai/models/layers/moe.rs:1429-1482; quantized_qwen3_moe.rs:70-141) over the C ABI
(``topk_softmax``, ``sort_expert_assignments``, ``moe_gemm_gguf``)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from ._lib import BackendError, check, lib, require_device
from .backend import DType, GgmlType, _cuda, _ptr, _stream


def topk_softmax(router_logits: torch.Tensor, topk: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """attention_rs::topk::topk_softmax: f32 [T, E] -> (weights f32 [T, k], ids u32-as-i32 [T, k])."""
    _cuda(router_logits, "router_logits"); require_device()
    if router_logits.dtype != torch.float32 or router_logits.dim() != 2:
        raise BackendError("topk_softmax expects f32 [tokens, experts]")
    T, E = router_logits.shape
    w = torch.empty((T, topk), dtype=torch.float32, device=router_logits.device)
    ids = torch.empty((T, topk), dtype=torch.int32, device=router_logits.device)
    with torch.cuda.device(router_logits.device):
        lib().topk_softmax(_ptr(router_logits.contiguous()), _ptr(w), _ptr(ids), C.c_int32(T), C.c_int32(E), C.c_int32(topk), _stream(router_logits.device))
    check("topk_softmax")
    return w, ids


def sort_expert_assignments(topk_ids: torch.Tensor, num_experts: int, is_prefill: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """moe.rs:35-45: (expert_ids ascending, sorted_token_ids = index into the flattened [T * k] pairs)."""
    _cuda(topk_ids, "topk_ids"); require_device()
    flat = topk_ids.reshape(-1).contiguous()
    e = torch.empty_like(flat); s = torch.empty_like(flat)
    with torch.cuda.device(flat.device):
        lib().sort_expert_assignments(_ptr(flat), _ptr(e), _ptr(s), C.c_int32(flat.numel()), C.c_int32(num_experts), _stream(flat.device))
    check("sort_expert_assignments")
    return e, s


def moe_gemm_gguf(xs: torch.Tensor, experts: torch.Tensor, ggml_type: int, shape, topk_weights: Optional[torch.Tensor],
                  sorted_token_ids: torch.Tensor, expert_ids: torch.Tensor, topk: int, is_prefill: bool = False) -> torch.Tensor:
    """attention_rs::moe::moe_gemm_gguf: xs f32 [T or T*k, K]; experts = stacked GGML blocks [E, N, K] (u8); -> f32 [T*k, N]."""
    _cuda(xs, "xs"); require_device()
    E, N, K = (int(v) for v in shape)
    if xs.dtype != torch.float32 or xs.dim() != 2 or xs.shape[1] != K:
        raise BackendError(f"moe_gemm_gguf: xs must be f32 [rows, {K}]")
    be, bb = GgmlType.BLOCK[ggml_type]
    if experts.dtype != torch.uint8 or experts.numel() != E * N * (K // be) * bb:
        raise BackendError("moe_gemm_gguf: experts must be the stacked GGML blocks [E, N, K]")
    P = sorted_token_ids.numel()
    out = torch.empty((P, N), dtype=torch.float32, device=xs.device)
    L = lib()
    need = int(L.moe_gemm_workspace_bytes(C.c_int32(P), C.c_int32(N), C.c_int32(K), C.c_int32(E)))
    ws = torch.empty(need + 256, dtype=torch.uint8, device=xs.device)
    off = (-ws.data_ptr()) % 256
    tw = None if topk_weights is None else topk_weights.reshape(-1).contiguous()
    with torch.cuda.device(xs.device):
        L.moe_gemm_gguf(_ptr(xs.contiguous()), _ptr(experts), _ptr(tw), _ptr(sorted_token_ids), _ptr(expert_ids), _ptr(out), C.c_int32(E), C.c_int32(topk),
                        C.c_int32(xs.shape[0]), C.c_int32(P), C.c_int32(N), C.c_int32(K), C.c_int32(ggml_type), C.c_int32(1 if is_prefill else 0),
                        C.c_void_p(ws.data_ptr() + off), C.c_size_t(need), _stream(xs.device))
    check("moe_gemm_gguf")
    return out


def moe_gemm_fp8(xs: torch.Tensor, experts: torch.Tensor, scale: torch.Tensor, topk_weights: Optional[torch.Tensor], sorted_token_ids: torch.Tensor,
                 expert_ids: torch.Tensor, topk: int, block_y: int = 128, block_x: int = 128, is_prefill: bool = False) -> torch.Tensor:
    """attention_rs::moe::moe_gemm_fp8 (moe.rs:1447-1473): xs f16 / bf16 [T or T*k, K]; experts e4m3 (or u8) [E, N, K]; scale f32
    [E, ceil(N/by), ceil(K/bx)]; -> [T*k, N] in xs' dtype."""
    _cuda(xs, "xs"); require_device()
    if experts.dim() != 3 or experts.element_size() != 1:
        raise BackendError("moe_gemm_fp8: experts must be an 8-bit [E, N, K] tensor")
    E, N, K = (int(v) for v in experts.shape)
    if xs.dtype not in (torch.float16, torch.bfloat16) or xs.dim() != 2 or xs.shape[1] != K:
        raise BackendError(f"moe_gemm_fp8: xs must be f16 / bf16 [rows, {K}]")
    want = (E, -(-N // block_y), -(-K // block_x))
    if scale.dtype != torch.float32 or tuple(scale.shape) != want:
        raise BackendError(f"moe_gemm_fp8: scale must be f32 {want}, got {tuple(scale.shape)} {scale.dtype}")
    P = sorted_token_ids.numel()
    out = torch.empty((P, N), dtype=xs.dtype, device=xs.device)
    L = lib()
    need = int(L.moe_gemm_fp8_workspace_bytes(C.c_int32(P), C.c_int32(N), C.c_int32(K), C.c_int32(E)))
    ws = torch.empty(need + 256, dtype=torch.uint8, device=xs.device)
    off = (-ws.data_ptr()) % 256
    tw = None if topk_weights is None else topk_weights.reshape(-1).float().contiguous()
    ex = experts.contiguous().view(torch.uint8)
    with torch.cuda.device(xs.device):
        L.moe_gemm_fp8(_ptr(xs.contiguous()), _ptr(ex), _ptr(scale.contiguous()), _ptr(tw), _ptr(sorted_token_ids), _ptr(expert_ids), _ptr(out),
                       C.c_int32(E), C.c_int32(topk), C.c_int32(xs.shape[0]), C.c_int32(P), C.c_int32(N), C.c_int32(K), C.c_int32(block_y),
                       C.c_int32(block_x), C.c_int32(DType.F16 if xs.dtype == torch.float16 else DType.BF16), C.c_int32(1 if is_prefill else 0),
                       C.c_void_p(ws.data_ptr() + off), C.c_size_t(need), _stream(xs.device))
    check("moe_gemm_fp8")
    return out


class FusedMoe:
    """GGUF fused MoE block: router ``gate`` f32 [E, H]; ``gate_experts`` / ``up_experts`` [E, I, H] and ``down_experts`` [E, H, I] as
    stacked GGML blocks (moe.rs:1429-1482).  forward(xs f32 [T, H]) -> f32 [T, H]."""

    def __init__(self, gate: torch.Tensor, gate_experts: torch.Tensor, up_experts: torch.Tensor, down_experts: torch.Tensor,
                 ggml_types: Tuple[int, int, int], num_experts: int, hidden: int, inter: int, num_experts_per_tok: int,
                 norm_topk_prob: bool = True, routed_scaling_factor: Optional[float] = None):
        self.gate = _cuda(gate, "gate").float().contiguous()
        self.ge, self.ue, self.de = gate_experts, up_experts, down_experts
        self.types = ggml_types
        self.E, self.H, self.I, self.k = num_experts, hidden, inter, num_experts_per_tok
        self.norm_topk_prob, self.routed_scaling_factor = norm_topk_prob, routed_scaling_factor

    def forward(self, xs: torch.Tensor, is_prefill: bool = False) -> torch.Tensor:
        T = xs.shape[0]
        # router: a [T, H] x [H, E] f32 product on the plumbing library (the reference's `gate` Linear, moe.rs:1435); every expert
        # GEMM below runs on this backend's kernels
        router_logits = xs.float() @ self.gate.t()
        w, ids = topk_softmax(router_logits, self.k)
        if self.norm_topk_prob:
            w = w / w.sum(dim=-1, keepdim=True)
        if self.routed_scaling_factor is not None:
            w = w * self.routed_scaling_factor
        expert_ids, sorted_token_ids = sort_expert_assignments(ids, self.E, is_prefill)
        g = moe_gemm_gguf(xs, self.ge, self.types[0], (self.E, self.I, self.H), None, sorted_token_ids, expert_ids, self.k, is_prefill)
        u = moe_gemm_gguf(xs, self.ue, self.types[1], (self.E, self.I, self.H), None, sorted_token_ids, expert_ids, self.k, is_prefill)
        down_in = torch.nn.functional.silu(g) * u                                   # (up * gate.apply(act)), moe.rs:1462
        ys = moe_gemm_gguf(down_in, self.de, self.types[2], (self.E, self.H, self.I), w, sorted_token_ids, expert_ids, self.k, is_prefill)
        return ys.reshape(T, self.k, self.H).sum(dim=1)
