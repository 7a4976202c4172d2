"""``QLinear``: the reference's dispatch over quantised linears (/root/reference/src/openai/models/linear.rs:419-916).

GPTQ / AWQ tuple (scales present) -> ``gptq_matmul``; GGUF tensor -> ``QMatMul`` on f32-cast activations (``forward_no_dequant``,
f32 result); GGUF tensor stored transposed -> dequantise to f16 and a plain (cuBLAS) matmul (``forward_via_dequant``).
Shape handling follows the reference: ``[b, 1, d]`` decode inputs are flattened to ``[b, d]`` for the quantised GEMM and restored."""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import BackendError
from .backend import QMatMul, QTensor, dequantize
from .gptq import gptq_matmul


class QLinear:
    def __init__(self, inner, bias: Optional[torch.Tensor] = None, scales: Optional[torch.Tensor] = None,
                 qzeros: Optional[torch.Tensor] = None, g_idx: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                 group_size: int = 0, bits: int = 0, is_awq: bool = False, transposed_weight: bool = False):
        self.inner, self.bias, self.scales, self.qzeros, self.g_idx, self.workspace = inner, bias, scales, qzeros, g_idx, workspace
        self.group_size, self.bits, self.is_awq, self.transposed_weight = group_size, bits, is_awq, transposed_weight
        if scales is None and not isinstance(inner, QMatMul):
            raise BackendError("QLinear: a GGUF QLinear wraps a QMatMul; a GPTQ/AWQ one needs scales")

    @classmethod
    def from_qtensor(cls, w: QTensor, bias: Optional[torch.Tensor] = None, transposed_weight: bool = False) -> "QLinear":
        """``QLinear::new`` (linear.rs:498-521)"""
        return cls(QMatMul.from_arc(w), bias, transposed_weight=transposed_weight)

    @classmethod
    def from_gptq(cls, qweight: torch.Tensor, scales: torch.Tensor, qzeros, g_idx, workspace, group_size: int, bits: int,
                  is_awq: bool = False, bias: Optional[torch.Tensor] = None) -> "QLinear":
        """``QLinear::from_linear`` (linear.rs:523-540)"""
        return cls(qweight, bias, scales, qzeros, g_idx, workspace, group_size, bits, is_awq)

    def _add_bias(self, y: torch.Tensor) -> torch.Tensor:
        return y if self.bias is None else y + self.bias.to(y.dtype)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.scales is not None:                                      # linear.rs:855-905
            if x.dim() not in (2, 3, 4):
                raise BackendError("Invalid input format!")
            y = gptq_matmul(x, self.inner, self.scales, self.qzeros, self.g_idx, self.workspace, self.bits, self.group_size, self.is_awq)
            return self._add_bias(y)
        return self.forward_via_dequant(x) if self.transposed_weight else self.forward_no_dequant(x)

    def forward_no_dequant(self, x: torch.Tensor) -> torch.Tensor:
        """linear.rs:765-806: f32 in, f32 out; a single-token sequence dimension is squeezed for the GEMM and restored"""
        xs = x.float()
        squeeze = x.dim() in (3, 4) and x.shape[1] == 1
        if squeeze:
            xs = xs.reshape(x.shape[0], *x.shape[2:])
        y = self.inner.forward(xs)
        if squeeze:
            y = y.reshape(x.shape[0], 1, *y.shape[1:])
        return self._add_bias(y)

    def forward_via_dequant(self, x: torch.Tensor) -> torch.Tensor:
        """linear.rs:808-842: the tensor holds W^T ([in, out]); dequantise (-> f16 -> x.dtype) and multiply"""
        w = dequantize(self.inner.w).half().to(x.dtype)
        if w.shape[0] != x.shape[-1]:
            raise BackendError(f"QLinear: transposed weight {tuple(w.shape)} vs x {tuple(x.shape)}")
        return self._add_bias(x @ w)
