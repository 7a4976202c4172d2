"""Operator-level mirror of the reference's backend surface, over the C ABI.

Names, argument meaning and error behaviour follow the reference so the parity tests read like
tests of the reference itself:

  copy_blocks(key_caches, value_caches, block_mapping)        src/backend/cache.rs:15-165
  swap_blocks(src, dst, mapping)                              attention_rs::cache::swap_blocks
                                                              (call site cache_engine.rs:527-535)
  InputMetadata / PagedAttention::{new, forward}              attention-rs; call sites
                                                              layers/attention.rs:566-575,707-718
  QTensor / QMatMul::{from_arc, forward}                      candle; models/linear.rs:765-806

Tensors are torch CUDA tensors (device memory + stream plumbing only); every op launches a
hand-written sm_100a kernel through libb200backend.so.  Errors -> ``BackendError`` (``bail!``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from ._lib import BackendError, check, lib, require_device


class DType:
    F32, F16, BF16, U8, FP8_E4M3, F16_K4 = 0, 1, 2, 3, 4, 5


class GgmlType:
    Q8_0, Q4_K, Q6_K = 8, 12, 14
    BLOCK = {8: (32, 34), 12: (256, 144), 14: (256, 210)}     # elems, bytes


class KvLayout:
    FLASH, PAGED = 0, 1


_TORCH2B200 = {torch.float32: DType.F32, torch.float16: DType.F16, torch.bfloat16: DType.BF16,
               torch.uint8: DType.U8}


def _dt(t: torch.Tensor) -> int:
    try:
        return _TORCH2B200[t.dtype]
    except KeyError:
        raise BackendError(f"unsupported dtype {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(dev=None) -> C.c_int64:
    return C.c_int64(torch.cuda.current_stream(dev).cuda_stream)


def _cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise BackendError(f"Expected `{name}` to be on a CUDA device.")
    return t


# ---------------------------------------------------------------------------------------------
# copy_blocks / swap_blocks
# ---------------------------------------------------------------------------------------------
def _pairs(block_mapping) -> np.ndarray:
    pairs = []
    if isinstance(block_mapping, dict):
        for s, ds in block_mapping.items():
            if isinstance(ds, (list, tuple)):
                pairs += [(int(s), int(d)) for d in ds]
            else:
                pairs.append((int(s), int(ds)))
    else:
        pairs = [(int(s), int(d)) for s, d in block_mapping]
    return np.asarray(pairs, np.int64).reshape(-1, 2)


def copy_blocks(key_caches: Sequence[torch.Tensor], value_caches: Sequence[torch.Tensor],
                block_mapping: Union[Dict[int, List[int]], Sequence]) -> None:
    """``backend::copy_blocks`` (cache.rs:15-165).  block_mapping: src -> [dst, ...]."""
    if len(key_caches) == 0:
        return                                                    # cache.rs:41-44
    k0, v0 = key_caches[0], value_caches[0]
    _cuda(k0, "key caches")
    if k0.device != v0.device:
        raise BackendError(f"`key` and `value` caches have different devices, got {k0.device} and {v0.device} respectively.")
    if k0.dtype != v0.dtype:
        raise BackendError(f"Key and value caches have different types, got {k0.dtype} and {v0.dtype}.")
    fn = {torch.bfloat16: "copy_blocks_bf16", torch.float16: "copy_blocks_f16", torch.float32: "copy_blocks_f32",
          torch.uint8: "copy_blocks_u8"}.get(k0.dtype)
    if fn is None:
        raise BackendError("only f32, f16, bf16 (and u8 for FP8 KV) input data type supported!")
    require_device()
    kptrs = np.asarray([k.data_ptr() for k in key_caches], np.uint64)
    vptrs = np.asarray([v.data_ptr() for v in value_caches], np.uint64)
    pairs = _pairs(block_mapping)
    numel_per_block = int(np.prod(k0.shape[1:]))
    with torch.cuda.device(k0.device):
        getattr(lib(), fn)(kptrs.ctypes.data_as(C.c_void_p), vptrs.ctypes.data_as(C.c_void_p),
                           pairs.ctypes.data_as(C.c_void_p), C.c_int32(len(key_caches)), C.c_int32(len(pairs)),
                           C.c_int32(numel_per_block), _stream(k0.device))
    check("copy_blocks")


def swap_blocks(src: torch.Tensor, dst: torch.Tensor, mapping: Dict[int, int]) -> None:
    """``attention_rs::cache::swap_blocks`` (cache_engine.rs:527-535): dst[d] <- src[s]; either side
    may be a CPU tensor (GPU cache <-> CPU swap tier)."""
    if not (src.is_cuda or dst.is_cuda):
        raise BackendError("swap_blocks: at least one of src/dst must be on a CUDA device")
    if src.dtype != dst.dtype or src.shape[1:] != dst.shape[1:]:
        raise BackendError(f"swap_blocks: src/dst block shape or dtype differ ({src.shape}, {dst.shape})")
    if not (src.is_contiguous() and dst.is_contiguous()):
        raise BackendError("swap_blocks: tensors must be contiguous")
    require_device()
    pairs = _pairs(mapping)
    if len(pairs) and (pairs[:, 0].max() >= src.shape[0] or pairs[:, 1].max() >= dst.shape[0] or pairs.min() < 0):
        raise BackendError("swap_blocks: block id out of range")
    bytes_per_block = int(np.prod(src.shape[1:])) * src.element_size()
    dev = src.device if src.is_cuda else dst.device
    with torch.cuda.device(dev):
        lib().swap_blocks(_ptr(src), _ptr(dst), pairs.ctypes.data_as(C.c_void_p), C.c_int32(len(pairs)),
                          C.c_int64(bytes_per_block), _stream(dev))
    check("swap_blocks")


# ---------------------------------------------------------------------------------------------
# reshape_and_cache + paged attention
# ---------------------------------------------------------------------------------------------
def _kv_layout(k_cache: torch.Tensor) -> int:
    if k_cache.dim() == 4:
        return KvLayout.FLASH          # [nb, bs, kvh, hd]
    if k_cache.dim() == 5:
        return KvLayout.PAGED          # [nb, kvh, hd/x, bs, x]
    raise BackendError(f"unexpected key cache rank {k_cache.dim()}")


def reshape_and_cache(key: torch.Tensor, value: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                      slot_mapping: torch.Tensor, fp8: bool = False) -> None:
    """key/value [T, kvh, hd] -> caches at ``slot_mapping`` (i64, -1 = pad)."""
    _cuda(key, "key"); _cuda(k_cache, "key cache")
    if slot_mapping.dtype != torch.int64:
        raise BackendError("slot_mapping must be i64")
    T, kvh, hd = key.shape
    layout = _kv_layout(k_cache)
    bs = k_cache.shape[1] if layout == KvLayout.FLASH else k_cache.shape[3]
    if key.stride(2) != 1 or value.stride(2) != 1 or key.stride(1) != hd or value.stride(1) != hd:
        key, value = key.contiguous(), value.contiguous()
    cache_dt = DType.FP8_E4M3 if (fp8 or k_cache.dtype == torch.uint8) else _dt(k_cache)
    require_device()
    with torch.cuda.device(key.device):
        lib().reshape_and_cache(_ptr(key), _ptr(value), _ptr(k_cache), _ptr(v_cache), _ptr(slot_mapping),
                                C.c_int32(T), C.c_int32(kvh), C.c_int32(hd), C.c_int32(bs),
                                C.c_int64(key.stride(0)), C.c_int64(value.stride(0)), C.c_int32(_dt(key)),
                                C.c_int32(cache_dt), C.c_int32(layout), _stream(key.device))
    check("reshape_and_cache")


@dataclass
class InputMetadata:
    """Mirror of attention-rs ``InputMetadata`` as built at inputs.rs:351-367 / :552-568."""
    is_prefill: bool
    slot_mapping: torch.Tensor                       # i64 [T]
    block_tables: Optional[torch.Tensor] = None      # u32/i32 [B, max_blocks], 0-padded
    context_lens: Optional[torch.Tensor] = None      # u32/i32 [B], includes the decoded token
    cu_seqlens_q: Optional[torch.Tensor] = None      # u32/i32 [n+1]
    cu_seqlens_k: Optional[torch.Tensor] = None
    max_seqlen_q: int = 0
    max_seqlen_k: int = 0
    max_context_len: int = 0
    is_mla: bool = False
    sequence_ids: Optional[list] = None
    mamba_slot_mapping: Optional[torch.Tensor] = None
    seqlens: Optional[list] = None
    flashinfer_metadata: Optional[object] = None
    is_mtp_verify: bool = False


@dataclass
class FlashInferMetadata:
    """The CSR page-table encoding of the reference's flashinfer build (``FlashInferMetadata``, inputs.rs:477-549): ``indptr`` u32 [B + 1],
    ``indices`` u32 [nnz] physical block ids sequence by sequence, ``last_len`` u32 [B] tokens in each sequence's last page."""
    indptr: torch.Tensor
    indices: torch.Tensor
    last_len: torch.Tensor
    max_blocks_per_seq: int = 0

    def to_paged(self, block_size: int):
        """-> (block_tables i32 [B, W] 0-padded, context_lens i32 [B]) on the device, through ``flashinfer_csr_to_paged`` (no host sync;
        W = max_blocks_per_seq, which the host knows from its own tables -- graph replay uses the static width)."""
        require_device()
        B = self.last_len.numel()
        W = int(self.max_blocks_per_seq)
        if W <= 0:
            raise BackendError("FlashInferMetadata.max_blocks_per_seq must be set (static table width)")
        dev = self.indptr.device
        bt = torch.empty((B, W), dtype=torch.int32, device=dev)
        cl = torch.empty((B,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            lib().flashinfer_csr_to_paged(_ptr(self.indptr), _ptr(self.indices), _ptr(self.last_len), _ptr(bt), _ptr(cl), C.c_int32(B), C.c_int32(W),
                                          C.c_int32(block_size), _stream(dev))
        check("flashinfer_csr_to_paged")
        return bt, cl


class PagedAttention:
    """``PagedAttention::new(num_heads, head_dim, scale, num_kv_heads, sliding_window, device, alibi,
    fp8_kvcache)`` / ``.forward(q, k, v, mask, k_cache, v_cache, &meta, softcap)`` (attention-rs;
    call sites layers/attention.rs:566-575,707-718,888-897,983-994)."""

    def __init__(self, num_attention_heads: int, head_dim: int, scale: float,
                 num_key_value_heads: Optional[int] = None, sliding_window: Optional[int] = None,
                 device=None, alibi_slopes=None, fp8_kvcache: bool = False):
        if alibi_slopes is not None:
            raise BackendError("alibi slopes are not supported (no in-tree caller passes them)")
        self.num_attention_heads = num_attention_heads
        self.head_dim = head_dim
        self.scale = float(scale)
        self.num_key_value_heads = num_key_value_heads or num_attention_heads
        if num_attention_heads % self.num_key_value_heads:
            raise BackendError("num_attention_heads must be divisible by num_key_value_heads")
        self.sliding_window = sliding_window
        self.fp8_kvcache = fp8_kvcache
        self.device = device
        self._ws = None

    def _workspace(self, B: int, max_blocks: int, bs: int, dev) -> torch.Tensor:
        need = lib().paged_attention_decode_workspace_bytes(C.c_int32(B), C.c_int32(self.num_attention_heads),
                                                            C.c_int32(self.head_dim), C.c_int32(max_blocks), C.c_int32(bs))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.zeros(int(need), dtype=torch.uint8, device=dev)
        return self._ws

    def forward(self, query: torch.Tensor, key: Optional[torch.Tensor], value: Optional[torch.Tensor],
                attention_mask, key_cache: Optional[torch.Tensor], value_cache: Optional[torch.Tensor],
                input_metadata: InputMetadata, softcapping: Optional[float] = None,
                out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """q [T, heads, hd], k/v [T, kvh, hd] (model dtype).  Writes k/v into the cache at
        ``slot_mapping`` first (K3), then attends over the paged cache (K1 decode / K2 prefill).
        Returns [T, heads, hd]."""
        _cuda(query, "query")
        if key_cache is None or value_cache is None:
            raise BackendError("PagedAttention.forward requires a KV cache on this backend")
        if query.dtype not in (torch.bfloat16, torch.float16):
            raise BackendError(f"PagedAttention: query dtype {query.dtype} unsupported (bf16/f16)")
        require_device()
        T, H, hd = query.shape
        if H != self.num_attention_heads or hd != self.head_dim:
            raise BackendError(f"query shape {tuple(query.shape)} does not match heads={self.num_attention_heads} head_dim={self.head_dim}")
        meta = input_metadata
        fp8 = self.fp8_kvcache or key_cache.dtype == torch.uint8
        if key is not None and value is not None:
            reshape_and_cache(key, value, key_cache, value_cache, meta.slot_mapping, fp8=fp8)
        layout = _kv_layout(key_cache)
        bs = key_cache.shape[1] if layout == KvLayout.FLASH else key_cache.shape[3]
        nb = key_cache.shape[0]
        q = query.contiguous()
        odt = out_dtype or query.dtype
        out = torch.empty((T, H, hd), dtype=odt, device=query.device)
        bt = meta.block_tables
        if bt is None and meta.flashinfer_metadata is not None and not meta.is_prefill:
            # the reference's default (flashinfer) build hands CSR page tables instead (inputs.rs:477-506): expand them on the device
            fm = meta.flashinfer_metadata
            bt, cl_csr = fm.to_paged(bs)
            meta = InputMetadata(False, meta.slot_mapping, bt, cl_csr)
        if bt is None:
            raise BackendError("InputMetadata.block_tables (or flashinfer_metadata for decode) is required")
        if bt.dtype not in (torch.int32, torch.uint32) or not bt.is_contiguous():
            raise BackendError("block_tables must be contiguous u32/i32")
        cache_dt = DType.FP8_E4M3 if fp8 else _dt(key_cache)
        win = int(self.sliding_window or 0)
        cap = float(softcapping or 0.0)
        L = lib()
        with torch.cuda.device(query.device):
            if meta.is_prefill:
                cq, ck = meta.cu_seqlens_q, meta.cu_seqlens_k
                if cq is None or ck is None:
                    raise BackendError("prefill requires cu_seqlens_q / cu_seqlens_k")
                n = cq.numel() - 1
                L.paged_attention_prefill(_ptr(out), _ptr(q), _ptr(key_cache), _ptr(value_cache), _ptr(bt), _ptr(cq), _ptr(ck),
                                          C.c_int32(n), C.c_int32(T), C.c_int32(meta.max_seqlen_q), C.c_int32(H),
                                          C.c_int32(self.num_key_value_heads), C.c_int32(hd), C.c_int32(bs),
                                          C.c_int32(bt.shape[1]), C.c_float(self.scale), C.c_float(cap), C.c_int32(win),
                                          C.c_int32(_dt(q)), C.c_int32(cache_dt), C.c_int32(layout), _stream(query.device))
            else:
                cl = meta.context_lens
                if cl is None:
                    raise BackendError("decode requires context_lens")
                ws = self._workspace(T, bt.shape[1], bs, query.device)
                L.paged_attention_decode(_ptr(out), _ptr(q), _ptr(key_cache), _ptr(value_cache), _ptr(bt), _ptr(cl),
                                         C.c_int32(T), C.c_int32(H), C.c_int32(self.num_key_value_heads), C.c_int32(hd),
                                         C.c_int32(bs), C.c_int32(bt.shape[1]), C.c_int64(nb), C.c_float(self.scale),
                                         C.c_float(cap), C.c_int32(win), C.c_int32(_dt(q)), C.c_int32(cache_dt),
                                         C.c_int32(layout), C.c_int32(_TORCH2B200[odt]), _ptr(ws), C.c_size_t(ws.numel()),
                                         _stream(query.device))
        check("PagedAttention.forward")
        return out


# ---------------------------------------------------------------------------------------------
# QTensor / QMatMul
# ---------------------------------------------------------------------------------------------
class QTensor:
    """GGML-quantised weight [n, k]: verbatim GGUF bytes on the device (candle ``QTensor``)."""

    def __init__(self, data: torch.Tensor, ggml_type: int, shape, allow_cpu: bool = False):
        n, k = int(shape[0]), int(shape[1])
        if ggml_type not in GgmlType.BLOCK:
            raise BackendError(f"unsupported ggml type {ggml_type}")
        be, bb = GgmlType.BLOCK[ggml_type]
        if k % be:
            raise BackendError(f"k={k} is not a multiple of the block size {be}")
        if data.dtype != torch.uint8 or data.numel() != n * (k // be) * bb:
            raise BackendError(f"QTensor: expected {n * (k // be) * bb} bytes (u8), got {data.numel()} {data.dtype}")
        # host-resident QTensors exist only for shard bookkeeping (tensor-parallel splitting); no op accepts them
        self.data = (data if allow_cpu else _cuda(data, "QTensor data")).contiguous()
        self.ggml_type = ggml_type
        self.shape = (n, k)

    @classmethod
    def from_numpy(cls, blocks: np.ndarray, ggml_type: int, shape, device="cuda"):
        return cls(torch.from_numpy(np.ascontiguousarray(blocks, np.uint8).reshape(-1)).to(device), ggml_type, shape)

    def dequantize(self) -> torch.Tensor:
        return dequantize(self)


def dequantize(w: QTensor) -> torch.Tensor:
    require_device()
    n, k = w.shape
    out = torch.empty((n, k), dtype=torch.float32, device=w.data.device)
    with torch.cuda.device(w.data.device):
        lib().dequantize_f32(_ptr(w.data), _ptr(out), C.c_int64(n), C.c_int64(k), C.c_int32(w.ggml_type), _stream(w.data.device))
    check("dequantize")
    return out


class QMatMul:
    """``QMatMul::from_arc(qtensor)`` / ``.forward(x)``: y = x . dequant(W)^T, f32 in -> f32 out
    (linear.rs:765-806; QuantizedAttention attention.rs:920-922)."""

    def __init__(self, qtensor: QTensor):
        self.w = qtensor
        self._ws = None

    @classmethod
    def from_arc(cls, qtensor: QTensor) -> "QMatMul":
        return cls(qtensor)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _cuda(x, "x")
        n, k = self.w.shape
        if x.shape[-1] != k:
            raise BackendError(f"QMatMul: shape mismatch, x {tuple(x.shape)} vs weight {self.w.shape}")
        if x.dtype not in (torch.float32, torch.float16):
            raise BackendError("QMatMul.forward expects f32 (or pre-cast f16) activations (linear.rs:769-798)")
        require_device()
        lead = x.shape[:-1]
        x2 = x.reshape(-1, k).contiguous()
        m = x2.shape[0]
        y = torch.empty((m, n), dtype=torch.float32, device=x.device)
        L = lib()
        with torch.cuda.device(x.device):
            if x2.dtype == torch.float16:
                xk4 = torch.empty_like(x2)          # the GEMM consumes fp16 activations in K4 order (b200_backend.h)
                L.cast(_ptr(x2), _ptr(xk4), C.c_int64(x2.numel()), C.c_int32(DType.F16), C.c_int32(DType.F16_K4), _stream(x.device))
                L.qmatmul_f16act(_ptr(xk4), _ptr(self.w.data), _ptr(y), C.c_int32(m), C.c_int32(n), C.c_int32(k),
                                 C.c_int32(self.w.ggml_type), C.c_int32(0), _stream(x.device))
            else:
                need = L.qmatmul_workspace_bytes(C.c_int32(m), C.c_int32(n), C.c_int32(k))
                if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                    self._ws = torch.empty(int(need), dtype=torch.uint8, device=x.device)
                L.qmatmul_f32(_ptr(x2), _ptr(self.w.data), _ptr(y), C.c_int32(m), C.c_int32(n), C.c_int32(k),
                              C.c_int32(self.w.ggml_type), C.c_int32(0), _ptr(self._ws), C.c_size_t(self._ws.numel()),
                              _stream(x.device))
        check("QMatMul.forward")
        return y.reshape(*lead, n)

    def forward_slabs(self, x: torch.Tensor) -> torch.Tensor:
        """Atomic-free form the fused decode layer uses (``qmatmul_f16act_slabs``): returns ``[S, m, n]`` partial sums whose
        sum over S is the product; S = 1 when no tile is split.  Bitwise deterministic."""
        _cuda(x, "x"); require_device()
        n, k = self.w.shape
        if x.dim() != 2 or x.shape[-1] != k:
            raise BackendError(f"QMatMul: shape mismatch, x {tuple(x.shape)} vs weight {self.w.shape}")
        L = lib()
        m = x.shape[0]
        L.qmatmul_slab_count.restype = C.c_int32
        L.qmatmul_f16act_slabs.restype = C.c_int32
        s_max = int(L.qmatmul_slab_count(C.c_int32(m), C.c_int32(n), C.c_int32(k), C.c_int32(self.w.ggml_type)))
        y = torch.full((max(s_max, 1), m, n), float("nan"), dtype=torch.float32, device=x.device)    # every element must be overwritten
        with torch.cuda.device(x.device):
            x2 = x.contiguous().half()
            xk4 = torch.empty_like(x2)
            L.cast(_ptr(x2), _ptr(xk4), C.c_int64(x2.numel()), C.c_int32(DType.F16), C.c_int32(DType.F16_K4), _stream(x.device))
            got = int(L.qmatmul_f16act_slabs(_ptr(xk4), _ptr(self.w.data), _ptr(y), C.c_int32(y.shape[0]), C.c_int32(m), C.c_int32(n),
                                             C.c_int32(k), C.c_int32(self.w.ggml_type), _stream(x.device)))
        check("QMatMul.forward_slabs")
        return y[:got]


# ---------------------------------------------------------------------------------------------
# weight-only low-precision float linears (LnFp8 / LnNvfp4 / LnMxfp4, linear.rs:1190-1221, :1913-1943, :1717-1757)
# ---------------------------------------------------------------------------------------------
def _lin_io(x: torch.Tensor, n: int, k: int, bias, who: str):
    _cuda(x, "x"); require_device()
    if x.dtype not in (torch.float16, torch.bfloat16):
        raise BackendError(f"{who}: activations must be f16 or bf16 (the reference casts f32 inputs to bf16, linear.rs:1719-1724)")
    if x.shape[-1] != k:
        raise BackendError(f"{who}: shape mismatch, x {tuple(x.shape)} vs weight [{n}, {k}]")
    if bias is not None and (bias.dtype != x.dtype or bias.numel() != n):
        raise BackendError(f"{who}: bias must be [{n}] of the activation dtype")
    x2 = x.reshape(-1, k).contiguous()
    out = torch.empty((x2.shape[0], n), dtype=x.dtype, device=x.device)
    return x2, out


class Linear:
    """Dense 16-bit linear: ``Linear::new(weight [N, K], bias)`` / ``.forward(x)`` (/root/reference/src/openai/models/linear.rs:124-172;
    the reference reaches cuBLAS through candle's matmul).  f16 or bf16 weights and activations, fp32 accumulation on tcgen05
    (``linear_16bit``, csrc/dense_gemm.cu)."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None):
        _cuda(weight, "weight")
        if weight.dim() != 2 or weight.dtype not in (torch.float16, torch.bfloat16):
            raise BackendError("Linear: weight must be a 2-D f16 / bf16 tensor")
        self.weight = weight.contiguous()
        self.bias = None if bias is None else _cuda(bias, "bias").to(weight.dtype).contiguous()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, k = self.weight.shape
        if x.dtype != self.weight.dtype:
            raise BackendError(f"Linear: activation dtype {x.dtype} != weight dtype {self.weight.dtype}")
        x2, out = _lin_io(x, n, k, self.bias, "Linear")
        if k % 8 or n % 8:
            raise BackendError("Linear: in / out features must be multiples of 8 (16-byte rows for TMA)")
        with torch.cuda.device(x.device):
            lib().linear_16bit(_ptr(x2), _ptr(self.weight), _ptr(self.bias), _ptr(out), C.c_int32(x2.shape[0]), C.c_int32(n), C.c_int32(k),
                               C.c_int32(_dt(x2)), _stream(x.device))
        check("Linear.forward")
        return out.reshape(*x.shape[:-1], n)


class LnFp8:
    """Block-scaled FP8 linear: ``weight`` e4m3 (torch.float8_e4m3fn or u8) [N, K], ``weight_scale`` f32
    [ceil(N/by), ceil(K/bx)] (linear.rs:944-973); ``forward`` = ``fp8_matmul`` (+ bias)."""

    def __init__(self, weight: torch.Tensor, weight_scale: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 weight_block_size=(128, 128)):
        if len(weight_block_size) != 2:
            raise BackendError("LnFp8: weight_block_size must have 2 elements")          # linear.rs:949-951
        self.weight = _cuda(weight, "weight").contiguous().view(torch.uint8)
        self.by, self.bx = int(weight_block_size[0]), int(weight_block_size[1])
        n, k = self.weight.shape
        want = ((n + self.by - 1) // self.by, (k + self.bx - 1) // self.bx)
        if tuple(weight_scale.shape) != want or weight_scale.dtype != torch.float32:
            raise BackendError(f"LnFp8: weight_scale must be f32 {want}, got {tuple(weight_scale.shape)} {weight_scale.dtype}")
        self.weight_scale = _cuda(weight_scale, "weight_scale").contiguous()
        self.bias = bias

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, k = self.weight.shape
        x2, out = _lin_io(x, n, k, self.bias, "LnFp8")
        with torch.cuda.device(x.device):
            lib().fp8_matmul(_ptr(x2), _ptr(self.weight), _ptr(self.weight_scale), _ptr(self.bias), _ptr(out), C.c_int32(x2.shape[0]),
                             C.c_int32(n), C.c_int32(k), C.c_int32(self.by), C.c_int32(self.bx), C.c_int32(_dt(x2)), _stream(x.device))
        check("LnFp8.forward")
        return out.reshape(*x.shape[:-1], n)


class LnNvfp4:
    """NVFP4 linear: ``blocks`` u8 [N, K/2], ``scales`` e4m3 [N, K/16], ``global_scale`` as stored by the reference
    (the reciprocal of weight_global_scale, or weight_scale_2; linear.rs:1829-1853).  ``input_scale`` is accepted and ignored."""

    def __init__(self, blocks: torch.Tensor, scales: torch.Tensor, global_scale: float = 1.0, input_scale: float = 1.0,
                 bias: Optional[torch.Tensor] = None):
        self.blocks = _cuda(blocks, "blocks").contiguous().view(torch.uint8)
        self.scales = _cuda(scales, "scales").contiguous().view(torch.uint8)
        n, k2 = self.blocks.shape
        if tuple(self.scales.shape) != (n, k2 // 8):
            raise BackendError(f"LnNvfp4: scales must be [{n}, {k2 // 8}], got {tuple(self.scales.shape)}")
        self.global_scale, self.input_scale, self.bias = float(global_scale), float(input_scale), bias

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, k = self.blocks.shape[0], self.blocks.shape[1] * 2
        x2, out = _lin_io(x, n, k, self.bias, "LnNvfp4")
        with torch.cuda.device(x.device):
            lib().nvfp4_matmul(_ptr(x2), _ptr(self.blocks), _ptr(self.scales), C.c_float(self.global_scale), C.c_float(self.input_scale),
                               _ptr(self.bias), _ptr(out), C.c_int32(x2.shape[0]), C.c_int32(n), C.c_int32(k), C.c_int32(_dt(x2)),
                               _stream(x.device))
        check("LnNvfp4.forward")
        return out.reshape(*x.shape[:-1], n)


class LnMxfp4:
    """MXFP4 linear: ``blocks`` u8 [N, K/2], ``scales`` e8m0 [N, K/32] (linear.rs:1686-1700)."""

    def __init__(self, blocks: torch.Tensor, scales: torch.Tensor, bias: Optional[torch.Tensor] = None):
        self.blocks = _cuda(blocks, "blocks").contiguous().view(torch.uint8)
        self.scales = _cuda(scales, "scales").contiguous().view(torch.uint8)
        n, k2 = self.blocks.shape
        if tuple(self.scales.shape) != (n, k2 // 16):
            raise BackendError(f"LnMxfp4: scales must be [{n}, {k2 // 16}], got {tuple(self.scales.shape)}")
        self.bias = bias

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, k = self.blocks.shape[0], self.blocks.shape[1] * 2
        x2, out = _lin_io(x, n, k, self.bias, "LnMxfp4")
        with torch.cuda.device(x.device):
            lib().mxfp4_matmul(_ptr(x2), _ptr(self.blocks), _ptr(self.scales), _ptr(self.bias), _ptr(out), C.c_int32(x2.shape[0]),
                               C.c_int32(n), C.c_int32(k), C.c_int32(_dt(x2)), _stream(x.device))
        check("LnMxfp4.forward")
        return out.reshape(*x.shape[:-1], n)


# ---------------------------------------------------------------------------------------------
# small ops
# ---------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float, out_dtype=torch.float32) -> torch.Tensor:
    _cuda(x, "x"); require_device()
    x2 = x.reshape(-1, x.shape[-1]).contiguous().float()
    out = torch.empty(x2.shape, dtype=out_dtype, device=x.device)
    with torch.cuda.device(x.device):
        lib().rms_norm(_ptr(x2), _ptr(weight.float().contiguous()), _ptr(out), C.c_int32(x2.shape[0]), C.c_int32(x2.shape[1]),
                       C.c_float(eps), C.c_int32(_TORCH2B200[out_dtype]), _stream(x.device))
    check("rms_norm")
    return out.reshape(x.shape)


def fused_rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, positions: torch.Tensor,
               is_rope_i: bool) -> None:
    """``FusedRope::apply_inplace(q, k, cos, sin, positions, is_rope_i)`` on f32 [T, h, hd] (in place)."""
    _cuda(q, "q"); require_device()
    if q.dtype != torch.float32 or k.dtype != torch.float32 or not q.is_contiguous() or not k.is_contiguous():
        raise BackendError("fused_rope: q, k must be contiguous f32")
    with torch.cuda.device(q.device):
        lib().fused_rope_f32(_ptr(q), _ptr(k), _ptr(cos), _ptr(sin), _ptr(positions), C.c_int32(q.shape[0]),
                             C.c_int32(q.shape[1]), C.c_int32(k.shape[1]), C.c_int32(q.shape[2]),
                             C.c_int32(1 if is_rope_i else 0), _stream(q.device))
    check("fused_rope")


def silu_mul(gate: torch.Tensor, up: torch.Tensor, out_dtype=torch.float32) -> torch.Tensor:
    _cuda(gate, "gate"); require_device()
    g, u = gate.contiguous().float(), up.contiguous().float()
    out = torch.empty(g.shape, dtype=out_dtype, device=g.device)
    with torch.cuda.device(g.device):
        lib().silu_mul(_ptr(g), _ptr(u), _ptr(out), C.c_int64(g.numel()), C.c_int32(_TORCH2B200[out_dtype]), _stream(g.device))
    check("silu_mul")
    return out


def argmax(logits: torch.Tensor) -> torch.Tensor:
    _cuda(logits, "logits"); require_device()
    x = logits.reshape(-1, logits.shape[-1]).contiguous().float()
    out = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        lib().argmax_f32(_ptr(x), _ptr(out), C.c_int32(x.shape[0]), C.c_int32(x.shape[1]), _stream(x.device))
    check("argmax")
    return out
