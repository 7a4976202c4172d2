"""GGUF-LLaMA decode engine binding: mirror of ``GGUFLLaMa`` (models/quantized_llama.rs) +
``GraphCapturer`` replay (backend/graph.rs) over ``b200_llama_*`` (include/b200_backend.h)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from ._lib import BackendError, check, lib, require_device
from .backend import DType, GgmlType, QTensor


@dataclass
class LlamaConfig:
    hidden: int = 4096
    num_layers: int = 32
    num_heads: int = 32
    num_kv_heads: int = 8
    head_dim: int = 128
    ffn: int = 14336
    vocab: int = 128256
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    max_pos: int = 8192
    block_size: int = 64
    max_num_seqs: int = 32
    max_blocks_per_seq: int = 128

    @classmethod
    def llama3_8b(cls, **kw):
        return cls(**kw)


def padded_vocab(vocab: int, world: int) -> int:
    """pad_vocab_size (/root/reference/src/openai/distributed.rs:1448-1454); identity at world 1."""
    if world == 1:
        return vocab
    padded = -(-vocab // 64) * 64
    per_rank = -(-padded // world) * world
    return -(-per_rank // 64) * 64


class _CCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("hidden", "num_layers", "num_heads", "num_kv_heads", "head_dim", "ffn", "vocab",
                                          "block_size", "max_num_seqs", "max_blocks_per_seq", "max_pos")] + \
               [("rms_eps", C.c_float), ("rope_theta", C.c_float)] + \
               [(n, C.c_int32) for n in ("kv_dtype", "tp_rank", "tp_world", "use_graph")]


class _CLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("attn_norm", "ffn_norm", "wq", "wk", "wv", "wo", "w1", "w2", "w3")] + \
               [(n, C.c_int32) for n in ("tq", "tk", "tv", "to", "t1", "t2", "t3")]


class _CLinear(C.Structure):
    _fields_ = [("kind", C.c_int32), ("type", C.c_int32), ("w", C.c_void_p), ("scales", C.c_void_p), ("zeros", C.c_void_p),
                ("group_size", C.c_int32), ("reserved", C.c_int32)]


class _CLayerEx(C.Structure):
    _fields_ = [("attn_norm", C.c_void_p), ("ffn_norm", C.c_void_p)] + [(n, _CLinear) for n in ("wq", "wk", "wv", "wo", "w1", "w2", "w3")]


@dataclass
class MarlinWeight:
    """GPTQ / AWQ int4 weight prepared like the reference prepares it for Marlin (linear.rs:300-413): ``qweight`` =
    ``marlin_weight_repack`` output, ``scales`` = ``marlin_permute_scales`` output (f16 / bf16), ``zeros`` = the converter's AWQ zero
    points or None (symmetric GPTQ)."""
    qweight: torch.Tensor
    scales: torch.Tensor
    group_size: int
    zeros: Optional[torch.Tensor] = None


_DT = {torch.float16: DType.F16, torch.bfloat16: DType.BF16}


def _clinear(w) -> _CLinear:
    """QTensor (GGML) | MarlinWeight (int4) | dense f16 / bf16 tensor [n, k] -> b200_linear."""
    if isinstance(w, QTensor):
        return _CLinear(0, w.ggml_type, w.data.data_ptr(), None, None, 0, 0)
    if isinstance(w, MarlinWeight):
        if w.scales.dtype not in _DT:
            raise BackendError("MarlinWeight: scales must be f16 / bf16")
        return _CLinear(1, _DT[w.scales.dtype], w.qweight.data_ptr(), w.scales.data_ptr(),
                        None if w.zeros is None else w.zeros.data_ptr(), int(w.group_size), 0)
    if isinstance(w, torch.Tensor) and w.dtype in _DT and w.dim() == 2 and w.is_contiguous():
        return _CLinear(2, _DT[w.dtype], w.data_ptr(), None, None, 0, 0)
    raise BackendError(f"unsupported linear weight {type(w)}")


class GGUFLLaMa:
    """weights: dict(tok_embeddings f32 [V,H], norm f32 [H], output, layers=[dict(attn_norm, ffn_norm f32; wq, wk, wv, wo, w1, w2, w3)])
    already sharded for (tp_rank, tp_world).  Every linear is a QTensor (GGML blocks: the GGUF models, quantized_llama.rs), a
    MarlinWeight (GPTQ / AWQ int4) or a dense f16 / bf16 tensor [n, k] (the safetensors models, llama.rs); ``rope_neox`` selects the
    rotation of the safetensors models (llama.rs:222) instead of GGUF's interleaved one."""

    def __init__(self, cfg: LlamaConfig, weights: dict, kv_cache: List, kv_dtype: int = DType.BF16,
                 tp_rank: int = 0, tp_world: int = 1, use_graph: bool = True, stream: Optional[torch.cuda.Stream] = None,
                 nccl_comm: Optional[int] = None, rope_neox: bool = False):
        require_device()
        self.cfg, self.weights, self.kv_cache = cfg, weights, kv_cache
        # Default = torch's current stream: every other op of this package (CacheEngine swap / copy, PagedAttention prefill,
        # the weight and KV uploads) is issued there, so a private stream would race them (decode reading blocks a swap-in
        # is still writing).  A caller that passes its own stream owns the ordering against those ops.
        self.stream = stream or torch.cuda.current_stream()
        self.tp_world = tp_world
        self.vocab_local = padded_vocab(cfg.vocab, tp_world) // tp_world
        c = _CCfg(cfg.hidden, cfg.num_layers, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.ffn, cfg.vocab,
                  cfg.block_size, cfg.max_num_seqs, cfg.max_blocks_per_seq, cfg.max_pos, cfg.rms_eps, cfg.rope_theta,
                  kv_dtype, tp_rank, tp_world, 1 if use_graph else 0)
        L = lib()
        self._h = C.c_void_p(L.b200_llama_create(C.byref(c)))
        check("b200_llama_create")
        if not self._h:
            raise BackendError("b200_llama_create returned null")
        names = ("wq", "wk", "wv", "wo", "w1", "w2", "w3")
        for i, lw in enumerate(weights["layers"]):
            if all(isinstance(lw[k], QTensor) for k in names):       # the reference's GGUF form
                cl = _CLayer(lw["attn_norm"].data_ptr(), lw["ffn_norm"].data_ptr(), *[lw[k].data.data_ptr() for k in names],
                             *[lw[k].ggml_type for k in names])
                L.b200_llama_set_layer(self._h, C.c_int32(i), C.byref(cl))
            else:
                ce = _CLayerEx(lw["attn_norm"].data_ptr(), lw["ffn_norm"].data_ptr(), *[_clinear(lw[k]) for k in names])
                L.b200_llama_set_layer_ex(self._h, C.c_int32(i), C.byref(ce))
            check("b200_llama_set_layer")
        out = _clinear(weights["output"])
        L.b200_llama_set_globals_ex(self._h, C.c_void_p(weights["tok_embeddings"].data_ptr()), C.c_void_p(weights["norm"].data_ptr()),
                                    C.byref(out), C.c_int32(1 if rope_neox else 0))
        check("b200_llama_set_globals")
        kp = (C.c_void_p * cfg.num_layers)(*[k.data_ptr() for k, _ in kv_cache])
        vp = (C.c_void_p * cfg.num_layers)(*[v.data_ptr() for _, v in kv_cache])
        L.b200_llama_set_kv_cache(self._h, kp, vp, C.c_int64(kv_cache[0][0].shape[0]))
        check("b200_llama_set_kv_cache")
        if nccl_comm is not None:
            L.b200_llama_set_comm(self._h, C.c_void_p(nccl_comm))
            check("b200_llama_set_comm")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().b200_llama_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def decode(self, prep: dict, want_logits: bool = False):
        """One decode step from HOST metadata (``inputs.prepare_decode`` output).  Returns
        (next_tokens i32[B] numpy, logits f32 [B, vocab] numpy or None); with tp_world > 1 the logits are the gathered full
        vocabulary (a collective: every rank passes want_logits in the same step)."""
        B = len(prep["tokens"])
        tokens = np.ascontiguousarray(prep["tokens"], np.uint32)
        pos = np.ascontiguousarray(prep["positions"], np.int64)
        slots = np.ascontiguousarray(prep["slot_mapping"], np.int64)
        ctx = np.ascontiguousarray(prep["context_lens"]).astype(np.uint32)
        bt = np.ascontiguousarray(prep["block_tables"]).astype(np.uint32)
        nxt = np.empty(B, np.int32)
        logits = np.empty((B, self.cfg.vocab), np.float32) if want_logits else None
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib().b200_llama_decode(self._h, p(tokens), p(pos), p(slots), p(ctx), p(bt), C.c_int32(bt.shape[1]), C.c_int32(B),
                                p(nxt), p(logits) if want_logits else C.c_void_p(0), C.c_int64(self.stream.cuda_stream))
        check("b200_llama_decode")
        return nxt, logits

    def decode_resident(self, num_seqs: int, advance: bool = True) -> None:
        """Replay on the device-resident metadata (tokens <- previous argmax, positions + 1)."""
        lib().b200_llama_decode_resident(self._h, C.c_int32(num_seqs), C.c_int32(1 if advance else 0),
                                         C.c_int64(self.stream.cuda_stream))
        check("b200_llama_decode_resident")

    def uses_layer_kernel(self, num_seqs: int) -> bool:
        """True when decode steps of this batch size run on the persistent layer kernel (deterministic split-K reduction)."""
        return bool(lib().b200_llama_uses_layer_kernel(self._h, C.c_int32(num_seqs)))

    def linear_chain(self, num_seqs: int) -> None:
        """Measurement aid (``b200_llama_linear_chain``): every projection of the model without the attention / KV stream."""
        lib().b200_llama_linear_chain(self._h, C.c_int32(num_seqs), C.c_int64(self.stream.cuda_stream))
        check("b200_llama_linear_chain")

    def read_next_tokens(self, n: int) -> np.ndarray:
        out = np.empty(n, np.int32)
        lib().b200_llama_read_next_tokens(self._h, out.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_int64(self.stream.cuda_stream))
        check("b200_llama_read_next_tokens")
        return out

    def read_logits(self, n: int) -> np.ndarray:
        out = np.empty((n, self.cfg.vocab), np.float32)
        lib().b200_llama_read_logits(self._h, out.ctypes.data_as(C.c_void_p), C.c_int32(n), C.c_int64(self.stream.cuda_stream))
        check("b200_llama_read_logits")
        return out

    def kernel_launches(self) -> int:
        return int(lib().b200_llama_kernel_launches(self._h))
