"""candle-vllm_b200 -- B200-native (sm_100a) batched-decode backend for candle-vllm.

Host-side mirror of the reference's operator interface for the hot path:
``copy_blocks`` / ``swap_blocks`` (src/backend/cache.rs, attention_rs::cache), ``PagedAttention`` +
``InputMetadata`` (attention-rs API used at layers/attention.rs:566-575,707-718), ``QMatMul`` /
``QTensor`` (candle; models/linear.rs:765-806), ``CacheEngine`` (src/scheduler/cache_engine.rs) and
the GGUF-LLaMA decode engine (models/quantized_llama.rs), all calling the C-ABI library
``libb200backend.so`` (include/b200_backend.h).  torch is used for device memory and streams only.
There is no CPU fallback: without the CUDA library / an sm_100 GPU every op raises.
"""
from ._lib import BackendError, lib, lib_path, device_ok  # noqa: F401
from .backend import (  # noqa: F401
    DType, GgmlType, KvLayout, copy_blocks, swap_blocks, reshape_and_cache, InputMetadata, FlashInferMetadata,
    PagedAttention, QTensor, QMatMul, Linear, LnFp8, LnNvfp4, LnMxfp4, rms_norm, fused_rope, silu_mul, argmax, dequantize,
)
from .cache_engine import CacheConfig, CacheEngine  # noqa: F401
from .inputs import prepare_decode, prepare_prompt, used_blocks_for_len, flashinfer_csr, PAD_SLOT_ID  # noqa: F401
from .llama import LlamaConfig, GGUFLLaMa, MarlinWeight  # noqa: F401
from .block_manager import BlockManager, PrefixCache, PrefixCacheConfig, Seq, SeqGroup, AllocStatus  # noqa: F401
from .gptq import gptq_matmul, marlin_weight_repack, marlin_permute_scales  # noqa: F401
from .linear import QLinear  # noqa: F401
from .moe import FusedMoe, topk_softmax, sort_expert_assignments, moe_gemm_gguf, moe_gemm_fp8  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
