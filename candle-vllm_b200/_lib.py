"""ctypes loader for libb200backend.so (include/b200_backend.h).  Fails loudly if missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libb200backend.so")
_lib = None


class BackendError(RuntimeError):
    """Raised where the reference would ``candle_core::bail!`` or surface a CUDA error."""


def lib_path() -> str:
    return _SO


# every symbol include/b200_backend.h declares (checked by tests/test_abi.py against the header)
SYMBOLS = [
    "b200_abi_version", "b200_last_error", "b200_last_error_message", "b200_device_sm_count", "b200_device_cc",
    "copy_blocks_bf16", "copy_blocks_f16", "copy_blocks_f32", "copy_blocks_u8", "swap_blocks", "flashinfer_csr_to_paged", "reshape_and_cache",
    "paged_attention_decode_workspace_bytes", "paged_attention_decode", "paged_attention_prefill",
    "qmatmul_workspace_bytes", "qmatmul_f32", "qmatmul_f16act", "qmatmul_slab_count", "qmatmul_f16act_slabs", "b200_llama_peer_inbox_bytes", "b200_llama_set_peer_inboxes", "b200_llama_peer_timeouts", "b200_ipc_alloc", "b200_ipc_open", "b200_ipc_close", "b200_ipc_free",
    "dequantize_f32", "linear_16bit", "fp8_matmul", "nvfp4_matmul", "mxfp4_matmul",
    "concat_and_cache_mla", "mla_paged_decode_workspace_bytes", "mla_paged_attention",
    "topk_softmax", "sort_expert_assignments", "moe_gemm_workspace_bytes", "moe_gemm_gguf", "moe_gemm_fp8_workspace_bytes", "moe_gemm_fp8",
    "rms_norm", "fused_rope_f32", "silu_mul", "add_f32", "cast", "embedding_f32", "argmax_f32", "rope_and_cache",
    "b200_llama_create", "b200_llama_destroy", "b200_llama_set_layer", "b200_llama_set_globals", "b200_llama_set_layer_ex", "b200_llama_set_globals_ex",
    "b200_llama_set_kv_cache", "b200_llama_set_comm", "b200_llama_decode", "b200_llama_decode_resident", "b200_llama_linear_chain", "b200_llama_uses_layer_kernel", "b200_llama_mega_trace",
    "b200_llama_logits", "b200_llama_next_tokens", "b200_llama_kernel_launches",
    "b200_llama_read_next_tokens", "b200_llama_read_logits",
    "b200_total_kernel_launches", "b200_allreduce_f32",
    "gptq_repack", "marlin_checkpoint_repack", "awq_repack", "marlin_4bit_f16", "marlin_4bit_bf16", "marlin_awq_4bit_f16", "marlin_awq_4bit_bf16",
    "gemm_half_q_half_alt", "b200_set_scratch",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise BackendError(
                f"{_SO} not found: build it with `python candle-vllm_b200/build.py` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(_SO)
        L.b200_last_error_message.restype = C.c_char_p
        L.paged_attention_decode_workspace_bytes.restype = C.c_size_t
        L.qmatmul_workspace_bytes.restype = C.c_size_t
        L.moe_gemm_workspace_bytes.restype = C.c_size_t
        L.moe_gemm_fp8_workspace_bytes.restype = C.c_size_t
        L.mla_paged_decode_workspace_bytes.restype = C.c_size_t
        L.b200_llama_create.restype = C.c_void_p
        L.b200_llama_peer_inbox_bytes.restype = C.c_size_t
        L.b200_ipc_alloc.restype = C.c_void_p
        L.b200_ipc_open.restype = C.c_void_p
        L.b200_llama_logits.restype = C.c_void_p
        L.b200_llama_next_tokens.restype = C.c_void_p
        L.b200_llama_kernel_launches.restype = C.c_int64
        L.b200_total_kernel_launches.restype = C.c_longlong
        for name in SYMBOLS:
            getattr(L, name)        # AttributeError if the library lacks a declared symbol
        _lib = L
    return _lib


def check(what: str = "") -> None:
    """Raise BackendError if the last call on this thread recorded an error."""
    L = lib()
    code = L.b200_last_error()
    if code:
        raise BackendError(f"{what}: {L.b200_last_error_message().decode()} (code {code})")


def device_ok() -> bool:
    """True when the current CUDA device is an sm_100 part."""
    try:
        return lib().b200_device_cc() >= 100
    except Exception:
        return False


def require_device() -> None:
    cc = lib().b200_device_cc()
    if cc < 100:
        raise BackendError(f"no sm_100 (B200) device available (compute capability {cc / 10:.1f}); "
                           "this backend has no CPU or other-GPU fallback")
