"""ctypes binding of oracle/_ref/libcpu_ref.so (C restatement of the reference CPU path).

Oracle / CPU-baseline infrastructure only -- see ``oracle/__init__.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libcpu_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpu_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.ref_vec_dot_q4_K_q8_K.restype = C.c_float
        _lib.ref_vec_dot_q6_K_q8_K.restype = C.c_float
        _lib.ref_layer_scratch_floats.restype = C.c_size_t
    return _lib


class RefCfg(C.Structure):
    _fields_ = [("hidden", C.c_int), ("heads", C.c_int), ("kv_heads", C.c_int), ("head_dim", C.c_int),
                ("ffn", C.c_int), ("block_size", C.c_int), ("max_blocks", C.c_int), ("rms_eps", C.c_float)]


class RefLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("attn_norm", "ffn_norm", "wq", "wk", "wv", "wo", "w1", "w2", "w3")]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().ref_num_threads())


def set_num_threads(n: int) -> None:
    lib().ref_set_num_threads(C.c_int(int(n)))


def qmatmul_q8k(x: np.ndarray, w: np.ndarray, ggml_type: int, n: int, k: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32).reshape(-1, k)
    w = np.ascontiguousarray(w, np.uint8)
    y = np.empty((x.shape[0], n), np.float32)
    rc = lib().ref_qmatmul_q8k(_p(x), x.shape[0], _p(w), ggml_type, n, k, _p(y))
    if rc:
        raise ValueError(f"ref_qmatmul_q8k rc={rc}")
    return y


def paged_attention_decode_bf16(q, kc_bits, vc_bits, block_tables, context_lens, scale):
    """q f32 [B,H,hd]; caches uint16 bf16 bits [nb,bs,kvh,hd]."""
    q = np.ascontiguousarray(q, np.float32)
    B, H, hd = q.shape
    nb, bs, kvh, _ = kc_bits.shape
    bt = np.ascontiguousarray(block_tables, np.uint32)
    cl = np.ascontiguousarray(context_lens, np.uint32)
    out = np.empty_like(q)
    lib().ref_paged_attention_decode_bf16(_p(q), _p(kc_bits), _p(vc_bits), _p(bt), _p(cl), B, H, kvh, hd, bs,
                                          bt.shape[1], C.c_float(scale), _p(out))
    return out
