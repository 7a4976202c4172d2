"""CPU tests: the C-ABI library loads and exports every symbol include/b200_backend.h declares
(no compute calls: there is no GPU here and the library has no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import candle_vllm_b200 as pkg
from candle_vllm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "b200_backend.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_exports_every_declared_symbol():
    L = pkg.lib()
    declared = _declared_functions()
    assert len(declared) >= 35
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, f"library lacks declared symbols: {missing}"
    # and the python-side list covers the header
    assert set(declared) <= set(_lib.SYMBOLS)


def test_reference_ffi_symbol_names_present():
    # the exact names the reference imports from attention_rs::kernels::ffi (src/backend/cache.rs:2)
    L = pkg.lib()
    for n in ("copy_blocks_bf16", "copy_blocks_f16", "copy_blocks_f32"):
        assert hasattr(L, n)


def test_abi_version_and_error_channel():
    L = pkg.lib()
    assert L.b200_abi_version() == 1
    assert L.b200_last_error() == 0
    # argument validation happens before any CUDA call: a null pointer is reported, not dereferenced
    L.swap_blocks(None, None, None, ctypes.c_int32(3), ctypes.c_int64(16), ctypes.c_int64(0))
    assert L.b200_last_error() == 1
    assert b"null pointer" in L.b200_last_error_message()
    assert L.b200_last_error() == 0          # reading clears
    # zero work is a no-op like the reference (cache.rs:41-44)
    L.copy_blocks_bf16(None, None, None, 0, 0, 0, ctypes.c_int64(0))
    assert L.b200_last_error() == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert not pkg.device_ok()
    with pytest.raises(pkg.BackendError):
        _lib.require_device()
    with pytest.raises(pkg.BackendError):
        pkg.copy_blocks([torch.zeros(2, 4, 1, 8, dtype=torch.bfloat16)], [torch.zeros(2, 4, 1, 8, dtype=torch.bfloat16)], {0: [1]})
    # engine creation refuses too
    cfg = pkg.llama._CCfg(256, 1, 2, 1, 128, 256, 512, 16, 4, 8, 128, 1e-5, 1e4, 2, 0, 1, 1)
    h = pkg.lib().b200_llama_create(ctypes.byref(cfg))
    assert not h and pkg.lib().b200_last_error() != 0


def test_product_does_not_import_oracle():
    pdir = os.path.join(ROOT, "candle-vllm_b200")
    for dp, _, fs in os.walk(pdir):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt or f.endswith(".md"), f
