"""Generates the committed golden vectors under tests/golden/ (run in the authoring container).

  ggml_dequant.npz : random GGML blocks (Q4_K, Q6_K, Q8_0) + their dequantisation by the `gguf`
                     Python package (gguf.quants.dequantize -- llama.cpp's own independent numpy
                     implementation of the GGUF block formats, v0.19) + Q8_0 quantisation of a
                     random vector by gguf.quants.quantize.  Pins oracle/ggml_quants.py.
  slot_mapping.json: known answers for the slot / block-table arithmetic restated from
                     /root/reference/src/openai/pipelines/inputs.rs:12-22, :410-430 (hand-computed
                     from the formulae; the reference holds no test vectors for them).
The reference itself (Rust, kernels in un-vendored deps) cannot be executed here, so no vectors
come from running it.
"""
import json
import os

import numpy as np
from gguf import GGMLQuantizationType as T
from gguf import quants

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260922)
    out = {}
    for name, t, bb in (("q4k", T.Q4_K, 144), ("q6k", T.Q6_K, 210), ("q8_0", T.Q8_0, 34)):
        nb = 24
        blocks = rng.integers(0, 256, (nb, bb), dtype=np.uint8)
        # keep the f16 scale fields finite and moderate
        if name == "q4k":
            blocks[:, 0:2] = (rng.uniform(0.5, 2, nb) * 2.0 ** -10).astype(np.float16).view(np.uint8).reshape(nb, 2)
            blocks[:, 2:4] = (rng.uniform(0.5, 2, nb) * 2.0 ** -8).astype(np.float16).view(np.uint8).reshape(nb, 2)
        elif name == "q6k":
            blocks[:, 208:210] = (rng.uniform(0.5, 2, nb) * 2.0 ** -12).astype(np.float16).view(np.uint8).reshape(nb, 2)
        else:
            blocks[:, 0:2] = (rng.uniform(0.5, 2, nb) * 2.0 ** -6).astype(np.float16).view(np.uint8).reshape(nb, 2)
        out[f"{name}_blocks"] = blocks
        out[f"{name}_deq"] = quants.dequantize(blocks, t).astype(np.float32)
    x = rng.standard_normal((4, 64)).astype(np.float32)
    out["q8_0_x"] = x
    out["q8_0_quant"] = quants.quantize(x, T.Q8_0)
    np.savez_compressed(os.path.join(HERE, "ggml_dequant.npz"), **out)

    cases = [
        # (seq_len incl. decoded token, block_size, table) -> position, slot, used blocks
        dict(seq_len=1, block_size=64, table=[7], position=0, slot=448, used=1),
        dict(seq_len=64, block_size=64, table=[7], position=63, slot=511, used=1),
        dict(seq_len=65, block_size=64, table=[7, 3], position=64, slot=192, used=2),
        dict(seq_len=4097, block_size=64, table=list(range(100, 165)), position=4096, slot=164 * 64, used=65),
        dict(seq_len=130, block_size=16, table=[5, 9, 2, 11, 4, 8, 1, 0, 3], position=129, slot=3 * 16 + 1, used=9),
        dict(seq_len=17, block_size=16, table=[5, 9, 2], position=16, slot=9 * 16, used=2),
    ]
    with open(os.path.join(HERE, "slot_mapping.json"), "w") as f:
        json.dump(cases, f, indent=1)


if __name__ == "__main__":
    main()


def marlin_perms():
    """tests/golden/marlin_perms.json: the scale permutation tables produced by EXECUTING the reference's own
    Python (`get_scale_perms`, /root/reference/examples/convert_awq_marlin.py:8-17; the function source is
    extracted with `ast` because the module imports safetensors)."""
    import ast
    src = open("/root/reference/examples/convert_awq_marlin.py").read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_scale_perms"][0]
    ns = {"List": list}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref", "exec"), ns)
    sp, sps = ns["get_scale_perms"]()
    with open(os.path.join(HERE, "marlin_perms.json"), "w") as f:
        json.dump({"scale_perm": sp, "scale_perm_single": sps, "source": "examples/convert_awq_marlin.py:8-17 (executed)"}, f)


if __name__ == "__main__" and os.path.exists("/root/reference/examples/convert_awq_marlin.py"):
    marlin_perms()


def awq_zero_points():
    """tests/golden/awq_zero_points.json: AWQ-packed zero points and what the reference's offline converter turns them into
    (`awq_to_marlin_zero_points`, /root/reference/examples/convert_awq_marlin.py:99-113, EXECUTED here with safetensors
    stubbed out) -- the layout `marlin_awq_4bit_*` takes as `qzeros`."""
    import importlib.util
    import sys
    import types
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle import gptq as OG
    for m in ("safetensors", "safetensors.torch"):
        mod = types.ModuleType(m); mod.load_file = mod.save_file = None; sys.modules.setdefault(m, mod)
    spec = importlib.util.spec_from_file_location("conv", "/root/reference/examples/convert_awq_marlin.py")
    conv = importlib.util.module_from_spec(spec); spec.loader.exec_module(conv)
    zp = np.random.default_rng(0).integers(0, 16, (3, 128), dtype=np.uint8)
    packed = OG.pack_awq(zp)
    ref = conv.awq_to_marlin_zero_points(torch.from_numpy(packed.view(np.int32)), 3, 128, 4).numpy().view(np.uint32)
    with open(os.path.join(HERE, "awq_zero_points.json"), "w") as f:
        json.dump({"zp": zp.tolist(), "awq_packed": packed.tolist(), "marlin_zp": ref.tolist()}, f)


if __name__ == "__main__" and os.path.exists("/root/reference/examples/convert_awq_marlin.py"):
    awq_zero_points()
