#!/usr/bin/env python
"""Timeline of one launch of the persistent layer kernel (csrc/layer_mega.cu) inside a warm decode step.

  B200_MEGA_TRACE=<launch index, 1 = layer 0's chain> python tools/mega_trace.py [batch] [ctx]

Stamps per CTA and phase (clock64, cycles): 0 phase entered, 1 units dequantised ahead / elementwise op reached, 2 previous phase complete
on every CTA, 3 op done, 4 last unit dequantised, 5 last epilogue stored + arrival, 6 first activations landed (MMA warp).
Prints, per phase, min / median / max over CTAs of the intervals, in microseconds at the SM clock."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200_MEGA_TRACE", "3")
import candle_vllm_b200 as pkg  # noqa: E402
from candle_vllm_b200 import synthetic  # noqa: E402
from candle_vllm_b200._lib import lib  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    layers = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    cfg = pkg.LlamaConfig(num_layers=layers, max_num_seqs=B, max_blocks_per_seq=80, max_pos=5200)
    w = synthetic.make_weights(cfg, "cuda", seed=0)
    nb = B * 80 + 8
    eng = pkg.CacheEngine(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim, pkg.CacheConfig(64, nb))
    synthetic.fill_kv_cache(eng.gpu_cache, seed=1)
    torch.cuda.synchronize()
    model = pkg.GGUFLLaMa(cfg, w, eng.gpu_cache)
    tables = synthetic.random_block_tables(B, 80, nb, seed=2)
    model.decode(pkg.prepare_decode([ctx + 1] * B, [1] * B, tables, 64))
    for _ in range(6):
        model.decode_resident(B, advance=True)
    torch.cuda.synchronize()
    G = lib().b200_device_sm_count()
    buf = np.zeros((G, 4, 8), np.int64)
    lib().b200_llama_mega_trace.restype = C.c_int32
    n = lib().b200_llama_mega_trace(model._h, buf.ctypes.data_as(C.c_void_p), C.c_int32(G))
    mhz = 1965.0
    t = buf[:n].astype(np.float64) / mhz          # us
    names = ["wo", "gate|up", "w2", "QKV(next)"]
    base = t[:, 0, 0]
    def st(a):
        a = a[np.isfinite(a)]
        return f"{a.min():7.2f} {np.median(a):7.2f} {a.max():7.2f}" if len(a) else "      -       -       -"
    print(f"launch {os.environ['B200_MEGA_TRACE']}, {n} CTAs; all times in us; columns = min median max over CTAs")
    for ph in range(4):
        e = t[:, ph]
        z = lambda k: np.where(buf[:n, ph, k] != 0, e[:, k] - base, np.nan)
        d = lambda a, b: np.where((buf[:n, ph, a] != 0) & (buf[:n, ph, b] != 0), e[:, a] - e[:, b], np.nan)
        print(f"-- phase {ph} {names[ph]}")
        print(f"   entered (since launch)      {st(z(0))}")
        if ph:
            print(f"   ahead-dequant -> op reached {st(d(1, 0))}")
            print(f"   wait prev phase everywhere  {st(d(2, 1))}")
            print(f"   elementwise op + publish    {st(d(3, 2))}")
            print(f"   op done -> X landed (MMA)   {st(d(6, 3))}")
        else:
            print(f"   entered -> X landed (MMA)   {st(d(6, 0))}")
        print(f"   X landed -> last unit deq   {st(d(4, 6))}")
        print(f"   last unit -> epilogue+arrive{st(d(5, 4))}")
        print(f"   phase total                 {st(d(5, 0))}")
    print(f"kernel (phase 0 entry -> phase 3 end): {st(t[:, 3, 5] - base)}")


if __name__ == "__main__":
    main()
