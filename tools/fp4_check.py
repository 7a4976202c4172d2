"""Timing of the weight-only FP4 / FP8 linears at decode sizes through the C ABI (profiling helper): bytes streamed per call against
the measured HBM peak.  python tools/fp4_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import candle_vllm_b200 as pkg

dev = torch.device("cuda:0")
peak = 6583.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rng = np.random.default_rng(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


print(f"weight-only linears, m = 32, bf16 activations; HBM peak {peak:.0f} GB/s (measured copy)")
for (n, k) in [(4096, 4096), (28672, 4096), (4096, 14336)]:
    x = torch.randn((32, k), device=dev).to(torch.bfloat16)
    blocks = torch.from_numpy(rng.integers(0, 256, (n, k // 2), dtype=np.uint8)).to(dev)
    sc = torch.from_numpy(rng.integers(0x28, 0x58, (n, k // 16), dtype=np.uint8)).to(dev)
    se = torch.from_numpy(rng.integers(117, 125, (n, k // 32), dtype=np.uint8)).to(dev)
    w8 = torch.from_numpy(rng.integers(0, 0x78, (n, k), dtype=np.uint8)).to(dev)
    s8 = torch.rand((n // 128, k // 128), device=dev) * 0.01 + 0.001
    for name, lin, byts in [("nvfp4", pkg.LnNvfp4(blocks, sc, 1.0 / 448, 1.0, None), n * k // 2 + n * k // 16),
                            ("mxfp4", pkg.LnMxfp4(blocks, se), n * k // 2 + n * k // 32),
                            ("fp8  ", pkg.LnFp8(w8, s8), n * k)]:
        ms = timed(lambda: lin.forward(x))
        print(f"  {name} n={n:5d} k={k:5d}: {ms*1e3:7.1f} us, {byts/ms/1e6:7.1f} GB/s ({byts/ms/1e6/peak:.2f} of peak) incl. activation cast + finishing pass")
