"""Stand-alone check/timing of the quantised GEMM through the C ABI (debug + profiling helper).
  python tools/gemm_check.py [m] [n] [k] [type: 12=Q4_K 14=Q6_K] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import candle_vllm_b200 as pkg
from candle_vllm_b200 import synthetic

m = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
t = int(sys.argv[4]) if len(sys.argv) > 4 else 12
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = synthetic.random_qtensor(g, t, n, k, "cuda")
x = torch.randn((m, k), device="cuda", generator=g)
mm = pkg.QMatMul(w)
y = mm.forward(x)
torch.cuda.synchronize()
wd = w.dequantize()
ref = (x.half().float().double() @ wd.double().T).float()      # fp16-rounded activations, exact weights
ref32 = (x.double() @ wd.double().T).float()
rel = ((y - ref).norm() / ref.norm()).item()
rel32 = ((y - ref32).norm() / ref32.norm()).item()
print(f"m={m} n={n} k={k} type={t}: rel-fro vs fp16-act ref {rel:.3e}, vs f32 ref {rel32:.3e}, max|y|={y.abs().max().item():.3f}")
if reps:
    L = 12
    ws = [synthetic.random_qtensor(g, t, n, k, "cuda") for _ in range(L)]
    mms = [pkg.QMatMul(wi) for wi in ws]
    xh = x.half()
    for i in range(3): mms[i % L].forward(xh)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(reps): mms[i % L].forward(xh)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = w.data.numel()
    print(f"  {ms*1e3:.1f} us/call (incl. K4 cast + memset), {byts/ms/1e6:.1f} GB/s weight stream")
