"""One int4 (GPTQ -> repacked) GEMM at a decode shape through the C ABI, for profiling: python tools/marlin_gemm_check.py [n] [k] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import candle_vllm_b200 as pkg
from oracle import gptq as OG
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28672
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
m, g = 32, 128
rng = np.random.default_rng(0)
q = rng.integers(0, 16, (k, n), dtype=np.uint8)
st = torch.from_numpy(rng.uniform(0.005, 0.02, (k // g, n)).astype(np.float32)).cuda().half()
qw = torch.from_numpy(OG.pack_gptq(q).view(np.int32)).cuda()
w_m = pkg.marlin_weight_repack(qw, 4, False)
s_m = pkg.marlin_permute_scales(st, k, n, g)
ws = torch.zeros(n, dtype=torch.int32, device="cuda")
xt = torch.randn((m, k), device="cuda").half()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    y = pkg.gptq_matmul(xt, w_m, s_m, None, None, ws, 4, g)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(reps):
    flush.zero_()
    e0.record(); y = pkg.gptq_matmul(xt, w_m, s_m, None, None, ws, 4, g); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
byts = n * k // 2 + n * (k // g) * 2
print(f"int4 n={n} k={k} m={m}: {tot/reps*1e3:.1f} us per call (cast + GEMM + finishing pass, eager), {byts/(tot/reps)/1e6:.0f} GB/s")
