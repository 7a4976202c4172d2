"""Timing of the M >= 128 GEMMs through the C ABI (profiling helper): the dense tcgen05 GEMM (`linear_16bit`) against the measured
bf16 peak in MEASURED_PEAKS.json (and cuBLAS via torch.matmul on the same shapes, for orientation only), and the quantised linears
on their prefill route (dequantise once -> dense GEMM).
  python tools/dense_check.py [reps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import candle_vllm_b200 as pkg
from candle_vllm_b200 import synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
peak = 1667.1
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
except Exception:
    pass
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda"); g.manual_seed(0)


def timed(fn, flush):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                      # 256 MB > L2: operands come from HBM
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"dense GEMM y[m,n] = x[m,k] . W[n,k]^T, bf16 in, fp32 accumulate; peak = {peak:.0f} TFLOP/s (measured cuBLAS burst)")
for (m, n, k) in [(128, 4096, 4096), (512, 4096, 4096), (2048, 4096, 4096), (8192, 4096, 4096), (8192, 28672, 4096), (8192, 4096, 14336), (8192, 8192, 8192)]:
    x = torch.randn((m, k), device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn((n, k), device=dev, generator=g) * 0.02).to(torch.bfloat16)
    lin = pkg.Linear(w)
    y = lin.forward(x)
    ref = x.float() @ w.float().t()
    err = ((y.float() - ref).norm() / ref.norm()).item()
    ms = timed(lambda: lin.forward(x), flush)
    ms_cublas = timed(lambda: torch.matmul(x, w.t()), flush)
    tf = 2.0 * m * n * k / ms / 1e9
    print(f"  m={m:5d} n={n:5d} k={k:5d}: {ms*1e3:8.1f} us = {tf:7.1f} TFLOP/s ({tf/peak:.2f} of peak), cuBLAS {2.0*m*n*k/ms_cublas/1e9:7.1f}; rel-Fro {err:.1e}")
    del x, w, lin, y, ref

print("quantised linears at prefill sizes (dequantise W once into scratch, dense GEMM)")
from candle_vllm_b200 import GgmlType
for (m, n, k, t) in [(2048, 4096, 4096, GgmlType.Q4_K), (8192, 4096, 4096, GgmlType.Q4_K), (8192, 14336, 4096, GgmlType.Q4_K), (8192, 4096, 4096, GgmlType.Q6_K)]:
    qm = pkg.QMatMul(synthetic.random_qtensor(g, t, n, k, dev))
    x = torch.randn((m, k), device=dev, generator=g)
    y = qm.forward(x)
    ms = timed(lambda: qm.forward(x), flush)
    tf = 2.0 * m * n * k / ms / 1e9
    print(f"  ggml type {int(t)} m={m:5d} n={n:5d} k={k:5d}: {ms*1e3:8.1f} us = {tf:7.1f} TFLOP/s ({tf/peak:.2f} of peak)")
    del qm, x, y

print("chunked-prefill attention on the paged cache (Llama-3-8B heads: 32 q / 8 kv, head 128, block 64), bf16")
for (nseq, qlen, cached) in [(4, 2048, 0), (1, 8192, 0), (4, 2048, 2048)]:
    H, kvh, hd, bs = 32, 8, 128, 64
    klen = qlen + cached
    nblk = -(-klen // bs)
    nb = nseq * nblk + 2
    perm = np.random.default_rng(1).permutation(nb)
    tables = [[int(v) for v in perm[i * nblk:(i + 1) * nblk]] for i in range(nseq)]
    prep = pkg.prepare_prompt([list(range(klen))] * nseq, tables, bs, [cached] * nseq, chunk_size=qlen)
    T = len(prep["tokens"])
    mk = lambda *shape: torch.randn(shape, device=dev, generator=g).to(torch.bfloat16)
    kc, vc = mk(nb, bs, kvh, hd), mk(nb, bs, kvh, hd)
    q, k, v = mk(T, H, hd), mk(T, kvh, hd), mk(T, kvh, hd)
    _, _, meta = pkg.inputs.to_device(prep)
    attn = pkg.PagedAttention(H, hd, hd ** -0.5, kvh)
    ms = timed(lambda: attn.forward(q, k, v, None, kc, vc, meta), flush)
    # causal: query i of a chunk sees cached + i + 1 keys; 4 flop per (query, key, dim) pair (QK^T and PV)
    pairs = nseq * (qlen * cached + qlen * (qlen + 1) // 2)
    tf = 4.0 * pairs * H * hd / ms / 1e9
    print(f"  {nseq} x {qlen} new tokens on {cached} cached: {ms*1e3:8.1f} us = {tf:7.1f} TFLOP/s causal-exact ({tf/peak:.2f} of peak; incl. the cache write)")
