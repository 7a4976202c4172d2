"""Timing of the MoE and MLA operators through the C ABI at the shapes of BASELINE configs 4 and 5 (profiling helper).
  python tools/moe_mla_check.py
MoE: Qwen3-30B-A3B (hidden 2048, 128 experts, top-8, expert FFN 768), Q4_K experts, batch 16 / 32 / 256: bytes of the experts that
were hit per grouped GEMM against the measured HBM peak.  MLA: DeepSeek-V3 absorbed decode (kv_lora 512 + rope 64), 128 heads (TP 1) and
16 heads (TP 8), batch 32, ctx 4096: bytes of the compressed cache read against the peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import candle_vllm_b200 as pkg
from candle_vllm_b200 import moe, mla, synthetic, GgmlType

dev = torch.device("cuda:0")
peak = 6583.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device="cuda"); g.manual_seed(0)


def timed(fn, reps=8):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


print(f"fused MoE, Qwen3-30B-A3B shapes (H 2048, 128 experts, top-8, I 768), Q4_K experts; HBM peak {peak:.0f} GB/s")
E, H, I, K = 128, 2048, 768, 8
ge = synthetic.random_qtensor(g, GgmlType.Q4_K, E * I, H, dev).data
ue = synthetic.random_qtensor(g, GgmlType.Q4_K, E * I, H, dev).data
de = synthetic.random_qtensor(g, GgmlType.Q4_K, E * H, I, dev).data
gate = torch.randn((E, H), device=dev, generator=g) * 0.05
for T in (16, 32, 256):
    xs = torch.randn((T, H), device=dev, generator=g)
    w, ids = moe.topk_softmax(xs @ gate.t(), K)
    eids, sids = moe.sort_expert_assignments(ids, E)
    hit = int(torch.unique(ids).numel())
    down_in = torch.randn((T * K, I), device=dev, generator=g)
    t_route = timed(lambda: moe.sort_expert_assignments(moe.topk_softmax(xs @ gate.t(), K)[1], E))
    t_up = timed(lambda: moe.moe_gemm_gguf(xs, ge, GgmlType.Q4_K, (E, I, H), None, sids, eids, K))
    t_dn = timed(lambda: moe.moe_gemm_gguf(down_in, de, GgmlType.Q4_K, (E, H, I), w, sids, eids, K))
    b_up = hit * I * H * 144 // 256
    b_dn = hit * H * I * 144 // 256
    print(f"  T={T:4d}: {hit:3d} experts hit; router+topk+sort {t_route*1e3:6.1f} us; gate/up GEMM {t_up*1e3:6.1f} us = {b_up/t_up/1e6:6.0f} GB/s "
          f"({b_up/t_up/1e6/peak:.2f}); down GEMM {t_dn*1e3:6.1f} us = {b_dn/t_dn/1e6:6.0f} GB/s ({b_dn/t_dn/1e6/peak:.2f})  [eager calls incl. gather / scatter]")

print("MLA absorbed decode (DeepSeek-V3: kv_lora 512 + rope 64), bf16 cache, batch 32, ctx 4096, block 64")
R, P, BS, B, ctx = 512, 64, 64, 32, 4096
nblk = ctx // BS
nb = B * nblk + 2
perm = np.random.default_rng(0).permutation(nb)
bt = torch.from_numpy(perm[:B * nblk].reshape(B, nblk).astype(np.int32)).to(dev)
cl = torch.full((B,), ctx, dtype=torch.int32, device=dev)
caches = [(torch.randn((nb, BS, 1, R), device=dev, generator=g).to(torch.bfloat16), torch.randn((nb, BS, 1, P), device=dev, generator=g).to(torch.bfloat16))
          for _ in range(4)]
for Hh in (128, 16):
    qa = (torch.randn((B, Hh, R), device=dev, generator=g) * 0.2).to(torch.bfloat16)
    qp = (torch.randn((B, Hh, P), device=dev, generator=g) * 0.2).to(torch.bfloat16)
    it = [0]

    def run():
        cc, pc = caches[it[0] % 4]; it[0] += 1
        mla.mla_paged_decode(qa, qp, cc, pc, bt, cl, (128 + 64) ** -0.5)

    ms = timed(run)
    byts = B * ctx * (R + P) * 2
    flops = 2.0 * B * Hh * ctx * (R + P + R)
    print(f"  {Hh:3d} heads: {ms*1e3:7.1f} us; cache bytes {byts/1e6:.0f} MB -> {byts/ms/1e6:6.0f} GB/s ({byts/ms/1e6/peak:.2f} of HBM peak); "
          f"{flops/ms/1e9:6.1f} TFLOP/s (the shape is compute-bound on mma.sync at 128 heads: {flops/byts:.0f} flop/B)")
