"""Stand-alone check/timing of paged attention decode through the C ABI (debug + profiling helper).
  python tools/attn_check.py [B] [ctx] [heads] [kv_heads] [reps] [fp8]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import candle_vllm_b200 as pkg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 100
H = int(sys.argv[3]) if len(sys.argv) > 3 else 32
kvh = int(sys.argv[4]) if len(sys.argv) > 4 else 8
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
fp8 = len(sys.argv) > 6 and sys.argv[6] == "fp8"
hd, bs = 128, 64
nblk = -(-ctx // bs)
nb = B * nblk + 4
g = torch.Generator(device="cuda"); g.manual_seed(0)
def mk():
    t = torch.randn((nb, bs, kvh, hd), device="cuda", generator=g)
    return t.to(torch.float8_e4m3fn).view(torch.uint8) if fp8 else t.to(torch.bfloat16)
def f32(t):
    return t.view(torch.float8_e4m3fn).float() if fp8 else t.float()
kc, vc = mk(), mk()
q = torch.randn((B, H, hd), device="cuda", generator=g).to(torch.bfloat16)
bt = torch.from_numpy(np.random.default_rng(0).permutation(nb)[:B * nblk].reshape(B, nblk).astype(np.int32)).cuda()
cl = torch.full((B,), ctx, dtype=torch.int32, device="cuda")
attn = pkg.PagedAttention(H, hd, hd ** -0.5, kvh, fp8_kvcache=fp8)
meta = pkg.InputMetadata(False, torch.zeros(0, dtype=torch.int64, device="cuda"), bt, cl)
out = attn.forward(q, None, None, None, kc, vc, meta)
torch.cuda.synchronize()
# torch reference (plumbing-only check, fp32)
ref = torch.empty_like(out, dtype=torch.float32)
for b in range(min(B, 4)):
    k = f32(kc[bt[b].long()]).reshape(-1, kvh, hd)[:ctx]; v = f32(vc[bt[b].long()]).reshape(-1, kvh, hd)[:ctx]
    k = k.repeat_interleave(H // kvh, dim=1); v = v.repeat_interleave(H // kvh, dim=1)
    s = torch.einsum("hd,khd->hk", q[b].float(), k) * hd ** -0.5
    ref[b] = torch.einsum("hk,khd->hd", torch.softmax(s, -1), v)
nchk = min(B, 4)
err = (out[:nchk].float() - ref[:nchk]).abs().max().item()
print(f"B={B} ctx={ctx} H={H} kvh={kvh}: max abs err vs torch fp32 = {err:.3e}")
if reps:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L = 8
    kcs = [mk() for _ in range(L)]
    vcs = [mk() for _ in range(L)]
    for i in range(3):
        attn.forward(q, None, None, None, kcs[i % L], vcs[i % L], meta)
    torch.cuda.synchronize(); e0.record()
    for i in range(reps):
        attn.forward(q, None, None, None, kcs[i % L], vcs[i % L], meta)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = B * ctx * 2 * kvh * hd * (1 if fp8 else 2)
    print(f"  {ms*1e3:.1f} us/call, {byts/ms/1e6:.1f} GB/s KV read")
