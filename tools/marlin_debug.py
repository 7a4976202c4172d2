import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import candle_vllm_b200 as pkg
from oracle import gptq as OG
m, k, n, g = int(os.environ.get("M", 32)), int(os.environ.get("K", 4096)), int(os.environ.get("N", 1024)), 128
rng = np.random.default_rng(0)
q = rng.integers(0, 16, (k, n), dtype=np.uint8)
scales = rng.uniform(0.005, 0.02, (k // g, n)).astype(np.float32)
st = torch.from_numpy(scales).cuda().half()
qw = torch.from_numpy(OG.pack_gptq(q).view(np.int32)).cuda()
w_m = pkg.marlin_weight_repack(qw, 4, False)
s_m = pkg.marlin_permute_scales(st, k, n, g)
ws = torch.zeros(n, dtype=torch.int32, device="cuda")
W = OG.dequant_gptq(OG.pack_gptq(q), st.float().cpu().numpy(), g)      # [n, k]
for sl in list(range(16)) + [-1]:
    x = np.zeros((m, k), np.float32)
    if sl >= 0: x[:, sl * 256:(sl + 1) * 256] = rng.standard_normal((m, 256))
    else: x = rng.standard_normal((m, k)).astype(np.float32)
    xt = torch.from_numpy(x).cuda().half()
    y = pkg.gptq_matmul(xt, w_m, s_m, None, None, ws, 4, g).float().cpu().numpy()
    ref = xt.float().cpu().numpy().astype(np.float64) @ W.T
    err = np.abs(y - ref)
    bad_cols = np.where(err.max(axis=0) > 0.05 * np.abs(ref).max())[0]
    bad_rows = np.where(err.max(axis=1) > 0.05 * np.abs(ref).max())[0]
    print(f"slice {sl:2d}: rel {np.linalg.norm(y-ref)/np.linalg.norm(ref):.3e}  bad cols {len(bad_cols)} {bad_cols[:6]} bad rows {len(bad_rows)} {bad_rows[:6]}")

# ---- isolate: is the repacked weight tensor what the layout definition says?
def repack_emul(inw, k_packed, n):
    total = k_packed * n
    i = np.arange(total, dtype=np.int64)
    col = i % n; wk = i // n; c = wk >> 3; w = wk & 7
    kp_lo = 8 * c + (w >> 1); kp_hi = kp_lo + 4; sh = (16 * (w & 1)).astype(np.uint32)
    lo = inw[kp_lo, col] >> sh; hi = inw[kp_hi, col] >> sh
    o = np.zeros(total, np.uint32)
    for j in range(4):
        o |= (((lo >> np.uint32(4 * j)) & 0xF) | (((hi >> np.uint32(4 * j)) & 0xF) << 4)) << np.uint32(8 * j)
    out = np.zeros((n, k_packed), np.uint32); out[col, wk] = o
    return out
got = w_m.cpu().numpy().view(np.uint32).reshape(n, k // 8)
exp = repack_emul(OG.pack_gptq(q), k // 8, n)
bad = np.argwhere(got != exp)
print("repack mismatches:", len(bad), bad[:5], "bad word-columns (k/8):", sorted(set(bad[:, 1].tolist()))[:20] if len(bad) else [])
