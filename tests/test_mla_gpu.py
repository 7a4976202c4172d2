"""GPU parity of multi-head latent attention: concat_and_cache_mla (bit-exact), absorbed MLA decode on the tensor-core split-KV
kernel and causal prefill on the generic kernel, against oracle/mla.py.  Tolerance as for paged attention (16-bit P and output)."""
import numpy as np
import pytest
import torch

from candle_vllm_b200 import mla
from oracle import mla as OM
from tests.gpu_util import DEV, rel_fro

pytestmark = pytest.mark.gpu
R, P, BS = 512, 64, 64


def _setup(rng, dtype, ctx, H, extra_rows=None):
    B = len(ctx)
    nblk = [-(-c // BS) for c in ctx]
    nb = sum(nblk) + 2
    perm = rng.permutation(nb)
    W = max(nblk)
    bt = np.zeros((B, W), np.int32); o = 0
    for b in range(B):
        bt[b, :nblk[b]] = perm[o:o + nblk[b]]; o += nblk[b]
    mk = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(DEV).to(dtype)
    cc, pc = mk(nb, BS, 1, R), mk(nb, BS, 1, P)
    cc[perm[-1]] = float("nan"); pc[perm[-1]] = float("nan")          # an unowned block must never be read
    rows = B if extra_rows is None else extra_rows
    qa, qp = mk(rows, H, R) * 0.2, mk(rows, H, P) * 0.2
    return bt, cc, pc, qa, qp


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("H,ctx", [(16, [1, 64, 65, 300]), (128, [700, 33]), (5, [31, 32, 33, 257, 512]), (16, [4096 + 31 * i for i in range(8)])])
def test_mla_decode_matches_oracle(dtype, H, ctx):
    rng = np.random.default_rng(H + len(ctx))
    bt, cc, pc, qa, qp = _setup(rng, dtype, ctx, H)
    cl = torch.tensor(ctx, dtype=torch.int32, device=DEV)
    scale = (128 + 64) ** -0.5
    out = mla.mla_paged_decode(qa, qp, cc, pc, torch.from_numpy(bt).to(DEV), cl, scale)
    ref = OM.attend(qa.float().cpu().numpy(), qp.float().cpu().numpy(), cc.float().cpu().numpy()[:, :, 0], pc.float().cpu().numpy()[:, :, 0], bt, ctx, scale)
    o = out.float().cpu().numpy()
    assert np.isfinite(o).all() and rel_fro(o, ref) < 4e-3, rel_fro(o, ref)
    assert torch.equal(out, mla.mla_paged_decode(qa, qp, cc, pc, torch.from_numpy(bt).to(DEV), cl, scale))


def test_mla_prefill_and_cache_write():
    rng = np.random.default_rng(9)
    dtype, H = torch.bfloat16, 8
    qlens, cached = [70, 5, 130], [20, 0, 64]
    klens = [q + c for q, c in zip(qlens, cached)]
    T = sum(qlens)
    bt, cc, pc, qa, qp = _setup(rng, dtype, klens, H, extra_rows=T)
    mk = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).to(DEV).to(dtype)
    ckv, kpe = mk(T, R), mk(T, P)
    slots, cu = [], [0]
    for i, (q, c) in enumerate(zip(qlens, cached)):
        slots += [int(bt[i, p // BS]) * BS + p % BS for p in range(c, c + q)]
        cu.append(cu[-1] + q)
    slots[3] = -1                                                       # a padded token is skipped
    cn, pn = cc.float().cpu().numpy()[:, :, 0].copy(), pc.float().cpu().numpy()[:, :, 0].copy()
    mla.concat_and_cache_mla(ckv, kpe, cc, pc, torch.tensor(slots, dtype=torch.int64, device=DEV))
    OM.concat_and_cache(ckv.float().cpu().numpy(), kpe.float().cpu().numpy(), cn, pn, slots)
    got_c, got_p = cc.float().cpu().numpy()[:, :, 0], pc.float().cpu().numpy()[:, :, 0]
    assert np.array_equal(np.nan_to_num(got_c), np.nan_to_num(cn)) and np.array_equal(np.nan_to_num(got_p), np.nan_to_num(pn))      # bit-exact cache write
    scale = 0.07
    out = mla.mla_paged_prefill(qa, qp, cc, pc, torch.from_numpy(bt).to(DEV), torch.tensor(klens, dtype=torch.int32, device=DEV),
                                torch.tensor(cu, dtype=torch.int32, device=DEV), scale)
    ref = OM.attend(qa.float().cpu().numpy(), qp.float().cpu().numpy(), got_c, got_p, bt, klens, scale, cu)
    assert rel_fro(out.float().cpu().numpy(), ref) < 4e-3
