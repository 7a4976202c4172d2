"""GPU parity (bit-exact): copy_blocks (K4), swap_blocks (K5), reshape_and_cache (K3) through the
C ABI against the oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

import candle_vllm_b200 as pkg
from oracle import cache_ops as OC
from oracle import llama as LL
from tests.gpu_util import DEV, bits, to_bf16_t

pytestmark = pytest.mark.gpu


def _caches(rng, L, nb, shape, dtype):
    ks, vs = [], []
    for _ in range(L):
        if dtype == torch.uint8:
            ks.append(torch.from_numpy(rng.integers(0, 256, (nb,) + shape, dtype=np.uint8)).to(DEV))
            vs.append(torch.from_numpy(rng.integers(0, 256, (nb,) + shape, dtype=np.uint8)).to(DEV))
        else:
            ks.append(torch.from_numpy(rng.standard_normal((nb,) + shape).astype(np.float32)).to(DEV).to(dtype))
            vs.append(torch.from_numpy(rng.standard_normal((nb,) + shape).astype(np.float32)).to(DEV).to(dtype))
    return ks, vs


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32, torch.uint8])
@pytest.mark.parametrize("shape,mapping", [
    ((64, 8, 128), {1: [3, 4], 0: [5], 7: [2]}),
    ((16, 2, 64), {0: [1]}),
    ((4, 1, 24), {2: [0, 1, 3, 4, 5, 6, 7]}),            # fan-out of one source
    ((3, 1, 5), {6: [0]}),                               # bytes_per_block not a multiple of 16
])
def test_copy_blocks_bit_exact(dtype, shape, mapping):
    rng = np.random.default_rng(0)
    ks, vs = _caches(rng, 3, 8, shape, dtype)
    kn = [bits(k).copy() for k in ks]; vn = [bits(v).copy() for v in vs]
    pkg.copy_blocks(ks, vs, mapping)
    torch.cuda.synchronize()
    OC.copy_blocks(kn, vn, mapping)
    for l in range(3):
        assert np.array_equal(bits(ks[l]), kn[l]) and np.array_equal(bits(vs[l]), vn[l])


def test_copy_blocks_many_pairs_and_layers():
    rng = np.random.default_rng(1)
    L, nb = 40, 600
    ks, vs = _caches(rng, L, nb, (4, 2, 16), torch.bfloat16)
    perm = rng.permutation(nb)
    mapping = {int(perm[i]): [int(perm[300 + i])] for i in range(300)}       # > kMaxParamPairs: chunked launches
    kn = [bits(k).copy() for k in ks]; vn = [bits(v).copy() for v in vs]
    pkg.copy_blocks(ks, vs, mapping)
    torch.cuda.synchronize()
    OC.copy_blocks(kn, vn, mapping)
    for l in range(L):
        assert np.array_equal(bits(ks[l]), kn[l]) and np.array_equal(bits(vs[l]), vn[l])


def test_copy_blocks_errors_like_reference():
    k = torch.zeros(4, 4, 1, 8, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(pkg.BackendError, match="different types"):
        pkg.copy_blocks([k], [k.float()], {0: [1]})
    with pytest.raises(pkg.BackendError, match="data type supported"):
        pkg.copy_blocks([k.to(torch.int32)], [k.to(torch.int32)], {0: [1]})
    pkg.copy_blocks([], [], {0: [1]})       # zero layers: Ok(())


def test_copy_blocks_full_size_idempotent_and_checksum():
    # Llama-3-8B block: 64*8*128 bf16 = 131072 B per K block; property: copying twice == once,
    # untouched blocks keep their checksum.
    rng = np.random.default_rng(2)
    ks, vs = _caches(rng, 32, 24, (64, 8, 128), torch.bfloat16)
    before = [k.view(torch.int16).long().sum(dim=(1, 2, 3)).cpu() for k in ks]
    mapping = {0: [10, 11], 3: [12]}
    pkg.copy_blocks(ks, vs, mapping)
    once = [k.clone() for k in ks]
    pkg.copy_blocks(ks, vs, mapping)
    torch.cuda.synchronize()
    for l in range(32):
        assert torch.equal(ks[l], once[l])
        after = ks[l].view(torch.int16).long().sum(dim=(1, 2, 3)).cpu()
        assert after[10] == before[l][0] and after[11] == before[l][0] and after[12] == before[l][3]
        keep = [i for i in range(24) if i not in (10, 11, 12)]
        assert torch.equal(after[keep], before[l][keep])


@pytest.mark.parametrize("pinned", [True, False])
def test_swap_blocks_roundtrip_bit_exact(pinned):
    rng = np.random.default_rng(3)
    gpu = torch.from_numpy(rng.standard_normal((16, 64, 8, 128)).astype(np.float32)).to(DEV).to(torch.bfloat16)
    cpu = torch.zeros((8, 64, 8, 128), dtype=torch.bfloat16, pin_memory=pinned)
    out_map = {1: 0, 2: 1, 3: 2, 9: 5, 15: 7}           # has a coalescable run 1,2,3 -> 0,1,2
    pkg.swap_blocks(gpu, cpu, out_map)                  # swap out (D2H)
    torch.cuda.synchronize()
    exp = np.zeros((8, 64, 8, 128, 2), np.uint8).reshape(8, -1)
    g = bits(gpu).reshape(16, -1)
    OC.swap_blocks(g, exp, out_map)
    assert np.array_equal(bits(cpu).reshape(8, -1), exp)
    gpu2 = torch.zeros_like(gpu)
    pkg.swap_blocks(cpu, gpu2, {v: k for k, v in out_map.items()})    # swap in (H2D)
    torch.cuda.synchronize()
    for s in out_map:
        assert torch.equal(gpu2[s], gpu[s])
    pkg.swap_blocks(gpu, gpu2, {0: 4})                                 # D2D
    torch.cuda.synchronize()
    assert torch.equal(gpu2[4], gpu[0])
    with pytest.raises(pkg.BackendError, match="out of range"):
        pkg.swap_blocks(gpu, cpu, {0: 99})


@pytest.mark.parametrize("layout", ["flash", "paged"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("fp8", [False, True])
def test_reshape_and_cache_bit_exact(layout, dtype, fp8):
    rng = np.random.default_rng(4)
    T, kvh, hd, bs, nb = 37, 8, 128, 64, 6
    k = torch.from_numpy(rng.standard_normal((T, kvh, hd)).astype(np.float32) * 3).to(DEV).to(dtype)
    v = torch.from_numpy(rng.standard_normal((T, kvh, hd)).astype(np.float32) * 3).to(DEV).to(dtype)
    slots = rng.permutation(nb * bs)[:T].astype(np.int64)
    slots[[3, 20]] = -1                                  # pad slots are skipped
    cdt = torch.uint8 if fp8 else dtype
    esz = 1 if fp8 else 2
    if layout == "flash":
        kshape, vshape = OC.flash_kv_shape(nb, bs, kvh, hd), OC.flash_kv_shape(nb, bs, kvh, hd)
    else:
        kshape, vshape = OC.paged_k_shape(nb, bs, kvh, hd, esz), OC.paged_v_shape(nb, bs, kvh, hd)
    kc = torch.zeros(kshape, dtype=cdt, device=DEV); vc = torch.zeros(vshape, dtype=cdt, device=DEV)
    pkg.reshape_and_cache(k, v, kc, vc, torch.from_numpy(slots).to(DEV), fp8=fp8)
    torch.cuda.synchronize()
    kn, vn = k.float().cpu().numpy(), v.float().cpu().numpy()
    ke = np.zeros(kshape, np.uint8 if fp8 else np.float32); ve = np.zeros(vshape, np.uint8 if fp8 else np.float32)
    (OC.reshape_and_cache_flash if layout == "flash" else OC.reshape_and_cache_paged)(kn, vn, ke, ve, slots, fp8=fp8)
    if fp8:
        assert np.array_equal(kc.cpu().numpy(), ke) and np.array_equal(vc.cpu().numpy(), ve)
    else:
        assert np.array_equal(kc.float().cpu().numpy(), ke) and np.array_equal(vc.float().cpu().numpy(), ve)


def test_reshape_and_cache_f32_input_and_strided():
    rng = np.random.default_rng(5)
    T, kvh, hd, bs, nb = 9, 2, 64, 16, 3
    big = torch.from_numpy(rng.standard_normal((T, 3 * kvh * hd)).astype(np.float32)).to(DEV)
    k = big[:, :kvh * hd].view(T, kvh, hd)                 # row-strided views of a packed qkv
    v = big[:, kvh * hd:2 * kvh * hd].view(T, kvh, hd)
    slots = torch.arange(T, device=DEV, dtype=torch.int64) * 2
    kc = torch.zeros((nb, bs, kvh, hd), dtype=torch.bfloat16, device=DEV); vc = torch.zeros_like(kc)
    pkg.reshape_and_cache(k, v, kc, vc, slots)
    torch.cuda.synchronize()
    flat = kc.view(-1, kvh, hd)
    assert torch.equal(flat[slots], k.to(torch.bfloat16)) and torch.equal(vc.view(-1, kvh, hd)[slots], v.to(torch.bfloat16))


def test_cache_engine_swap_and_copy():
    cfg = pkg.CacheConfig(block_size=16, num_gpu_blocks=12, num_cpu_blocks=6)
    eng = pkg.CacheEngine(num_layers=2, num_kv_heads=2, head_dim=64, cache_config=cfg)
    for k, v in eng.gpu_cache:
        k.normal_(); v.normal_()
    ref = [(k.clone(), v.clone()) for k, v in eng.gpu_cache]
    n = eng.swap_out({3: 0, 4: 1})
    assert n == 2 * 2 * 2 * 16 * 2 * 64 * 2
    for k, v in eng.gpu_cache:
        k[3].zero_(); v[4].zero_()
    eng.swap_in({0: 3, 1: 4})
    eng.copy({3: [7, 8]})
    torch.cuda.synchronize()
    for (k, v), (rk, rv) in zip(eng.gpu_cache, ref):
        assert torch.equal(k[3], rk[3]) and torch.equal(v[4], rv[4]) and torch.equal(k[7], rk[3]) and torch.equal(v[8], rv[3])
