"""Synthetic GGUF-LLaMA weights and KV state: no model files, no network.

Weights are random but VALID GGML blocks generated on the device with a seeded torch generator:
Q4_K: 12 scale bytes + 128 nibble bytes uniform, d ~ U(0.5,2)*2^-14, dmin ~ 7.5*U(0.5,2)*2^-14
(zero-mean weights of std ~0.02, so activations stay O(1) through 32 layers); Q6_K: 208 bytes
uniform, d ~ U(0.5,2)*2^-16.  Norm weights ~ U(0.5,1.5), embeddings N(0,1).
Tensor-parallel shards follow the reference: column split dim 0, row split dim 1 by raw blocks
(/root/reference/src/openai/models/layers/quantized_var_builder.rs:234-269).
"""
from __future__ import annotations

import torch

from .backend import GgmlType, QTensor
from .llama import LlamaConfig


def _f16_bytes(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float16).view(torch.uint8).reshape(*x.shape, 2)


def random_q4k(gen, n: int, k: int, device) -> torch.Tensor:
    nb = n * (k // 256)
    blocks = torch.randint(0, 256, (nb, 144), dtype=torch.uint8, device=device, generator=gen)
    d = (torch.rand(nb, device=device, generator=gen) * 1.5 + 0.5) * 2.0 ** -14
    dmin = (torch.rand(nb, device=device, generator=gen) * 1.5 + 0.5) * 7.5 * 2.0 ** -14
    blocks[:, 0:2] = _f16_bytes(d)
    blocks[:, 2:4] = _f16_bytes(dmin)
    return blocks.reshape(-1)


def random_q6k(gen, n: int, k: int, device) -> torch.Tensor:
    nb = n * (k // 256)
    blocks = torch.randint(0, 256, (nb, 210), dtype=torch.uint8, device=device, generator=gen)
    d = (torch.rand(nb, device=device, generator=gen) * 1.5 + 0.5) * 2.0 ** -16
    blocks[:, 208:210] = _f16_bytes(d)
    return blocks.reshape(-1)


def random_qtensor(gen, ggml_type: int, n: int, k: int, device) -> QTensor:
    fn = {GgmlType.Q4_K: random_q4k, GgmlType.Q6_K: random_q6k}[ggml_type]
    return QTensor(fn(gen, n, k, device), ggml_type, (n, k), allow_cpu=True)


def shard_rows(w: QTensor, rank: int, world: int) -> QTensor:
    """column-parallel: split dim 0 (distributed.rs:767-811)."""
    n, k = w.shape
    be, bb = GgmlType.BLOCK[w.ggml_type]
    rows = w.data.reshape(n, (k // be) * bb)
    nl = n // world
    return QTensor(rows[rank * nl:(rank + 1) * nl].contiguous().reshape(-1), w.ggml_type, (nl, k), allow_cpu=True)


def pad_rows(w: QTensor, n_pad: int) -> QTensor:
    """append all-zero GGML blocks (d = 0 -> weights 0) so that the tensor has n_pad rows."""
    n, k = w.shape
    if n_pad == n:
        return w
    be, bb = GgmlType.BLOCK[w.ggml_type]
    extra = torch.zeros((n_pad - n) * (k // be) * bb, dtype=torch.uint8, device=w.data.device)
    return QTensor(torch.cat([w.data, extra]), w.ggml_type, (n_pad, k), allow_cpu=True)


def shard_cols(w: QTensor, rank: int, world: int) -> QTensor:
    """row-parallel: split dim 1 along whole blocks (raw-byte shard, quantized_var_builder.rs:234-269)."""
    n, k = w.shape
    be, bb = GgmlType.BLOCK[w.ggml_type]
    nbk = k // be
    if nbk % world:
        raise ValueError(f"k={k}: {nbk} blocks per row not divisible by world {world}")
    blocks = w.data.reshape(n, nbk, bb)
    bl = nbk // world
    return QTensor(blocks[:, rank * bl:(rank + 1) * bl].contiguous().reshape(-1), w.ggml_type, (n, k // world), allow_cpu=True)


def make_weights(cfg: LlamaConfig, device="cuda", seed: int = 0, tp_rank: int = 0, tp_world: int = 1,
                 linear_type: int = GgmlType.Q4_K, output_type: int = GgmlType.Q6_K) -> dict:
    """Full (unsharded) tensors are drawn from the seeded stream on every rank, then sharded, so all
    ranks agree on the global model (embedding + norms replicated)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    H, hd = cfg.hidden, cfg.head_dim
    qd, kd = cfg.num_heads * hd, cfg.num_kv_heads * hd
    col = (lambda w: shard_rows(w, tp_rank, tp_world)) if tp_world > 1 else (lambda w: w)
    row = (lambda w: shard_cols(w, tp_rank, tp_world)) if tp_world > 1 else (lambda w: w)
    kcol = col
    if tp_world > 1 and cfg.num_kv_heads < tp_world:
        # kv_head_shard with fewer kv heads than ranks (/root/reference/src/openai/distributed.rs:753-764): every rank holds ONE
        # kv head, replicated over world / kvh consecutive ranks
        if tp_world % cfg.num_kv_heads:
            raise ValueError(f"world {tp_world} not a multiple of {cfg.num_kv_heads} kv heads")
        kcol = lambda w: shard_rows(w, tp_rank * cfg.num_kv_heads // tp_world, cfg.num_kv_heads)
    w = dict(tok_embeddings=torch.randn((cfg.vocab, H), device=device, generator=gen, dtype=torch.float32),
             norm=torch.rand(H, device=device, generator=gen) + 0.5, layers=[])
    for _ in range(cfg.num_layers):
        w["layers"].append(dict(
            attn_norm=torch.rand(H, device=device, generator=gen) + 0.5,
            ffn_norm=torch.rand(H, device=device, generator=gen) + 0.5,
            wq=col(random_qtensor(gen, linear_type, qd, H, device)),
            wk=kcol(random_qtensor(gen, linear_type, kd, H, device)),
            wv=kcol(random_qtensor(gen, linear_type, kd, H, device)),
            wo=row(random_qtensor(gen, linear_type, H, qd, device)),
            w1=col(random_qtensor(gen, linear_type, cfg.ffn, H, device)),
            w2=row(random_qtensor(gen, linear_type, H, cfg.ffn, device)),
            w3=col(random_qtensor(gen, linear_type, cfg.ffn, H, device)),
        ))
    out = random_qtensor(gen, output_type, cfg.vocab, H, device)
    if tp_world > 1:
        # vocab-parallel lm_head: rows padded with zero blocks to pad_vocab_size (distributed.rs:1448-1454), then split
        from .llama import padded_vocab
        out = shard_rows(pad_rows(out, padded_vocab(cfg.vocab, tp_world)), tp_rank, tp_world)
    w["output"] = out
    return w


def make_weights_16bit(cfg: LlamaConfig, device="cuda", seed: int = 0, dtype=torch.bfloat16):
    """Dense 16-bit Llama (BASELINE config 2; the safetensors path of the reference, llama.rs): every linear a [n, k] tensor of
    ``dtype`` ~ N(0, 0.02), lm_head included.  Returns (engine weights, oracle weights); the oracle sees exactly the 16-bit values."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    H, hd = cfg.hidden, cfg.head_dim
    qd, kd = cfg.num_heads * hd, cfg.num_kv_heads * hd
    mk = lambda n, k: (torch.randn((n, k), device=device, generator=gen, dtype=torch.float32) * 0.02).to(dtype).contiguous()
    w = dict(tok_embeddings=torch.randn((cfg.vocab, H), device=device, generator=gen, dtype=torch.float32),
             norm=torch.rand(H, device=device, generator=gen) + 0.5, layers=[])
    for _ in range(cfg.num_layers):
        w["layers"].append(dict(attn_norm=torch.rand(H, device=device, generator=gen) + 0.5, ffn_norm=torch.rand(H, device=device, generator=gen) + 0.5,
                                wq=mk(qd, H), wk=mk(kd, H), wv=mk(kd, H), wo=mk(H, qd), w1=mk(cfg.ffn, H), w2=mk(H, cfg.ffn), w3=mk(cfg.ffn, H)))
    w["output"] = mk(cfg.vocab, H)
    return w


def make_weights_gptq(cfg: LlamaConfig, device="cuda", seed: int = 0, group_size: int = 128, dtype=torch.float16, output_type: int = GgmlType.Q6_K,
                      with_oracle: bool = True):
    """GPTQ symmetric int4 Llama prepared for Marlin like the reference does at load time (linear.rs:300-413: ``gptq_repack`` +
    ``marlin_permute_scales``): BASELINE config 3.  int4 uniform, scales ~ U(0.005, 0.02) of ``dtype``; the lm_head stays a GGML tensor
    (the reference never quantises it to int4).  Returns (engine weights, oracle weights)."""
    import numpy as np
    from .gptq import marlin_permute_scales, marlin_weight_repack
    from .llama import MarlinWeight
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    H, hd = cfg.hidden, cfg.head_dim
    qd, kd = cfg.num_heads * hd, cfg.num_kv_heads * hd

    def mk(n, k):
        q = torch.randint(0, 16, (k, n), device=device, generator=gen, dtype=torch.int32)          # [K, N] like the checkpoint
        packed = torch.zeros((k // 8, n), dtype=torch.int32, device=device)
        for i in range(8):
            packed |= q[i::8] << (4 * i)                                                              # nibble i of word kp = q[8 kp + i]
        ng = 1 if group_size == -1 else k // group_size
        scales = (torch.rand((ng, n), device=device, generator=gen) * 0.015 + 0.005).to(dtype)
        eng = MarlinWeight(marlin_weight_repack(packed, 4, False), marlin_permute_scales(scales, k, n, group_size).contiguous(), group_size)
        orc = ("gptq", packed.cpu().numpy().view(np.uint32), scales.float().cpu().numpy(), group_size) if with_oracle else None
        return eng, orc

    w = dict(tok_embeddings=torch.randn((cfg.vocab, H), device=device, generator=gen, dtype=torch.float32),
             norm=torch.rand(H, device=device, generator=gen) + 0.5, layers=[])
    ow = dict(tok_embeddings=w["tok_embeddings"].cpu().numpy() if with_oracle else None, norm=w["norm"].cpu().numpy(), layers=[])
    for _ in range(cfg.num_layers):
        lw = dict(attn_norm=torch.rand(H, device=device, generator=gen) + 0.5, ffn_norm=torch.rand(H, device=device, generator=gen) + 0.5)
        lo = dict(attn_norm=lw["attn_norm"].cpu().numpy(), ffn_norm=lw["ffn_norm"].cpu().numpy())
        for name, (n, k) in dict(wq=(qd, H), wk=(kd, H), wv=(kd, H), wo=(H, qd), w1=(cfg.ffn, H), w2=(H, cfg.ffn), w3=(cfg.ffn, H)).items():
            lw[name], lo[name] = mk(n, k)
        w["layers"].append(lw); ow["layers"].append(lo)
    out = random_qtensor(gen, output_type, cfg.vocab, H, device)
    w["output"] = out
    ow["output"] = (out.data.cpu().numpy(), out.ggml_type, cfg.vocab, H) if with_oracle else None
    return w, ow


def fill_kv_cache(kv_cache, seed: int = 1, tp_rank: int = 0, tp_world: int = 1, num_kv_heads: int = 0) -> None:
    """KV contents N(0,1) in the cache dtype (bf16, or e4m3 bits for u8 caches).  With tp_world > 1 (and the model's total
    ``num_kv_heads``) every rank draws the FULL layer [nb, bs, kvh, hd] from the same seeded stream and keeps its own kv heads
    (kv_head_shard: a slice, or one replicated head when kvh < world), so the sharded caches are exactly the shards of the
    tp_world = 1 cache -- which is what lets bench.py compare a TP run with a TP = 1 run of the same seed."""
    dev = kv_cache[0][0].device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    for k, v in kv_cache:
        for t in (k, v):
            if tp_world > 1 and num_kv_heads > 0:
                nb, bs, kl, hd = t.shape
                full = torch.randn((nb, bs, num_kv_heads, hd), device=dev, generator=gen, dtype=torch.float32)
                h0 = tp_rank * kl if num_kv_heads >= tp_world else tp_rank * num_kv_heads // tp_world
                r = full[:, :, h0:h0 + kl, :]
                del full
            else:
                r = torch.randn(t.shape, device=dev, generator=gen, dtype=torch.float32)
            if t.dtype == torch.uint8:
                t.copy_(r.to(torch.float8_e4m3fn).view(torch.uint8))
            else:
                t.copy_(r)
            del r


def random_block_tables(num_seqs: int, blocks_per_seq: int, num_blocks: int, seed: int = 2):
    """Random permutation of physical blocks (non-contiguous pages), seeded; list of lists."""
    import numpy as np
    rng = np.random.default_rng(seed)
    perm = rng.permutation(num_blocks)[: num_seqs * blocks_per_seq]
    return perm.reshape(num_seqs, blocks_per_seq).tolist()
