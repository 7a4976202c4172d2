"""KV-cache allocator and block-op entry points: mirror of ``CacheEngine``
(/root/reference/src/scheduler/cache_engine.rs).

  allocate_kv_cache   cache_engine.rs:122-294   (flash layout by default, legacy paged optional)
  shapes              cache_engine.rs:298-341
  swap_in / swap_out  cache_engine.rs:345-385   -> swap_blocks (K5), K then V of every layer
  copy                cache_engine.rs:387-399   -> copy_blocks (K4)

Unlike the reference the block ops must be applied on EVERY tensor-parallel rank:
each rank owns its kv-head shard of every block.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

from ._lib import BackendError
from .backend import KvLayout, copy_blocks, swap_blocks


@dataclass
class CacheConfig:
    block_size: int
    num_gpu_blocks: int
    num_cpu_blocks: int = 0
    fully_init: bool = True
    kvcache_dtype: str = "auto"        # "auto" (model dtype) | "fp8" (e4m3 stored as u8, main.rs:263-267)


def kv_head_shard(num_kv_heads: int, rank: int, world: int) -> Tuple[int, int]:
    """(/root/reference/src/openai/distributed.rs:725-765) -> (local kv heads, first global head)."""
    if world <= 1:
        return num_kv_heads, 0
    if num_kv_heads >= world:
        if num_kv_heads % world:
            raise BackendError(f"num_key_value_heads {num_kv_heads} not divisible by world {world}")
        n = num_kv_heads // world
        return n, rank * n
    if world % num_kv_heads:
        raise BackendError(f"world {world} not divisible by num_key_value_heads {num_kv_heads}")
    return 1, rank // (world // num_kv_heads)      # replicate each head over world/kvh ranks


def allocate_mla_cache(num_layers: int, num_blocks: int, block_size: int, kv_lora_rank: int, qk_rope_head_dim: int,
                       dtype: torch.dtype = torch.bfloat16, device="cuda") -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """MLA models (cache_engine.rs:172-185): per layer the compressed KV ``[nb, bs, 1, kv_lora_rank]`` and the rope keys
    ``[nb, bs, 1, qk_rope_head_dim]`` -- what ``mla.concat_and_cache_mla`` writes and ``mla.mla_paged_decode`` reads."""
    return [(torch.zeros((num_blocks, block_size, 1, kv_lora_rank), dtype=dtype, device=device),
             torch.zeros((num_blocks, block_size, 1, qk_rope_head_dim), dtype=dtype, device=device)) for _ in range(num_layers)]


def turboquant_layer_shapes(mode: str, num_blocks: int, block_size: int, num_kv_heads: int, head_dim: int, num_shards: int = 1) -> Dict[str, tuple]:
    """STORAGE shapes of the TurboQuant KV cache (cache_engine.rs:401-482); ``mode`` in {"turbo8", "turbo4", "turbo3"}.  Only the
    allocation is visible in the reference -- the transform (WHT + absmax quantisation) lives in attention-rs -- so this backend
    offers the shapes for sizing and block bookkeeping, and no kernels."""
    if mode not in ("turbo8", "turbo4", "turbo3"):
        raise BackendError(f"unknown TurboQuant mode {mode!r}")
    kvh = max(num_kv_heads // max(num_shards, 1), 1)
    shapes = {"v_absmax": (num_blocks, block_size, kvh), "v_quant": (num_blocks, block_size, kvh, head_dim // 2)}
    if mode == "turbo4":
        shapes.update(k_absmax=(num_blocks, block_size, kvh), k_quant=(num_blocks, block_size, kvh, head_dim // 2))
    elif mode == "turbo3":
        shapes.update(k_absmax=(num_blocks, block_size, kvh), k_quant=(num_blocks, block_size, kvh, (head_dim * 3 + 7) // 8))
    return shapes            # turbo8: K stays in the regular FP8 cache (no k_* tensors), as upstream


class CacheEngine:
    def __init__(self, num_layers: int, num_kv_heads: int, head_dim: int, cache_config: CacheConfig,
                 dtype: torch.dtype = torch.bfloat16, device="cuda", num_shards: int = 1,
                 layout: int = KvLayout.FLASH, pin_cpu: bool = True):
        self.num_layers = num_layers
        self.head_dim = head_dim
        self.num_kv_heads = max(num_kv_heads // max(num_shards, 1), 1)       # cache_engine.rs:306,320,335
        self.cfg = cache_config
        self.block_size = cache_config.block_size
        self.layout = layout
        self.device = torch.device(device)
        self.dtype = torch.uint8 if cache_config.kvcache_dtype == "fp8" else dtype
        self.cpu_swap_enabled = cache_config.num_cpu_blocks > 0
        self.gpu_cache = self._allocate(cache_config.num_gpu_blocks, self.device, False)
        self.cpu_cache = self._allocate(cache_config.num_cpu_blocks, torch.device("cpu"), pin_cpu) if self.cpu_swap_enabled else []

    # shapes: cache_engine.rs:298-341
    def key_block_shape(self):
        if self.layout == KvLayout.FLASH:
            return (self.block_size, self.num_kv_heads, self.head_dim)
        x = 16 // torch.empty((), dtype=self.dtype).element_size()
        return (self.num_kv_heads, self.head_dim // x, self.block_size, x)

    def value_block_shape(self):
        if self.layout == KvLayout.FLASH:
            return (self.block_size, self.num_kv_heads, self.head_dim)
        return (self.num_kv_heads, self.head_dim, self.block_size)

    def _allocate(self, num_blocks: int, device, pin: bool) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        out = []
        for _ in range(self.num_layers):
            kw = dict(dtype=self.dtype, device=device)
            if device.type == "cpu" and pin and torch.cuda.is_available():
                kw["pin_memory"] = True
            k = torch.zeros((num_blocks,) + self.key_block_shape(), **kw)
            v = torch.zeros((num_blocks,) + self.value_block_shape(), **kw)
            out.append((k, v))
        return out

    def get_kv_cache(self):
        return self.gpu_cache

    def bytes_per_block(self) -> int:
        k, v = self.gpu_cache[0]
        return (k[0].numel() + v[0].numel()) * k.element_size() * self.num_layers

    def swap_in(self, src_to_dst: Dict[int, int]) -> int:
        if not self.cpu_swap_enabled:
            raise BackendError("CPU KV cache swap-in is disabled for this device")
        nbytes = 0
        for (sk, sv), (dk, dv) in zip(self.cpu_cache, self.gpu_cache):
            swap_blocks(sk, dk, src_to_dst)
            swap_blocks(sv, dv, src_to_dst)
            nbytes += (sk[0].numel() + sv[0].numel()) * sk.element_size() * len(src_to_dst)
        return nbytes

    def swap_out(self, src_to_dst: Dict[int, int]) -> int:
        if not self.cpu_swap_enabled:
            raise BackendError("CPU KV cache swap-out is disabled for this device")
        nbytes = 0
        for (sk, sv), (dk, dv) in zip(self.gpu_cache, self.cpu_cache):
            swap_blocks(sk, dk, src_to_dst)
            swap_blocks(sv, dv, src_to_dst)
            nbytes += (sk[0].numel() + sv[0].numel()) * sk.element_size() * len(src_to_dst)
        return nbytes

    def copy(self, src_to_dst: Dict[int, List[int]]) -> None:
        copy_blocks([k for k, _ in self.gpu_cache], [v for _, v in self.gpu_cache], src_to_dst)
