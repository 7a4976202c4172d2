"""Absorbed multi-head latent attention over the paged latent cache, numpy fp64.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.  PARITY UNPINNED: the kernels live in attention-rs.  Semantics follow the call sites
in /root/reference/src/openai/models/layers/mla_attention.rs:479-552: score = (q_absorbed . ckv + q_pe . kpe) * sm_scale over the keys gathered
through the block table, softmax, out = P . ckv (the caller applies W_uv); prefill is causal with the query rows being the LAST positions of the
sequence's context (same convention as inputs.rs:351-367)."""
from __future__ import annotations

import numpy as np


def _gather(cache, table, n):
    bs = cache.shape[1]
    nblk = -(-n // bs)
    return cache[np.asarray(table[:nblk], np.int64)].reshape(nblk * bs, cache.shape[-1])[:n].astype(np.float64)


def concat_and_cache(ckv, kpe, ckv_cache, kpe_cache, slots):
    nb, bs = ckv_cache.shape[:2]
    c = ckv_cache.reshape(nb * bs, -1); p = kpe_cache.reshape(nb * bs, -1)
    for t, s in enumerate(np.asarray(slots, np.int64)):
        if s >= 0:
            c[s] = ckv[t]; p[s] = kpe[t]


def attend(q_abs, q_pe, ckv_cache, kpe_cache, block_tables, context_lens, sm_scale, cu_seqlens_q=None):
    """q_abs [rows, H, R], q_pe [rows, H, P]; caches [nb, bs, R] / [nb, bs, P] -> [rows, H, R] f32."""
    q_abs = np.asarray(q_abs, np.float64); q_pe = np.asarray(q_pe, np.float64)
    out = np.zeros(q_abs.shape, np.float32)
    for i in range(len(context_lens)):
        L = int(context_lens[i])
        if L == 0:
            continue
        c = _gather(ckv_cache, block_tables[i], L); p = _gather(kpe_cache, block_tables[i], L)
        if cu_seqlens_q is None:
            rows, last = [i], [L - 1]
        else:
            q0, q1 = int(cu_seqlens_q[i]), int(cu_seqlens_q[i + 1])
            rows = list(range(q0, q1)); last = [L - (q1 - q0) + j for j in range(q1 - q0)]
        for r, lp in zip(rows, last):
            s = (q_abs[r] @ c[:lp + 1].T + q_pe[r] @ p[:lp + 1].T) * sm_scale        # [H, keys]
            s = s - s.max(axis=-1, keepdims=True)
            w = np.exp(s); w /= w.sum(axis=-1, keepdims=True)
            out[r] = w @ c[:lp + 1]
    return out
