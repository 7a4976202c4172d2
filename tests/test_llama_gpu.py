"""GPU parity: the whole decode step (GGUFLLaMa::forward_inner) through b200_llama_decode vs the
numpy oracle on a small synthetic model; greedy tokens and logits.  Tolerance: logits within 1e-3
(normalised by max|logit|, SURVEY.md §8c / north_star) of the fp oracle."""
import numpy as np
import pytest
import torch

import candle_vllm_b200 as pkg
from candle_vllm_b200 import synthetic
from oracle import llama as OL
from tests.gpu_util import DEV, weights_to_oracle

pytestmark = pytest.mark.gpu


def _small_cfg(**kw):
    d = dict(hidden=512, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=128, ffn=1024, vocab=768,
             max_pos=512, block_size=16, max_num_seqs=8, max_blocks_per_seq=16)
    d.update(kw)
    return pkg.LlamaConfig(**d)


def _ocfg(cfg):
    return dict(hidden=cfg.hidden, heads=cfg.num_heads, kv_heads=cfg.num_kv_heads, head_dim=cfg.head_dim,
                rms_eps=cfg.rms_eps, max_pos=cfg.max_pos, rope_theta=cfg.rope_theta)


@pytest.mark.parametrize("use_graph", [False, True])
def test_decode_steps_match_oracle(use_graph):
    cfg = _small_cfg()
    w = synthetic.make_weights(cfg, DEV, seed=0)
    ow = weights_to_oracle(w)
    nb = 40
    eng = pkg.CacheEngine(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim, pkg.CacheConfig(cfg.block_size, nb))
    synthetic.fill_kv_cache(eng.gpu_cache, seed=1)
    okc = [k.float().cpu().numpy().copy() for k, _ in eng.gpu_cache]
    ovc = [v.float().cpu().numpy().copy() for _, v in eng.gpu_cache]
    model = pkg.GGUFLLaMa(cfg, w, eng.gpu_cache, use_graph=use_graph)
    B = 5
    lens = [1, 16, 17, 40, 100]
    tables = synthetic.random_block_tables(B, 8, nb, seed=2)
    tokens = [5, 700, 33, 0, 123]
    for step in range(3):
        prep = pkg.prepare_decode(lens, tokens, tables, cfg.block_size)
        nxt, logits = model.decode(prep, want_logits=True)
        ref = OL.forward(_ocfg(cfg), ow, prep["tokens"].astype(np.int64), prep["positions"], okc, ovc,
                         dict(slot_mapping=prep["slot_mapping"], block_tables=prep["block_tables"], context_lens=prep["context_lens"]))
        scale = np.abs(ref).max()
        err = np.abs(logits - ref).max() / scale
        assert err < 1e-3, (step, err)
        # greedy tokens agree wherever the oracle's top-1 margin exceeds the logit error bound
        for b in range(B):
            if nxt[b] != ref[b].argmax():
                assert ref[b].max() - ref[b, nxt[b]] <= 2 * err * scale, (step, b)
        # the engine wrote this step's K/V into the paged cache exactly where the oracle did
        for l in range(cfg.num_layers):
            assert np.abs(eng.gpu_cache[l][0].float().cpu().numpy() - okc[l]).max() < 2e-2
        tokens = [int(t) for t in ref.argmax(axis=1)]
        lens = [L + 1 for L in lens]


def test_resident_replay_equals_host_driven_steps():
    cfg = _small_cfg()
    w = synthetic.make_weights(cfg, DEV, seed=3)
    nb = 40
    tables = synthetic.random_block_tables(4, 8, nb, seed=4)

    def run(resident):
        eng = pkg.CacheEngine(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim, pkg.CacheConfig(cfg.block_size, nb))
        synthetic.fill_kv_cache(eng.gpu_cache, seed=5)
        model = pkg.GGUFLLaMa(cfg, w, eng.gpu_cache)
        lens, toks, out = [10, 31, 32, 50], [1, 2, 3, 4], []
        nxt, _ = model.decode(pkg.prepare_decode(lens, toks, tables, cfg.block_size))
        out.append(nxt.copy())
        for _ in range(4):
            if resident:
                model.decode_resident(4, advance=True)
                nxt = model.read_next_tokens(4)
            else:
                lens = [L + 1 for L in lens]
                nxt, _ = model.decode(pkg.prepare_decode(lens, [int(x) for x in nxt], tables, cfg.block_size))
            out.append(np.asarray(nxt).copy())
        return np.stack(out)

    a, b = run(False), run(True)
    assert np.array_equal(a, b)


def test_engine_argument_errors():
    cfg = _small_cfg()
    w = synthetic.make_weights(cfg, DEV, seed=0)
    eng = pkg.CacheEngine(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim, pkg.CacheConfig(cfg.block_size, 8))
    model = pkg.GGUFLLaMa(cfg, w, eng.gpu_cache)
    prep = pkg.prepare_decode([3], [cfg.vocab + 5], [[1]], cfg.block_size)
    with pytest.raises(pkg.BackendError, match="vocab"):
        model.decode(prep)
    prep = pkg.prepare_decode([3] * 9, [1] * 9, [[1]] * 9, cfg.block_size)
    with pytest.raises(pkg.BackendError, match="num_seqs"):
        model.decode(prep)
