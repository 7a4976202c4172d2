"""CPU oracle for the candle-vllm batched-decode hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker.  The product
path (``candle-vllm_b200``) never imports this package and has no CPU fallback.

Parity status:
  * GGML block codecs (Q4_K / Q6_K / Q8_0 dequantise) are PINNED against the
    independent llama.cpp-derived implementation in the ``gguf`` Python package
    (``gguf.quants.dequantize``) through committed fixtures in ``tests/golden/``.
  * Slot-mapping / block-table arithmetic is pinned by known-answer cases restated
    from ``src/openai/pipelines/inputs.rs:12-22,410-430`` (reference).
  * Marlin scale permutations are pinned against the reference's own Python
    (``examples/convert_awq_marlin.py:8-17``) through committed fixtures.
  * PARITY UNPINNED for paged attention, reshape_and_cache, copy/swap_blocks,
    QMatMul activations path, FP8 KV: the arithmetic lives in the un-vendored
    dependencies attention-rs @a97f519 (v0.6.5) and candle-core fork @cafd231
    (v0.8.3); the reference tree holds no tests, fixtures or golden vectors for
    them.  Those oracles follow the in-tree call sites and ``NaiveAttention``
    (``src/openai/models/mod.rs:1268-1307``), the only in-tree statement of the
    attention semantics.
"""
