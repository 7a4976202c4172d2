#!/usr/bin/env python
"""bench.py -- headline benchmark: decode tokens/s, Llama-3-8B Q4_K (GGUF), batch 32, ctx 4096 -> 5120,
paged KV (bf16), on N B200s (tensor parallel = N), synthetic weights / KV / block tables.

One "step" = one decode step of the whole batch through the hot path (32 decoder layers of
quantised GEMMs + paged attention + lm_head + greedy argmax).  Prints ONE JSON line (see the
driver contract):  value = tokens/s with all inputs resident in HBM (CUDA-graph replay, metadata
advanced on the device); e2e = the same through the public host API (prepare_decode on the host,
H2D metadata copies, D2H token read-back inside the timed region); roofline = the paged-attention
decode kernel (HBM-bound) measured live with CUDA events; cpu_baseline = the C port of the
reference's CPU/GGML arithmetic (oracle/cpu_ref.c) on a bounded sample.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU; TP = N)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "decode tokens/s Llama-3-8B Q4_K batch=32"
UNIT = "tokens/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=4096, help="context length at the first decode step")
    ap.add_argument("--max-ctx", type=int, default=5120)
    ap.add_argument("--layers", type=int, default=32, help="debug only: fewer layers => number is INVALID")
    ap.add_argument("--kv", default=None, choices=["bf16", "fp8"])
    ap.add_argument("--config", default="q4k", choices=["q4k", "dense_bf16", "gptq_fp8kv"],
                    help="q4k = the metric (BASELINE config: Llama-3-8B Q4_K GGUF); dense_bf16 = BASELINE config 2 (Llama-3-8B BF16); "
                         "gptq_fp8kv = BASELINE config 3 (GPTQ/Marlin int4 + FP8 KV).  The extra configs are single-GPU.")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--parity-steps", type=int, default=3, help="decode steps compared against a TP=1 engine before timing (0 = off)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: C port of the reference's CPU/GGML path (oracle/cpu_ref.c) on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample_setup(batch: int, ctx: int):
    """Times ONE decoder layer (Q4_K x Q8_K integer-dot GEMMs + f32 paged attention over a bf16
    cache) for `batch` sequences at context `ctx`, plus a row slice of the Q6_K lm_head, and scales to
    the full model: t_step = 32 * t_layer + t_head * (vocab / rows).  Returns (tokens/s, info)."""
    import ctypes as C
    from oracle import cpu_ref
    from oracle import attention as OA
    lib = cpu_ref.lib()
    rng = np.random.default_rng(0)
    H, heads, kvh, hd, F, V, bs = 4096, 32, 8, 128, 14336, 128256, 64
    nblk = -(-(ctx + 1) // bs)

    def q4k(n, k):
        b = rng.integers(0, 256, (n * (k // 256), 144), dtype=np.uint8)
        b[:, 0:2] = (rng.uniform(0.5, 2, len(b)) * 2.0 ** -14).astype(np.float16).view(np.uint8).reshape(-1, 2)
        b[:, 2:4] = (rng.uniform(0.5, 2, len(b)) * 7.5 * 2.0 ** -14).astype(np.float16).view(np.uint8).reshape(-1, 2)
        return b.reshape(-1)

    ws = dict(wq=q4k(heads * hd, H), wk=q4k(kvh * hd, H), wv=q4k(kvh * hd, H), wo=q4k(H, heads * hd),
              w1=q4k(F, H), w2=q4k(H, F), w3=q4k(F, H))
    norms = dict(attn_norm=rng.uniform(0.5, 1.5, H).astype(np.float32), ffn_norm=rng.uniform(0.5, 1.5, H).astype(np.float32))
    nb = batch * nblk
    kc = rng.integers(0, 2 ** 16, (nb, bs, kvh, hd), dtype=np.uint16) & 0xBFFF     # finite bf16 bit patterns
    vc = rng.integers(0, 2 ** 16, (nb, bs, kvh, hd), dtype=np.uint16) & 0xBFFF
    kc = (kc & 0x807F) | 0x3F00; vc = (vc & 0x807F) | 0x3F00                         # |x| in [0.5, 1)
    bt = rng.permutation(nb).reshape(batch, nblk).astype(np.uint32)
    ctx_lens = np.full(batch, ctx + 1, np.uint32)
    pos = np.full(batch, ctx, np.int64)
    slots = (bt[:, ctx // bs].astype(np.int64) * bs + ctx % bs)
    cos, sin = OA.rope_tables(hd, ctx + 8, 500000.0)
    cfg = cpu_ref.RefCfg(H, heads, kvh, hd, F, bs, nblk, 1e-5)
    layer = cpu_ref.RefLayer(*[a.ctypes.data for a in (norms["attn_norm"], norms["ffn_norm"], ws["wq"], ws["wk"], ws["wv"],
                                                       ws["wo"], ws["w1"], ws["w2"], ws["w3"])])
    x = rng.standard_normal((batch, H)).astype(np.float32)
    lib.ref_layer_scratch_floats.restype = C.c_size_t
    scratch = np.empty(lib.ref_layer_scratch_floats(C.byref(cfg), batch), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)

    def one_layer():
        t0 = time.perf_counter()
        rc = lib.ref_llama_layer_decode(C.byref(cfg), C.byref(layer), p(x), batch, p(pos), p(slots), p(bt), p(ctx_lens),
                                        p(kc), p(vc), p(cos), p(sin), p(scratch))
        assert rc == 0
        return time.perf_counter() - t0

    rows = 8192
    w6 = rng.integers(0, 256, (rows * (H // 256), 210), dtype=np.uint8)
    w6[:, 208:210] = (rng.uniform(0.5, 2, len(w6)) * 2.0 ** -16).astype(np.float16).view(np.uint8).reshape(-1, 2)
    w6 = w6.reshape(-1)
    xh = rng.standard_normal((batch, H)).astype(np.float32)

    keep = (ws, norms, kc, vc, bt, ctx_lens, pos, slots, cos, sin, x, scratch, layer, cfg)   # raw pointers inside

    # ONE thread count, chosen by rule and printed (r01 re-calibrated per run and moved 4x between boxes): the physical cores this
    # process may use = half the schedulable logical CPUs (SMT pairs; on the GPU box 128 logical threads ran 10x slower than 64)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    n_thr = max(1, avail // 2) if avail >= 16 else max(1, avail)
    cpu_ref.set_num_threads(n_thr)
    one_layer()                                                    # warm the team and the page tables (not counted)

    def step():
        """one decode step of the whole model, nothing extrapolated: 32 decoder-layer passes (one layer's weights and KV reused -- 123 MB +
        1 GB per pass, far beyond the CPU caches) + the Q6_K lm_head over all `V` rows in slices of `rows`.  Returns seconds."""
        assert keep
        t_layers = sum(one_layer() for _ in range(32))
        t0 = time.perf_counter()
        done = 0
        while done < V:
            r = min(rows, V - done)
            cpu_ref.qmatmul_q8k(xh, w6[:r * (H // 256) * 210], 14, r, H)
            done += r
        t_head = time.perf_counter() - t0
        return t_layers + t_head, t_layers / 32, t_head

    info = dict(cores=cpu_ref.num_threads(), kind="port",
                sample=f"per step: all 32 decoder layers (batch {batch}, ctx {ctx}, Q4_K x Q8_K int dot + f32 paged attention; the "
                       f"weights and KV of one layer are reused for every pass) + the full {V}-row Q6_K lm_head -- a complete decode step, "
                       f"nothing scaled; {n_thr} threads = "
                       f"{'half the ' + str(avail) + ' schedulable logical CPUs' if avail >= 16 else 'all schedulable CPUs'}; mean over samples")
    return step, info


def cpu_decode_sample(batch: int, ctx: int, budget_s: float):
    step, info = cpu_sample_setup(batch, ctx)
    vals = []
    t_start = time.perf_counter()
    while len(vals) < 3 and (not vals or (time.perf_counter() - t_start) * (len(vals) + 1) / len(vals) < budget_s):
        vals.append(step()[0])
    mean = statistics.mean(vals)
    info = dict(info)
    info["sample"] += f"; {len(vals)} samples, mean {mean:.2f} s/step (min {min(vals):.2f}, max {max(vals):.2f})"
    return batch / mean, mean, info


def timed_window(args):
    """context lengths (including the decoded token) of the K timed steps: identical for both arms"""
    ctx0 = args.ctx + args.parity_steps
    return ctx0 + 1 + args.warmup + 1, ctx0 + 1 + args.warmup + args.steps


def workload_config(args, world: int, layers: int = 32) -> dict:
    """`config` of the JSON line -- the SAME dict from `--impl reference` and from the GPU arm (same workload, same window)."""
    first, last = timed_window(args)
    name = {"q4k": "Llama-3-8B Q4_K (lm_head Q6_K)", "dense_bf16": "Llama-3-8B BF16 (dense)",
            "gptq_fp8kv": "Llama-3-8B GPTQ int4 g128 (Marlin-prepared; lm_head Q6_K)"}[args.config]
    kv = args.kv or {"q4k": "bf16", "dense_bf16": "bf16", "gptq_fp8kv": "fp8"}[args.config]
    return {"workload": f"{name} decode, batch {args.batch}, ctx {first}->{last} of 4096->5120, block_size 64, {kv} paged KV, "
                        f"random non-contiguous block tables",
            "parallelism": f"tp{world}", "global_batch": args.batch, "layers": layers,
            "l2_policy": "inputs larger than L2 (KV 17+ GB and weights 4.4 GB streamed per step; 126 MB L2)"}


def run_reference(args):
    """--impl reference: the reference's CPU path (C port; the Rust reference cannot be built here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, info = cpu_sample_setup(args.batch, timed_window(args)[0])       # the first context of the timed window
    vals, t_begin = [], time.perf_counter()
    for i in range(args.warmup + args.steps):
        t_step, _, _ = step()
        if i >= args.warmup or (time.perf_counter() - t_begin) > 60:
            vals.append((args.batch / t_step, t_step))
        if vals and (time.perf_counter() - t_begin) > 170:       # keep the whole run within a few minutes (the line reports how many steps ran)
            break
    ms = statistics.mean(v[1] for v in vals) * 1e3                  # same statistic as the cpu_baseline leg: mean seconds per step
    tps = args.batch / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "q4_k weights x q8_k activations (int8 dot), f32 attention", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "note": "CPU arm (host cores only, rank 0): every step is a complete decode step of the same workload (one layer's weights / KV reused across the 32 layer passes)",
            "cpu_baseline": {"value": tps, "unit": UNIT, **info},
            "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
CONFIGS = {
    "q4k": dict(metric=METRIC, kv="bf16", dtype="f16 activations x q4_k/q6_k weights (fp32 accumulate), bf16 attention",
                name="Llama-3-8B Q4_K (lm_head Q6_K)"),
    "dense_bf16": dict(metric="decode tokens/s Llama-3-8B BF16 batch=32 (BASELINE config 2)", kv="bf16",
                       dtype="bf16 activations x bf16 weights (fp32 accumulate, tcgen05 dense GEMM), bf16 attention", name="Llama-3-8B BF16 (dense)"),
    "gptq_fp8kv": dict(metric="decode tokens/s Llama-3-8B GPTQ/Marlin int4 + FP8 KV batch=32 (BASELINE config 3)", kv="fp8",
                       dtype="f16 activations x int4 weights g128 (fp32 accumulate), f16 attention over e4m3 KV",
                       name="Llama-3-8B GPTQ int4 g128 (Marlin-prepared; lm_head Q6_K)"),
}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import candle_vllm_b200 as pkg
    from candle_vllm_b200 import synthetic
    from candle_vllm_b200.distributed import Comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = Comm(rank, world) if world > 1 else None

    conf = CONFIGS[args.config]
    if args.kv is None:
        args.kv = conf["kv"]
    if args.config != "q4k" and world > 1:
        raise SystemExit(f"bench.py --config {args.config} is a single-GPU configuration")
    B, bs = args.batch, 64
    K, W = args.steps, args.warmup
    PS = args.parity_steps
    args.max_ctx = max(args.max_ctx, args.ctx + 3 * (K + W) + PS + 16, 4608 + 2 * (K + W) + 16)      # room for every leg below
    blocks_per_seq = -(-args.max_ctx // bs)
    cfg = pkg.LlamaConfig(num_layers=args.layers, max_num_seqs=B, max_blocks_per_seq=blocks_per_seq, max_pos=args.max_ctx + 64,
                          block_size=bs)
    kv_dtype = pkg.DType.FP8_E4M3 if args.kv == "fp8" else pkg.DType.BF16
    num_blocks = B * blocks_per_seq + 16
    tables = synthetic.random_block_tables(B, blocks_per_seq, num_blocks, seed=2)
    tables_np = np.asarray(tables, np.int32)              # rectangular tables: prepare_decode's vectorised path
    stream = torch.cuda.Stream()

    def build(tp_rank, tp_world, nccl):
        """model + KV state for (tp_rank, tp_world): same seeds on every rank and for every tp_world, so a TP run and the TP = 1
        run see the same global weights and the same global KV cache (each rank holds its shard of both)."""
        if args.config == "dense_bf16":
            weights = synthetic.make_weights_16bit(cfg, dev, seed=0, dtype=torch.bfloat16)
        elif args.config == "gptq_fp8kv":
            weights, _ = synthetic.make_weights_gptq(cfg, dev, seed=0, group_size=128, dtype=torch.float16, with_oracle=False)
        else:
            weights = synthetic.make_weights(cfg, dev, seed=0, tp_rank=tp_rank, tp_world=tp_world)
        eng = pkg.CacheEngine(cfg.num_layers, cfg.num_kv_heads, cfg.head_dim,
                              pkg.CacheConfig(bs, num_blocks, kvcache_dtype="fp8" if args.kv == "fp8" else "auto"),
                              device=dev, num_shards=tp_world)
        synthetic.fill_kv_cache(eng.gpu_cache, seed=1, tp_rank=tp_rank, tp_world=tp_world, num_kv_heads=cfg.num_kv_heads)
        torch.cuda.synchronize()
        model = pkg.GGUFLLaMa(cfg, weights, eng.gpu_cache, kv_dtype=kv_dtype, tp_rank=tp_rank, tp_world=tp_world, nccl_comm=nccl,
                              stream=stream, rope_neox=args.config != "q4k")
        return model, eng, weights

    model, eng, weights = build(rank, world, comm.handle.value if comm else None)
    inboxes = None
    if world > 1:
        # row-parallel all-reduce + residual add + next RMSNorm as one kernel over NVLink peer memory (falls back to NCCL if
        # CUDA IPC is not available between the ranks, or with B200_TP_NCCL=1)
        from candle_vllm_b200.distributed import PeerInboxes
        inboxes = PeerInboxes(model, rank, world)
        if not inboxes.active and rank == 0:
            print(f"[bench] peer inboxes inactive, all-reduce through NCCL: {inboxes.error or 'B200_TP_NCCL set'}", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    toks0 = [int(t) for t in np.random.default_rng(3).integers(0, cfg.vocab, B)]

    # ---- (0) parity, OUTSIDE every timed region: the first PS decode steps of this run (graph replay, fused all-reduce and
    # all) against the same steps of a TP = 1 engine with the same seeds on rank 0: greedy tokens and full-vocabulary logits.
    # At N = 1 the second engine runs eagerly (no graph) -- that leg checks graph capture / replay and bounds the run-to-run
    # spread of the red.add reductions.
    def parity_steps(mdl, want, forced=None):
        """PS decode steps; with `forced` (the reference's tokens, [PS, B]) step s + 1 consumes forced[s] instead of the engine's own
        argmax -- teacher forcing, so that a greedy near-tie that flips under the run-to-run noise of the fp32 reductions cannot
        make the later steps decode different sequences and every step's logits stay comparable"""
        lens, toks, out_t, out_l = np.full(B, args.ctx + 1), list(toks0), [], []
        for st in range(PS):
            nxt, lg = mdl.decode(pkg.prepare_decode(lens, toks, tables_np, bs), want_logits=want)
            out_t.append(np.asarray(nxt).copy()); out_l.append(lg)
            toks, lens = [int(t) for t in (nxt if forced is None else forced[st])], lens + 1
        return np.stack(out_t), out_l

    parity = None
    if PS > 0:
        ref_t = torch.zeros(PS, B, dtype=torch.int64, device=dev)
        ref_l = None
        if rank == 0:
            if world == 1:
                ref_model = pkg.GGUFLLaMa(cfg, weights, eng.gpu_cache, kv_dtype=kv_dtype, use_graph=False, stream=stream,
                                          rope_neox=args.config != "q4k")
                ref_eng = None                                      # same cache: the steps rewrite the same slots with the same values
            else:
                ref_model, ref_eng, ref_w = build(0, 1, None)
            t, ref_l = parity_steps(ref_model, True)
            ref_t.copy_(torch.from_numpy(t.astype(np.int64)))
            del ref_model
            if world > 1:
                del ref_eng, ref_w
            torch.cuda.empty_cache()
        if world > 1:
            dist.broadcast(ref_t, src=0)
        ref_t = ref_t.cpu().numpy()
        got_t, got_l = parity_steps(model, True, forced=ref_t)      # collective at N > 1 (logits all-gather)
        barrier()
        if rank == 0:
            scale = max(float(np.abs(l).max()) for l in ref_l)
            tol = 3e-2 if args.config == "dense_bf16" else 5e-3
            errs = [float(np.abs(a - b).max()) / scale for a, b in zip(got_l, ref_l)]
            err = max(errs)
            # a greedy token may differ only on a near-tie inside the logit error
            flips = mismatches = 0
            for st in range(PS):
                for b in np.nonzero(got_t[st] != ref_t[st])[0]:
                    mismatches += 1
                    margin = float(ref_l[st][b].max() - ref_l[st][b, got_t[st][b]])
                    flips += int(margin > 2 * err * scale + 1e-6)
            parity = {"steps": PS, "logits_max_err": err, "logits_max_err_per_step": errs,
                      "tokens_equal_tp1": mismatches == 0, "token_mismatches": mismatches,
                      "token_mismatches_beyond_logit_error": flips, "logit_scale": scale,
                      "reference": "TP=1 engine, same seeds, rank 0" + (" (eager, no CUDA graph)" if world == 1 else "")
                                   + "; teacher-forced: every step of both engines consumes the reference's tokens",
                      "tolerance": tol, "within_tolerance": bool(err < tol and flips == 0),
                      "tolerance_note": ("bf16 activations (relative step 3.9e-3) amplify the fp32 reduction-order difference between the two "
                                         "engines layer by layer" if args.config == "dense_bf16" else
                                         "fp16 activations; fp32 reduction order differs between the two engines "
                                         "(profiles/r02_rounding_floor.md)") + "; unit: fraction of max|logit|"}
            del ref_l
        barrier()

    # ---- (1) device-resident: metadata advanced on the device, graph replay only ----------------
    def resident_leg(ctx0):
        prep = pkg.prepare_decode([ctx0 + 1] * B, toks0, tables, bs)   # context INCLUDING the token being decoded
        model.decode(prep)                                     # loads the static buffers (captures the graph on first use)
        for _ in range(W):
            model.decode_resident(B, advance=True)
        barrier()
        sampler = ClockSampler(local); sampler.start()
        l0 = model.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(K):
                model.decode_resident(B, advance=True)
            e1.record(stream)
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        return ms_total / K, sampler.stop(), model.kernel_launches() - l0, (ctx0 + 1 + W + 1, ctx0 + 1 + W + K)

    ms_step, clocks, launches, (ctx_first, ctx_last) = resident_leg(args.ctx + PS)
    assert (ctx_first, ctx_last) == timed_window(args)
    value = B / (ms_step * 1e-3)
    # the metric is quoted over ctx 4096 -> 5120 (mean 4608): the same leg centred on the mean context
    mid0 = max(args.ctx, 4608 - (W + K) // 2 - 1)
    ms_mid, clocks_mid, _, (mid_first, mid_last) = resident_leg(mid0)

    # ---- (2) end to end through the host API -------------------------------------------------------
    # over the SAME context window as `value` (ctx_first -> ctx_last), so that the two numbers differ by the host path only
    n_warm = min(W, 3)
    cur = ctx_first - n_warm
    nxt = model.read_next_tokens(B)
    h2d = B * (8 + 8 + 8 + 4) + B * blocks_per_seq * 4
    d2h = B * 4
    for _ in range(n_warm):
        prep = pkg.prepare_decode(np.full(B, cur), nxt, tables_np, bs)
        nxt, _ = model.decode(prep); cur += 1
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K):
        prep = pkg.prepare_decode(np.full(B, cur), nxt, tables_np, bs)              # host: block tables, slots
        nxt, _ = model.decode(prep); cur += 1                                       # H2D + replay + D2H + sync
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms)) / K
    e2e = B / (e2e_ms * 1e-3)

    # ---- (3) rooflines: paged-attention decode alone per layer; the weight stream of all projections alone ---------
    roof = attention_roofline(pkg, model, cfg, eng, B, cur, tables, world, stream, dev)
    roof_gemm = gemm_roofline(model, cfg, B, world, stream, args.config)
    peer_ar = inboxes is not None and inboxes.active
    if inboxes is not None:
        timed_out = inboxes.timed_out()
        inboxes.close()                            # collective (barrier): before any rank leaves
        if timed_out:
            raise SystemExit("bench: the fused all-reduce timed out on a peer -- numbers invalid")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6.65 TB/s"
    roof.update(peak=hbm_peak, peak_source=src, frac=roof["achieved"] / hbm_peak)
    roof_gemm.update(peak=hbm_peak, peak_source=src, frac=roof_gemm["achieved"] / hbm_peak)
    line = {"metric": conf["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": conf["dtype"], "data": "synthetic",
            "config": workload_config(args, world, cfg.num_layers),
            "all_reduce": None if world == 1 else ("fused all-reduce + add + norm over NVLink peer memory" if peer_ar else "NCCL all-reduce"),
            "value_mean_ctx": {"value": B / (ms_mid * 1e-3), "unit": UNIT, "ms_per_step": ms_mid, "ctx": f"{mid_first}->{mid_last}",
                               "note": "same leg centred on the metric's mean context 4608 (attention bytes +12 % over ctx 4096)",
                               "clocks": clocks_mid},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ctx": f"{cur - K}->{cur - 1}",
                    "api": "prepare_decode (host) + GGUFLLaMa.decode -> b200_llama_decode (C ABI, host buffers)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_gemm": roof_gemm}
    if parity is not None:
        line["parity"] = parity
    if cfg.num_layers != 32:
        line["invalid"] = "debug run with fewer layers"
    if not args.no_cpu_baseline and world == 1 and args.config == "q4k":      # the CPU arm is the reference's GGUF/GGML path
        tps, t_step, info = cpu_decode_sample(B, timed_window(args)[0], args.cpu_seconds)
        line["cpu_baseline"] = {"value": tps, "unit": UNIT, **info}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(csv_name: str, kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first launch of `kernel_substr` in a TRACKED `ncu --page raw --csv`
    export under profiles/ (None when the file or the metrics are missing): roofline.traffic is computed, never hard-coded."""
    import csv
    path = os.path.join(ROOT, "profiles", csv_name)
    try:
        rows = list(csv.reader(open(path, newline="")))
    except OSError:
        return None, None
    hdr = next((r for r in rows if "Kernel Name" in r), None)
    if hdr is None:
        return None, None
    ik = hdr.index("Kernel Name")
    units = rows[rows.index(hdr) + 1]
    def col(name):
        return hdr.index(name) if name in hdr else None
    ir, iw = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    if ir is None or iw is None:
        return None, None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[rows.index(hdr) + 2:]:
        if len(r) > max(ik, ir, iw) and kernel_substr in r[ik]:
            try:
                rd = float(r[ir].replace(",", "")) * mult.get(units[ir], 1.0)
                wr = float(r[iw].replace(",", "")) * mult.get(units[iw], 1.0)
                return rd + wr, os.path.join("profiles", csv_name)
            except ValueError:
                return None, None
    return None, None


def gemm_roofline(model, cfg, B, world, stream, config="q4k"):
    """The weight stream alone: every quantised projection of every layer + the lm_head + the small ops between them, without
    RoPE / cache write / attention (b200_llama_linear_chain), timed eagerly with CUDA events on the launching stream.  4.36 GB
    of weights per pass at TP = 1: far beyond the 126 MB L2."""
    import torch
    H, F, V = cfg.hidden, cfg.ffn, cfg.vocab
    qd, kd = cfg.num_heads * cfg.head_dim, cfg.num_kv_heads * cfg.head_dim
    params = (H * (qd + 2 * kd) + qd * H + 3 * H * F) * cfg.num_layers
    if config == "dense_bf16":
        alg = (params + V * H) * 2
    elif config == "gptq_fp8kv":
        alg = params // 2 + params // 128 * 2 + V * H * 210 // 256        # int4 + one f16 scale per 128 weights; lm_head Q6_K
    else:
        alg = (params * 144 // 256 + V * H * 210 // 256) // world
    with torch.cuda.stream(stream):
        for _ in range(3):
            model.linear_chain(B)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        stream.synchronize()
        e0.record(stream)
        for _ in range(reps):
            model.linear_chain(B)
        e1.record(stream)
        stream.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"kernel": "all quantised projections of one decode step (QKV, wo, gate|up, w2 per layer + lm_head) incl. norm / SiLU between them, eager launches",
            "bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "unit": "GB/s", "algorithmic_bytes_per_launch": alg,
            "ms_per_launch": ms, "launch": f"one pass over all layers (weights {alg / 1e9:.2f} GB on this rank)",
            "traffic": None}


def attention_roofline(pkg, model, cfg, eng, B, ctx, tables, world, stream, dev):
    """Times PagedAttention decode alone (all layers round-robin, so every launch streams a different
    layer's KV: 0.5+ GB per launch >> 126 MB L2) with CUDA events on the launching stream."""
    import torch
    heads_l = cfg.num_heads // world
    kv_l = max(cfg.num_kv_heads // world, 1)
    attn = pkg.PagedAttention(heads_l, cfg.head_dim, cfg.head_dim ** -0.5, kv_l, fp8_kvcache=eng.dtype == torch.uint8)
    prep = pkg.prepare_decode([ctx] * B, [0] * B, tables, cfg.block_size)
    bt = np.zeros((B, cfg.max_blocks_per_seq), np.int32); bt[:, :prep["block_tables"].shape[1]] = prep["block_tables"]
    with torch.cuda.stream(stream):
        meta = pkg.InputMetadata(False, torch.zeros(0, dtype=torch.int64, device=dev), torch.from_numpy(bt).to(dev),
                                 torch.from_numpy(prep["context_lens"]).to(dev))
        q = torch.randn((B, heads_l, cfg.head_dim), device=dev, dtype=torch.float32).to(torch.bfloat16)
        for l in range(min(3, cfg.num_layers)):
            attn.forward(q, None, None, None, eng.gpu_cache[l][0], eng.gpu_cache[l][1], meta, out_dtype=torch.float16)
        reps = max(cfg.num_layers, 32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def run_all():
            for i in range(reps):
                k, v = eng.gpu_cache[i % cfg.num_layers]
                attn.forward(q, None, None, None, k, v, meta, out_dtype=torch.float16)

        # the launches are captured once and replayed: at 50-100 us per launch an eager Python loop would measure the interpreter
        graph, how = None, "CUDA-graph replay of the launches"
        try:
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                run_all()
            graph.replay()                                   # warm replay
        except Exception as exc:                             # capture unavailable: eager launches (an upper bound on the kernel time)
            graph, how = None, f"eager launches ({type(exc).__name__})"
        stream.synchronize()
        e0.record(stream)
        if graph is not None:
            graph.replay()
        else:
            run_all()
        e1.record(stream)
        stream.synchronize()
    ms = e0.elapsed_time(e1) / reps
    esz = eng.gpu_cache[0][0].element_size()
    kv_bytes = B * ctx * 2 * kv_l * cfg.head_dim * esz
    io_bytes = 2 * B * heads_l * cfg.head_dim * 2
    alg = kv_bytes + io_bytes
    return {"kernel": "paged_attention_decode (one layer: split-KV kernel + merge)", "bound": "hbm",
            "achieved": alg / (ms * 1e-3) / 1e9, "unit": "GB/s", "algorithmic_bytes_per_launch": alg,
            "ms_per_launch": ms, "ctx": ctx, "timed_as": how,
            # dram__bytes_read + dram__bytes_write of ONE `ncu --set full` capture of this kernel inside this benchmark, parsed
            # from the tracked raw export (profiles/); only quoted for the configuration it was captured on
            **traffic_fields(world, B, esz)}


def traffic_fields(world, B, esz):
    if not (world == 1 and B == 32):
        return {"traffic": None}
    if esz == 1:                                  # FP8 KV
        t, src = ncu_traffic("r02_attention_decode_fp8_raw.csv", "paged_attn_decode_kernel")
    else:
        t, src = ncu_traffic("r02_attention_decode_raw.csv", "paged_attn_decode_kernel")
        if not t:
            t, src = ncu_traffic("r01_attention_decode_raw.csv", "paged_attn_decode_kernel")
    return {"traffic": t, "traffic_source": src} if t else {"traffic": None, "traffic_source": "no tracked ncu export found"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
