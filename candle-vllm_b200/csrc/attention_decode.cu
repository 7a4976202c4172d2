// attention_decode.cu -- paged-attention decode for B200: TMA-staged KV pages, split-KV persistent
// CTAs, tensor-core QK^T / PV with warp-level online softmax, plus the split merge.
//
// Why this shape (DESIGN.md "paged attention decode"): the op is HBM-bound -- per layer it streams
// B * ctx * 2 * kvh * 128 * 2 B of KV (537-671 MB at B=32, ctx 4-5k) against ~4 flop/B of math.
// So the kernel is organised around keeping >= 190 KB of KV reads in flight per SM:
//   * KV cache = 4-D tensor [num_blocks, 64, kvh, 128] (flash layout,
//     /root/reference/src/scheduler/cache_engine.rs:326-340); one TMA box {64 dims, 1 head, 32 tokens,
//     1 block} lands a quarter page (4 KB) in shared memory with the 128-byte swizzle, 4 boxes per tile
//     (K lo/hi, V lo/hi), completion on an mbarrier (cp.async.bulk.tensor -> UTMALDG).
//   * work item = (sequence, kv head, chunk of <= 8 pages); items are enumerated from the DEVICE-side
//     context_lens (graph-replay safe, graph.rs:604) with the kv head fastest so concurrently running
//     warps touch the same DRAM pages; a persistent grid of one CTA per SM pulls items from a global
//     atomic queue (no tail imbalance).
//   * every warp (6 per CTA) is an independent pipeline: it gathers the chunk's block ids from the block
//     table, issues its own TMA loads two 32-token tiles ahead into a private 2-stage ring (16 KB per
//     stage, 192 KB in flight per SM) and consumes them: S = Q K^T with mma.sync m16n8k16 (the GQA
//     group's <= 8 query heads are the M rows, K read with ldmatrix from the swizzled tile), online
//     softmax in registers (exp2, fp32), O += P V with ldmatrix.trans on V.  Tensor cores are needed
//     only to keep the issue slots free: the arithmetic (16 flop/B padded) is far below the tensor peak.
//     A warp only waits on mbarrier phases it armed itself, so there is no cross-warp phase aliasing
//     and no CTA barrier in the steady state.
//   * per item the warp writes an fp32 partial (m, l, O) straight from registers; a tiny merge kernel
//     folds the chunks of each (sequence, head), rounds to bf16 and re-zeroes the queue head.
//
// Semantics: softmax(q k^T * scale) v over the block table, GQA by head grouping
// (NaiveAttention::forward /root/reference/src/openai/models/mod.rs:1268-1307; call sites
// layers/attention.rs:707-718, :983-994; metadata pipelines/inputs.rs:552-568).
#include <cuda.h>

#include <cstdlib>

#include <type_traits>

#include "attention.cuh"

namespace b200 {

namespace {

constexpr int kHeadDim = 128;
constexpr int kPage = 64;                     // tokens per KV block (main.rs:364-366 default)
constexpr int kTile = 32;                     // tokens per pipeline stage (half a page)
constexpr int kMaxChunkPages = 8;
constexpr int kSubTileBytes = kTile * 64 * 2;             // 32 tokens x 64 dims x 2 B = 4 KB
constexpr int kMaxSeqs = 1024;

// Shared-memory plan.  16-bit KV: 6 independent warp pipelines x 2 stages of 16 KB (K lo, K hi, V lo, V hi sub-tiles) = 192 KB of KV
// reads in flight per SM.  FP8 (e4m3) KV: a tile is half the bytes (one 4 KB box for K, one for V: 32 tokens x 128 B), so 8 warps x 3
// stages of 8 KB keep the same 192 KB in flight.  The e4m3 bytes are expanded to f16 IN REGISTERS on their way into the MMA fragments
// (K: LDS.128 of a token row; V: the transposing 8-bit ldmatrix of sm_100a) -- there is no f16 staging tile.  [A first version
// expanded each tile into a per-warp f16 staging tile (LDS + cvt + STS, then the 16-bit ldmatrix code): 0.45-0.59 of the HBM peak,
// bound by the shared-memory round trip and the 24 KB per warp it cost; profiles/r02_fp8_attention.md.]
#ifndef B200_FP8_WARPS          // tuning aids (tools/gpu_run_r02m.sh builds variants)
#define B200_FP8_WARPS 8
#endif
#ifndef B200_FP8_STAGES
#define B200_FP8_STAGES 3
#endif
template <bool kFp8>
struct Plan {
    static constexpr int kWarps = kFp8 ? B200_FP8_WARPS : 6;
    static constexpr int kStages = kFp8 ? B200_FP8_STAGES : 2;
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kStageBytes = kFp8 ? 2 * kTile * kHeadDim : 4 * kSubTileBytes;      // 8 KB / 16 KB
    static constexpr int kWarpBytes = kStages * kStageBytes;
    // stages first (1024-byte aligned for the 128B swizzle): [warp][stage...]
    static constexpr int kBars = kWarps * kWarpBytes;                               // full[kWarps][kStages]
    static constexpr int kHdr = kBars + kWarps * kStages * 8;                       // int[kWarps][8][8] item headers (the producer cursor runs <= kStages + 2 items ahead)
    static constexpr int kPrefix = kHdr + kWarps * 8 * 8 * 4;                       // int[kMaxSeqs + 1]
    static constexpr int kCtx = kPrefix + (kMaxSeqs + 1) * 4;                       // int[kMaxSeqs]: context lengths (read once from global)
    static constexpr int kMail = (kCtx + kMaxSeqs * 4 + 7) / 8 * 8;                             // CTA ticket mailbox: u64[8] tickets + u32[8] acknowledgements
    static constexpr int kHalf = kMail + 64 + 32 + 16;                              // int[kMaxSeqs + 1]: chunk prefix for the halved chunk size (+ the decision word before it)
    static constexpr int kTotal = kHalf + (kMaxSeqs + 1) * 4;
    static_assert(kTotal <= 232448, "exceeds the 227 KB shared memory of an SM");
};

// ---- PTX helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct DecodeParams {
    const void* q;                    // [B, H, 128] T
    const uint32_t* block_tables;     // [B, max_blocks]
    const uint32_t* context_lens;     // [B]
    float* part_o;                    // [B*kvh*max_chunks][group][128]
    float* part_ml;                   // [B*kvh*max_chunks][group][2]  (m in log2 domain, l)
    unsigned int* counter;            // dynamic work queue head (zero on entry; the merge kernel re-zeroes it)
    int num_seqs, num_heads, num_kv_heads, max_blocks, chunk_pages, max_chunks;
    int adaptive_chunks;           // 1: the kernel may halve chunk_pages from the device-side context lengths (B200_ATTN_ADAPTIVE=0: off)
    int static_walk;               // queue mode: 0 per-warp tickets, 1 CTA tickets (FP8 cache), 2 static walk (B200_ATTN_STATIC overrides)
    float scale_log2;                 // scale * log2(e)
};

// One work item = (sequence b, kv head h, chunk c of <= chunk_pages pages), owned by ONE warp.
// Header words in shared memory: {valid, b, h, c, ctx, ntiles}.
//
// Every warp is an independent pipeline: it claims items from a global queue, gathers the chunk's block
// ids from the block table, issues its own TMA loads two (FP8: three) tiles ahead into a private ring
// (completion on an mbarrier) and consumes them with ldmatrix + mma.sync.  No CTA-wide barrier in the
// steady state, and a warp only ever waits on a barrier phase it armed itself (no phase aliasing).
// T = model dtype (q, and the cache when kFp8 is false).  kFp8: the cache holds e4m3 bytes (scale 1.0, the reference passes none);
// K / V are expanded to f16 (exact), q is converted bf16 -> f16 and the MMAs run in f16.

// four e4m3 bytes -> two packed f16x2 (bytes 0,1 | bytes 2,3); the 16-bit halves of w are register sub-words, no byte permutes
__device__ __forceinline__ void cvt_e4m3x4(uint32_t w, uint32_t& lo, uint32_t& hi) {
    asm("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %2;\ncvt.rn.f16x2.e4m3x2 %0, l;\ncvt.rn.f16x2.e4m3x2 %1, h;\n}\n" : "=r"(lo), "=r"(hi) : "r"(w));
}
__device__ __forceinline__ void ldmatrix_x2_trans_b8(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <typename T, int kGroup, bool kFp8>
__global__ void __launch_bounds__(Plan<kFp8>::kThreads, 1)
paged_attn_decode_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                         const DecodeParams p) {
    using PL = Plan<kFp8>;
    using TM = typename std::conditional<kFp8, __half, T>::type;          // MMA operand type
    constexpr int kStagesPerWarp = PL::kStages, kStageBytes = PL::kStageBytes, kThreads = PL::kThreads;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    int* prefix = reinterpret_cast<int*>(smem + PL::kPrefix);
    int* ctxs = reinterpret_cast<int*>(smem + PL::kCtx);
    int* hdr = reinterpret_cast<int*>(smem + PL::kHdr) + warp * 64;          // [8][8]
    const uint32_t my_stages = smem_base + warp * PL::kWarpBytes;
    const uint32_t my_bars = smem_base + PL::kBars + warp * kStagesPerWarp * 8;

    if (lane == 0) {
        for (int s = 0; s < kStagesPerWarp; ++s) mbar_init(my_bars + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 8) {
        reinterpret_cast<unsigned long long*>(smem + PL::kMail)[threadIdx.x] = 0ull;
        reinterpret_cast<unsigned int*>(smem + PL::kMail + 64)[threadIdx.x] = 0u;
    }
    pdl_wait();            // context_lens / q / KV written by the previous kernels
    pdl_trigger();
    // Chunk size from the DEVICE-side context lengths: the host proposes p.chunk_pages (pick_chunk_pages, from the table width);
    // items are whole chunks, so the makespan is rounds x chunk length -- when the proposed size leaves an awkward last round (e.g. 2.16
    // items per warp = 3 rounds of 16 tiles at ctx 4664 with FP8 KV: 74 us instead of 63), half the size wins.  Every CTA takes the same
    // decision from the same numbers; CTA 0 publishes it in the workspace header for the merge kernel.
    int* prefix_half = reinterpret_cast<int*>(smem + PL::kHalf);           // chunk prefix for the halved chunk size
    int* decision = reinterpret_cast<int*>(smem + PL::kMail + 96);
    {
        const int ct0 = p.chunk_pages * kPage, ct1 = (p.chunk_pages > 1 ? p.chunk_pages / 2 : 1) * kPage;
        for (int b = threadIdx.x; b < p.num_seqs; b += kThreads) {
            const int ctx = (int)p.context_lens[b];
            ctxs[b] = ctx;
            prefix[b + 1] = (ctx + ct0 - 1) / ct0; prefix_half[b + 1] = (ctx + ct1 - 1) / ct1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        prefix[0] = prefix_half[0] = 0;
        for (int b = 0; b < p.num_seqs; ++b) { prefix[b + 1] += prefix[b]; prefix_half[b + 1] += prefix_half[b]; }
        int cp = p.chunk_pages;
        if (p.chunk_pages > 1 && p.adaptive_chunks) {
            const int64_t slots = (int64_t)gridDim.x * PL::kWarps;
            const int64_t r0 = ((int64_t)prefix[p.num_seqs] * p.num_kv_heads + slots - 1) / slots;
            const int64_t r1 = ((int64_t)prefix_half[p.num_seqs] * p.num_kv_heads + slots - 1) / slots;
            // cost in tiles: rounds x (tiles per chunk + ~1 tile of per-item overhead: claim, Q fragments, partial write, merge)
            if (r1 * (p.chunk_pages + 1) < r0 * (2 * p.chunk_pages + 1)) cp = p.chunk_pages / 2;
        }
        decision[0] = cp;
        if (blockIdx.x == 0) p.counter[4] = (unsigned int)cp;
    }
    __syncthreads();
    const int chunk_pages = decision[0];
    const int chunk_tokens = chunk_pages * kPage;
    if (chunk_pages != p.chunk_pages) prefix = prefix_half;
    const int total_items = prefix[p.num_seqs] * p.num_kv_heads;

    // Claiming the next item of the global queue is split in two so that neither the atomic's round trip (~1 us) nor the block-table
    // load sits in the consumer's instruction stream: claim_begin() issues the atomicAdd, claim_finish() -- one tile later, or at
    // the latest when the item is needed -- turns the ticket into a header slot and issues the block-table load, whose result
    // (the lane-distributed block ids) is first touched when the item's first tile is issued.
    int n_claimed = 0;
    unsigned int pend_id = 0;
    bool pending = false;
    // Queue modes.  0: every warp draws its own ticket (one item).  1 (FP8 cache): the CTA draws tickets -- warp 0 does the atomic and
    // posts the ticket in a shared-memory mailbox, ticket T covers items T * kWarps + warp, so the warps of a CTA stream neighbouring
    // kv heads of the same (sequence, chunk) side by side: their 128-byte rows of one token are contiguous in the cache and, requested
    // together, open a DRAM page once (+7 % at 8 kv heads); balancing stays dynamic at CTA granularity.  2: static walk (ticket k of
    // a CTA = blockIdx + k * grid; uniform batches only; kept for comparison).
    // Mailbox protocol: slot j % 8 holds {j + 1, T}; a reader acknowledges, and warp 0 re-uses a slot only after all kWarps - 1
    // readers of its previous ticket have acknowledged.  Warp 0 never waits for a ticket, readers only for warp 0's own progress.
    const int qmode = p.static_walk == 2 ? (prefix[p.num_seqs] == p.num_seqs * prefix[1] ? 2 : 0) : p.static_walk;
    unsigned long long* mail = reinterpret_cast<unsigned long long*>(smem + PL::kMail);
    unsigned int* acks = reinterpret_cast<unsigned int*>(smem + PL::kMail + 64);
    unsigned int n_begun = 0;
    auto claim_begin = [&]() {
        // the first ticket of every warp / CTA is its own index: no atomic round trip (~1 us) before the first TMA of the launch
        if (qmode == 2) pend_id = (blockIdx.x + n_begun * gridDim.x) * PL::kWarps + warp;
        else if (n_begun == 0) pend_id = qmode == 1 ? blockIdx.x : blockIdx.x * PL::kWarps + warp;
        else if (qmode == 1) {
            if (warp == 0 && lane == 0) {
                const unsigned int slot = n_begun & 7;
                if (n_begun >= 9) {                      // ticket 0 never went through the mailbox
                    while (*reinterpret_cast<volatile unsigned int*>(&acks[slot]) != (unsigned int)(PL::kWarps - 1)) {}
                    *reinterpret_cast<volatile unsigned int*>(&acks[slot]) = 0u;
                    __threadfence_block();
                }
                pend_id = gridDim.x + atomicAdd(p.counter, 1u);
            }
        } else if (lane == 0) pend_id = gridDim.x * PL::kWarps + atomicAdd(p.counter, 1u);
        ++n_begun;
        pending = true;
    };
    auto claim_finish = [&](bool& valid, int& h, int& ntiles) -> uint32_t {
        unsigned int id = __shfl_sync(0xffffffffu, pend_id, 0);
        if (qmode == 1 && n_begun == 1) id = id * PL::kWarps + warp;          // ticket 0 = blockIdx, known to every warp
        else if (qmode == 1) {
            const unsigned int j = n_begun - 1, slot = j & 7;
            unsigned int T = id;
            if (lane == 0) {
                if (warp == 0) {
                    *reinterpret_cast<volatile unsigned long long*>(&mail[slot]) = ((unsigned long long)(j + 1) << 32) | T;
                } else {
                    unsigned long long v;
                    do { v = *reinterpret_cast<volatile unsigned long long*>(&mail[slot]); } while ((unsigned int)(v >> 32) != j + 1);
                    T = (unsigned int)v;
                    atomicAdd(&acks[slot], 1u);
                }
            }
            T = __shfl_sync(0xffffffffu, T, 0);
            id = T * PL::kWarps + warp;
        }
        pending = false;
        int* hd = hdr + (n_claimed & 7) * 8;
        ++n_claimed;
        valid = id < (unsigned int)total_items;
        h = 0; ntiles = 0;
        if (!valid) {
            if (lane == 0) hd[0] = 0;
            return 0u;
        }
        const int pair = (int)id / p.num_kv_heads;
        h = (int)id - pair * p.num_kv_heads;              // kv head fastest: neighbours share DRAM pages
        int lo = 0, hi = p.num_seqs;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= pair) lo = mid; else hi = mid; }
        const int c = pair - prefix[lo];
        const int ctx = ctxs[lo];
        const int ntok = min(chunk_tokens, ctx - c * chunk_tokens);
        ntiles = (ntok + kTile - 1) / kTile;
        const int npages = (ntok + kPage - 1) / kPage;
        if (lane == 0) { hd[0] = 1; hd[1] = lo; hd[2] = h; hd[3] = c; hd[4] = ctx; hd[5] = ntiles; }
        return lane < npages ? p.block_tables[(int64_t)lo * p.max_blocks + c * chunk_pages + lane] : 0u;
    };

    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    // ---- producer cursor: runs kStagesPerWarp tiles ahead of the consumer ---------------------------
    bool pc_valid, pn_valid;
    int pc_h, pc_ntiles, pn_h, pn_ntiles, p_tile = 0;
    claim_begin();
    uint32_t pc_blk = claim_finish(pc_valid, pc_h, pc_ntiles);
    uint32_t pn_blk = 0u;
    pn_valid = false; pn_h = 0; pn_ntiles = 0;
    claim_begin();
    unsigned int issued = 0;
    auto issue_one = [&]() {
        if (pending) pn_blk = claim_finish(pn_valid, pn_h, pn_ntiles);
        if (!pc_valid) return;
        const int blk = (int)__shfl_sync(0xffffffffu, pc_blk, p_tile >> 1);
        if (lane == 0) {
            const int s = issued % kStagesPerWarp;
            const uint32_t bar = my_bars + s * 8, dst = my_stages + s * kStageBytes;
            const int tok = (p_tile & 1) * kTile;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // our generic accesses to this stage are done
            mbar_expect_tx(bar, kStageBytes);
            if constexpr (kFp8) {               // one {128 B, 32 tokens} box each for K and V
                tma_load_4d(dst, &kmap, bar, 0, pc_h, tok, blk, policy);
                tma_load_4d(dst + kTile * kHeadDim, &vmap, bar, 0, pc_h, tok, blk, policy);
            } else {
                tma_load_4d(dst, &kmap, bar, 0, pc_h, tok, blk, policy);
                tma_load_4d(dst + kSubTileBytes, &kmap, bar, 64, pc_h, tok, blk, policy);
                tma_load_4d(dst + 2 * kSubTileBytes, &vmap, bar, 0, pc_h, tok, blk, policy);
                tma_load_4d(dst + 3 * kSubTileBytes, &vmap, bar, 64, pc_h, tok, blk, policy);
            }
        }
        ++issued;
        if (++p_tile == pc_ntiles) {
            pc_valid = pn_valid; pc_h = pn_h; pc_ntiles = pn_ntiles; pc_blk = pn_blk; p_tile = 0;
            if (pc_valid) claim_begin(); else pn_valid = false;           // an empty queue stays empty: no ticket past the end needed
        }
    };
#pragma unroll
    for (int s = 0; s < kStagesPerWarp; ++s) issue_one();

    // ---- consumer -----------------------------------------------------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const T* qbase = static_cast<const T*>(p.q);
    unsigned int consumed = 0;
    for (int c_slot = 0;; ++c_slot) {
        __syncwarp();
        const int* hd = hdr + (c_slot & 7) * 8;
        if (hd[0] == 0) break;
        const int b = hd[1], h = hd[2], c = hd[3], ctx = hd[4], ntiles = hd[5];

        // Q fragments: rows = the group's query heads (g < kGroup), zero padding otherwise.  16-bit cache: k-step ks covers dims
        // 16 ks .. 16 ks + 15 in the fragment's natural order.  FP8 cache: the K operand comes straight from 16-byte chunks of the
        // e4m3 rows (see below), so k-step ks pairs slots (2t, 2t+1 | 2t+8, 2t+9) with dims D .. D+3, D = 16 (t + 4 (ks / 4)) + 4 (ks % 4)
        // -- any permutation of the contraction index is fine as long as both operands use it
        uint32_t qa[8][2];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            if (g < kGroup) {
                const int d0 = kFp8 ? 16 * (t + 4 * (ks >> 2)) + 4 * (ks & 3) : ks * 16 + 2 * t, d1 = kFp8 ? d0 + 2 : d0 + 8;
                const T* qr = qbase + ((int64_t)b * p.num_heads + h * kGroup + g) * kHeadDim;
                if constexpr (std::is_same<T, TM>::value) {
                    qa[ks][0] = *reinterpret_cast<const uint32_t*>(qr + d0);
                    qa[ks][1] = *reinterpret_cast<const uint32_t*>(qr + d1);
                } else {                        // model dtype bf16, f16 MMAs (FP8 cache): q -> f16 (saturating; exact for |q| in fp16's normal range)
                    qa[ks][0] = pack2<TM>(to_f32(qr[d0]), to_f32(qr[d0 + 1]));
                    qa[ks][1] = pack2<TM>(to_f32(qr[d1]), to_f32(qr[d1 + 1]));
                }
            } else { qa[ks][0] = qa[ks][1] = 0u; }
        }
        float o[16][4];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;         // log2-domain running max; this thread's partial row sum

        for (int tl = 0; tl < ntiles; ++tl) {
            const int s = consumed % kStagesPerWarp;
            mbar_wait(my_bars + s * 8, (consumed / kStagesPerWarp) & 1);
            const uint32_t kt = my_stages + s * kStageBytes, vt = kt + (kFp8 ? kTile * kHeadDim : 2 * kSubTileBytes);
            const int valid = min(kTile, ctx - (c * chunk_tokens + tl * kTile));

            // ---- S = Q K^T : 4 n-tiles (8 tokens each) x 8 k-steps --------------------------------
            float sacc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
            if constexpr (kFp8) {
                // FP8 cache: no staging tile.  Lane (g, t) reads the two 16-byte chunks t and t + 4 of ONE e4m3 token row per n-tile
                // (32 dims), expands them in registers (cvt.rn.f16x2.e4m3x2: two weights per instruction, exact) and feeds the words
                // to the MMA as its B fragment.  Column n = g of n-tile nt is token 8 nt + (g >> 1) + 4 (g & 1): the two rows a
                // quarter-warp touches then differ in bit 2 of the swizzle key and the LDS.128 are conflict free; thread t ends up with
                // the scores of tokens 8 nt + t and 8 nt + t + 4 -- the order the transposing 8-bit ldmatrix of V hands out below.
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int row = nt * 8 + (g >> 1) + 4 * (g & 1);
                    const uint8_t* kr = smem + (kt - smem_base) + row * 128;
                    const uint4 c0 = *reinterpret_cast<const uint4*>(kr + ((t ^ (row & 7)) << 4));
                    const uint4 c1 = *reinterpret_cast<const uint4*>(kr + (((t + 4) ^ (row & 7)) << 4));
                    const uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        uint32_t b0, b1;
                        cvt_e4m3x4(w[ks], b0, b1);
                        const uint32_t a0[4] = {qa[ks][0], 0u, qa[ks][1], 0u};
                        mma_16816<TM>(sacc[nt], a0, b0, b1);
                    }
                }
            } else {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int row = nt * 8 + (lane & 7);
#pragma unroll
                    for (int kp = 0; kp < 4; ++kp) {            // 32 dims (2 k-steps) per ldmatrix.x4
                        const int chunk = kp * 4 + (lane >> 3);  // 16-byte chunk index 0..15 along the 128 dims
                        const uint32_t addr = kt + (chunk >> 3) * kSubTileBytes + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
                        uint32_t kb[4];
                        ldmatrix_x4(kb, addr);
                        const uint32_t a0[4] = {qa[2 * kp][0], 0u, qa[2 * kp][1], 0u};
                        const uint32_t a1[4] = {qa[2 * kp + 1][0], 0u, qa[2 * kp + 1][1], 0u};
                        mma_16816<TM>(sacc[nt], a0, kb[0], kb[1]);
                        mma_16816<TM>(sacc[nt], a1, kb[2], kb[3]);
                    }
                }
            }
            // ---- mask + online softmax (rows g; rows g+8 are padding) -----------------------------
            float mx = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int tok0 = kFp8 ? nt * 8 + t : nt * 8 + 2 * t, tok1 = kFp8 ? tok0 + 4 : tok0 + 1;      // tokens of this thread's two columns
                sacc[nt][0] = tok0 < valid ? sacc[nt][0] * p.scale_log2 : -INFINITY;
                sacc[nt][1] = tok1 < valid ? sacc[nt][1] * p.scale_log2 : -INFINITY;
                mx = fmaxf(mx, fmaxf(sacc[nt][0], sacc[nt][1]));
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m_run, mx);            // finite: every tile holds >= 1 valid token
            const float corr = fast_exp2(m_run - m_new);
            m_run = m_new;
            l_run *= corr;
            uint32_t pa[4];                                  // P as packed 16-bit pairs, per n-tile
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float p0 = fast_exp2(sacc[nt][0] - m_new), p1 = fast_exp2(sacc[nt][1] - m_new);
                l_run += p0 + p1;
                pa[nt] = pack2<TM>(p0, p1);
            }
            if (corr != 1.f) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { o[i][0] *= corr; o[i][1] *= corr; }
            }
            // V rows past the context may hold non-finite garbage: 0 * NaN would poison O
            if (valid < kTile) {
                if constexpr (kFp8) {
                    for (int r = valid + (lane >> 3); r < kTile; r += 4)
                        *reinterpret_cast<int4*>(smem + (vt - smem_base) + r * 128 + ((lane & 7) << 4)) = make_int4(0, 0, 0, 0);
                } else {
                    for (int r = valid + (lane >> 4); r < kTile; r += 2) {
                        const int ch = lane & 15;
                        *reinterpret_cast<int4*>(smem + (vt - smem_base) + (ch >> 3) * kSubTileBytes + r * 128 + (((ch & 7) ^ (r & 7)) << 4)) =
                            make_int4(0, 0, 0, 0);
                    }
                }
                __syncwarp();
            }
            // ---- O += P V : 2 k-steps (16 tokens) x 16 n-tiles (8 dims) ---------------------------
            if constexpr (kFp8) {
                // ldmatrix.m16n16.trans.b8 (sm_100a; fragment layout probed, tools/probes/ldmatrix_b8_probe.cu): with row addresses
                // r = 0..15 from lanes 0..15 (x2: a second matrix from lanes 16..31), lane (g, t) receives bytes (rows 4t..4t+3, column g)
                // and (rows 4t..4t+3, column g + 8).  Rows are tokens, columns dims of a 16-byte chunk; matrix row r is pointed at
                // token (r >> 2) + 4 (r & 3) of the k-step, so the low half-word of a register holds tokens (t, t + 4) = the k slots
                // (2t, 2t+1) that P's a0 carries, the high half-word tokens (t + 8, t + 12) = slots (2t+8, 2t+9) = a2.
#pragma unroll
                for (int ktk = 0; ktk < 2; ++ktk) {
                    const uint32_t a[4] = {pa[2 * ktk], 0u, pa[2 * ktk + 1], 0u};
                    const int r = lane & 15, row = ktk * 16 + (r >> 2) + 4 * (r & 3);
#pragma unroll
                    for (int cp = 0; cp < 4; ++cp) {            // two 16-dim chunks (four 8-dim n-tiles) per x2
                        const int chunk = 2 * cp + (lane >> 4);
                        uint32_t vb[4];
                        ldmatrix_x2_trans_b8(vb, vt + row * 128 + ((chunk ^ (row & 7)) << 4));
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t b0, b1;
                            cvt_e4m3x4(vb[j], b0, b1);
                            mma_16816<TM>(o[4 * cp + j], a, b0, b1);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int ktk = 0; ktk < 2; ++ktk) {
                    const uint32_t a[4] = {pa[2 * ktk], 0u, pa[2 * ktk + 1], 0u};
                    const int row = ktk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
                    for (int np = 0; np < 8; ++np) {            // two 8-dim n-tiles per ldmatrix.x4.trans
                        const int chunk = np * 2 + (lane >> 4);
                        const uint32_t addr = vt + (chunk >> 3) * kSubTileBytes + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
                        uint32_t vb[4];
                        ldmatrix_x4_trans(vb, addr);
                        mma_16816<TM>(o[2 * np], a, vb[0], vb[1]);
                        mma_16816<TM>(o[2 * np + 1], a, vb[2], vb[3]);
                    }
                }
            }
            __syncwarp();
            issue_one();                                     // refill the stage we just drained (tile consumed + kStages)
            ++consumed;
        }

        // ---- fp32 partial of this item straight from registers ----------------------------------------
        l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
        l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
        if (g < kGroup) {
            const int64_t slot = (((int64_t)b * p.num_kv_heads + h) * p.max_chunks + c) * kGroup + g;
            float* orow = p.part_o + slot * kHeadDim;
#pragma unroll
            for (int i = 0; i < 16; ++i) *reinterpret_cast<float2*>(orow + i * 8 + 2 * t) = make_float2(o[i][0], o[i][1]);
            if (t == 0) { p.part_ml[slot * 2] = m_run; p.part_ml[slot * 2 + 1] = l_run; }
        }
    }
}

// merge the chunk partials of one (sequence, head); one CTA of 128 threads (= dims) each
template <typename T, typename TOut, bool kK4 = false>
__global__ void __launch_bounds__(kHeadDim)
paged_attn_merge_kernel(TOut* __restrict__ out, const float* __restrict__ part_o, const float* __restrict__ part_ml,
                        const uint32_t* __restrict__ context_lens, unsigned int* __restrict__ counter, int num_heads,
                        int num_kv_heads, int group, int max_chunks) {
    pdl_wait();
    pdl_trigger();
    const int head = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
    const int chunk_tokens = (int)counter[4] * kPage;        // the chunk size the decode kernel settled on
    if (head == 0 && b == 0 && d == 0) *counter = 0u;        // leave the work queue ready for the next launch
    const int h = head / group, r = head - h * group;
    const int ctx = (int)context_lens[b];
    const int nchunks = (ctx + chunk_tokens - 1) / chunk_tokens;
    const int64_t base = ((int64_t)b * num_kv_heads + h) * max_chunks;
    // batches of 8 chunks with all loads of a batch in flight together (the kernel is pure L2 latency), folded online
    constexpr int kU = 8;
    float M = -INFINITY, acc = 0.f, L = 0.f;
    for (int c0 = 0; c0 < nchunks; c0 += kU) {
        float2 ml[kU];
        float o[kU];
#pragma unroll
        for (int i = 0; i < kU; ++i) {
            const bool ok = c0 + i < nchunks;
            const int64_t slot = (base + (ok ? c0 + i : c0)) * group + r;
            ml[i] = *reinterpret_cast<const float2*>(part_ml + slot * 2);
            o[i] = part_o[slot * kHeadDim + d];
            if (!ok) { ml[i] = make_float2(-INFINITY, 0.f); o[i] = 0.f; }
        }
        float bm = M;
#pragma unroll
        for (int i = 0; i < kU; ++i) bm = fmaxf(bm, ml[i].x);
        const float corr = exp2f(M - bm);            // first batch: exp2(-inf) = 0 on zero accumulators
        acc *= corr; L *= corr; M = bm;
#pragma unroll
        for (int i = 0; i < kU; ++i) {
            const float w = exp2f(ml[i].x - M);      // padding: exp2(-inf) = 0
            acc += w * o[i];
            L += w * ml[i].y;
        }
    }
    const float res = nchunks > 0 && L > 0.f ? acc / L : 0.f;
    out[((int64_t)b * num_heads + head) * kHeadDim + (kK4 ? (int)k4_index(d) : d)] = from_f32<TOut>(to_f32(from_f32<T>(res)));
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool make_kv_map(CUtensorMap* map, const void* cache, int64_t num_blocks, int kvh, int dtype, bool fp8) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error(kErrCuda, "paged_attention_decode: cuTensorMapEncodeTiled unavailable"); return false; }
    const cuuint64_t es = fp8 ? 1 : 2;                  // e4m3 bytes: a token row of one head is 128 B = one swizzle span = ONE box
    const cuuint64_t dims[4] = {(cuuint64_t)kHeadDim, (cuuint64_t)kvh, (cuuint64_t)kPage, (cuuint64_t)num_blocks};
    const cuuint64_t strides[3] = {(cuuint64_t)kHeadDim * es, (cuuint64_t)kvh * kHeadDim * es, (cuuint64_t)kPage * kvh * kHeadDim * es};
    const cuuint32_t box[4] = {(cuuint32_t)(fp8 ? 128 : 64), 1, (cuuint32_t)kTile, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16), 4,
                           const_cast<void*>(cache), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error(kErrCuda, "paged_attention_decode: cuTensorMapEncodeTiled failed (%d)", (int)r); return false; }
    return true;
}

int pick_chunk_pages(int num_seqs, int kvh, int max_blocks) {
    static const int forced = [] { const char* e = getenv("B200_ATTN_CHUNK_PAGES"); return e ? atoi(e) : 0; }();   // tuning aid
    if (forced >= 1 && forced <= kMaxChunkPages) return forced;
    // Measured (tools/attn_check.py, B = 32, ctx 4400): with 8 kv heads 8-page chunks win (5.82 TB/s vs 5.76 / 5.65 for 4 / 2
    // pages); with 4 or 1 kv heads per rank (tensor-parallel shards) 4-page chunks win by 7-8 % -- too few items leave the
    // dynamic queue nothing to balance.  So: 8 pages while that still gives every warp ~2 items, otherwise 4, and smaller
    // only when the whole problem is tiny.
    auto items = [&](int c) { return (int64_t)num_seqs * kvh * ((max_blocks + c - 1) / c); };
    int chunk = kMaxChunkPages;
    if (items(chunk) < 2ll * sm_count() * 6) chunk >>= 1;
    const int64_t want = 3ll * sm_count();
    while (chunk > 1 && items(chunk) < want) chunk >>= 1;
    return chunk;
}

template <typename T, typename TOut, int kGroup, bool kK4, bool kFp8>
void launch(const DecodeArgs& a, const CUtensorMap& kmap, const CUtensorMap& vmap, const DecodeParams& p, cudaStream_t st) {
    using PL = Plan<kFp8>;
    auto kern = paged_attn_decode_kernel<T, kGroup, kFp8>;
    ensure_dynamic_smem(reinterpret_cast<const void*>(kern), PL::kTotal);
    const int64_t max_items = (int64_t)a.num_seqs * a.num_kv_heads * p.max_chunks;
    const int64_t want = (max_items + PL::kWarps - 1) / PL::kWarps;
    const int grid = (int)(want < sm_count() ? want : sm_count());
    launch_pdl(kern, dim3(grid), dim3(PL::kThreads), PL::kTotal, st, kmap, vmap, p);
    count_launch();
    launch_pdl(paged_attn_merge_kernel<T, TOut, kK4>, dim3(a.num_heads, a.num_seqs), dim3(kHeadDim), 0, st,
               static_cast<TOut*>(a.out), (const float*)p.part_o, (const float*)p.part_ml, a.context_lens, p.counter, (int)a.num_heads,
               (int)a.num_kv_heads, (int)kGroup, (int)p.max_chunks);
    count_launch();
}

template <typename T, typename TOut, bool kK4, bool kFp8>
void launch_group(const DecodeArgs& a, const CUtensorMap& km, const CUtensorMap& vm, const DecodeParams& p, int group, cudaStream_t st) {
    switch (group) {
        case 1: launch<T, TOut, 1, kK4, kFp8>(a, km, vm, p, st); break;
        case 2: launch<T, TOut, 2, kK4, kFp8>(a, km, vm, p, st); break;
        case 4: launch<T, TOut, 4, kK4, kFp8>(a, km, vm, p, st); break;
        case 8: launch<T, TOut, 8, kK4, kFp8>(a, km, vm, p, st); break;
    }
}

template <bool kFp8>
void launch_dtype(const DecodeArgs& a, const CUtensorMap& km, const CUtensorMap& vm, const DecodeParams& p, int group, cudaStream_t st) {
    if (a.dtype == B200_BF16) {
        if (a.out_dtype == B200_F16_K4) launch_group<__nv_bfloat16, __half, true, kFp8>(a, km, vm, p, group, st);
        else if (a.out_dtype == B200_F16) launch_group<__nv_bfloat16, __half, false, kFp8>(a, km, vm, p, group, st);
        else launch_group<__nv_bfloat16, __nv_bfloat16, false, kFp8>(a, km, vm, p, group, st);
    } else {
        if (a.out_dtype == B200_F16_K4) launch_group<__half, __half, true, kFp8>(a, km, vm, p, group, st);
        else launch_group<__half, __half, false, kFp8>(a, km, vm, p, group, st);
    }
}

}  // namespace

bool paged_attention_decode_tma_supported(const DecodeArgs& a, float softcap, int window, int cache_dtype, int layout) {
    const int group = a.num_heads / a.num_kv_heads;
    const bool fp8 = cache_dtype == B200_FP8_E4M3 || cache_dtype == B200_U8;
    return layout == B200_KV_FLASH && a.head_dim == kHeadDim && a.block_size == kPage && (cache_dtype == a.dtype || fp8) &&
           (a.dtype == B200_BF16 || a.dtype == B200_F16) && softcap <= 0.f && window <= 0 &&
           (group == 1 || group == 2 || group == 4 || group == 8) && a.num_seqs <= kMaxSeqs && a.num_blocks > 0 &&
           ((uintptr_t)a.kc & 15) == 0 && ((uintptr_t)a.vc & 15) == 0 && ((uintptr_t)a.q & 3) == 0;
}

size_t paged_attention_decode_tma_workspace(int num_seqs, int num_heads, int head_dim, int max_blocks, int block_size) {
    (void)block_size;
    // worst case: 1-page chunks -> max_blocks partials per (sequence, head)
    return (size_t)num_seqs * num_heads * (size_t)max_blocks * ((size_t)head_dim * 4 + 8) + 512;
}

void paged_attention_decode_tma(const DecodeArgs& a, cudaStream_t st) {
    const int group = a.num_heads / a.num_kv_heads;
    DecodeParams p;
    p.q = a.q; p.block_tables = a.block_tables; p.context_lens = a.context_lens;
    p.num_seqs = a.num_seqs; p.num_heads = a.num_heads; p.num_kv_heads = a.num_kv_heads; p.max_blocks = a.max_blocks;
    p.chunk_pages = pick_chunk_pages(a.num_seqs, a.num_kv_heads, a.max_blocks);
    { static const int sw = [] { const char* e = getenv("B200_ATTN_STATIC"); return e ? atoi(e) : 1; }(); p.static_walk = a.fp8 ? sw : 0; }     // queue mode (see the kernel): CTA tickets with FP8 KV (+7 %; -2 % with 16-bit KV, 6 warps)
    { static const int ad = [] { const char* e = getenv("B200_ATTN_ADAPTIVE"); return e ? atoi(e) : 1; }(); p.adaptive_chunks = ad; }
    const int min_pages = p.adaptive_chunks && p.chunk_pages > 1 ? p.chunk_pages / 2 : p.chunk_pages;      // the kernel may halve the chunk
    p.max_chunks = (a.max_blocks + min_pages - 1) / min_pages;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    const size_t n_part = (size_t)a.num_seqs * a.num_kv_heads * p.max_chunks * group;
    const size_t need = n_part * (kHeadDim * 4 + 8) + 256;
    if (!a.workspace || a.workspace_bytes < need) {
        set_error(kErrBadArg, "paged_attention_decode: workspace too small (%zu < %zu bytes)", a.workspace_bytes, need);
        return;
    }
    // layout: [counter (256 B)] [partial O] [partial (m, l)]
    p.counter = static_cast<unsigned int*>(a.workspace);
    p.part_o = reinterpret_cast<float*>(static_cast<char*>(a.workspace) + 256);
    p.part_ml = p.part_o + n_part * kHeadDim;
    CUtensorMap km, vm;
    if (!make_kv_map(&km, a.kc, a.num_blocks, a.num_kv_heads, a.dtype, a.fp8) || !make_kv_map(&vm, a.vc, a.num_blocks, a.num_kv_heads, a.dtype, a.fp8)) return;
    if (a.fp8) launch_dtype<true>(a, km, vm, p, group, st);
    else launch_dtype<false>(a, km, vm, p, group, st);
    check_launch("paged_attention_decode");
}

}  // namespace b200
