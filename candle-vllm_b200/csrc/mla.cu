// mla.cu -- multi-head latent attention over the paged latent cache (DeepSeek-V2/V3, BASELINE config 5).
//
// Replaces attention_rs::mla::{concat_and_cache_mla, mla_paged_decode, mla_paged_prefill} as called from
// /root/reference/src/openai/models/layers/mla_attention.rs:479-552; cache shapes /root/reference/src/scheduler/cache_engine.rs:172-185:
//   ckv_cache [num_blocks, block_size, 1, kv_lora_rank = 512], kpe_cache [num_blocks, block_size, 1, qk_rope_head_dim = 64] (model dtype).
// "Absorbed" formulation: the caller has folded W_uk into q (q_absorbed [T, H, 512]) and applies W_uv to the result, so attention runs in
// the latent space: score(t) = (q_abs . ckv_t + q_pe . kpe_t) * sm_scale, out = sum_t softmax(score)_t ckv_t  -> [T, H, 512].
// Every head of a sequence reads the SAME 1152 bytes per token: the op is MQA with a 576-wide key and a 512-wide value.
//
// Decode kernel (tensor cores, flash-decoding split over the context): a CTA of 4 warps owns (sequence, chunk of <= 256 tokens,
// group of 16 heads).  32-token latent tiles land by TMA (nine {64 dims, 32 tokens} boxes, 128-byte swizzle, 2 stages of 36 KB); the
// 16 heads are the M rows of mma.sync m16n8k16.  Per tile: warp w computes S for tokens [8w, 8w+8) over all 576 dims (Q fragments by
// ldmatrix from shared memory), the warps agree on the running row maxima through shared memory, P goes to shared memory as 16-bit,
// and warp w accumulates O for latent dims [128w, 128w+128) over all 32 tokens.  Partials (m, l, O) per chunk are folded by a merge
// kernel.  Bytes dominate at 16 heads per rank (TP 8): 30 flop / byte; with all 128 heads on one GPU the same tiles serve 8 head groups
// from L2.
// Prefill (causal, chunked) runs a shape-generic kernel: one CTA per (query row, head).
#include <cuda.h>

#include <type_traits>

#include "attention.cuh"

namespace b200 {

namespace {

constexpr int kLat = 512, kRope = 64, kKd = kLat + kRope;      // latent / rope / key width
constexpr int kPg = 64, kTl = 32;                               // page, tile (tokens)
constexpr int kHg = 16;                                         // heads per CTA (MMA M)
constexpr int kSubB = kTl * 64 * 2;                             // one {64 dims, 32 tokens} box: 4 KB
constexpr int kStageB = 9 * kSubB;                              // 36 KB
constexpr int kChunkTok = 256;
constexpr int kQB = kHg * kKd * 2;                              // Q tile 16 x 576 x 2 B = 18 KB (row pitch 1152 B)
constexpr int kSmQ = 0, kSmStage = 18432, kSmP = kSmStage + 2 * kStageB, kSmMax = kSmP + kHg * kTl * 2, kSmBar = kSmMax + 4 * kHg * 4;
constexpr int kMlaSmem = kSmBar + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map),
                 "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) { __half2 v = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- concat_and_cache_mla: one CTA per token, 16-byte vectors -------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
concat_and_cache_mla_kernel(const T* __restrict__ ckv, const T* __restrict__ kpe, T* __restrict__ ckv_cache, T* __restrict__ kpe_cache,
                            const int64_t* __restrict__ slots, int lat, int rope) {
    const int t = blockIdx.x;
    const int64_t slot = slots[t];
    if (slot < 0) return;
    const int lv = lat * (int)sizeof(T) / 16, rv = rope * (int)sizeof(T) / 16;
    const int4* s0 = reinterpret_cast<const int4*>(ckv + (int64_t)t * lat);
    int4* d0 = reinterpret_cast<int4*>(ckv_cache + slot * lat);
    for (int i = threadIdx.x; i < lv; i += blockDim.x) d0[i] = s0[i];
    const int4* s1 = reinterpret_cast<const int4*>(kpe + (int64_t)t * rope);
    int4* d1 = reinterpret_cast<int4*>(kpe_cache + slot * rope);
    for (int i = threadIdx.x; i < rv; i += blockDim.x) d1[i] = s1[i];
}

struct MlaParams {
    const void* q_abs; const void* q_pe;       // [T, H, 512], [T, H, 64]
    const uint32_t* block_tables; const uint32_t* context_lens;
    float* part_o; float* part_ml;             // [T][H][chunks][512], [T][H][chunks][2]
    int num_heads, max_blocks, max_chunks;
    float scale_log2;
};

template <typename T>
__global__ void __launch_bounds__(128, 2)
mla_decode_kernel(const __grid_constant__ CUtensorMap cmap, const __grid_constant__ CUtensorMap pmap, const MlaParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int chunk = blockIdx.x, seq = blockIdx.y, hg = blockIdx.z;
    const int ctx = (int)p.context_lens[seq];
    const int tok_begin = chunk * kChunkTok;
    if (tok_begin >= ctx) return;                                     // uniform per CTA
    const int ntok = min(kChunkTok, ctx - tok_begin), ntiles = (ntok + kTl - 1) / kTl;
    const uint32_t sb = smem_u32(smem);
    const uint32_t bars = sb + kSmBar;
    float* wmax = reinterpret_cast<float*>(smem + kSmMax);            // [4 warps][16 heads]
    const int heads = min(kHg, p.num_heads - hg * kHg);

    if (threadIdx.x == 0) {
        mbar_init(bars, 1); mbar_init(bars + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // Q tile -> shared memory: row h = [q_abs (512) | q_pe (64)], 1152-byte rows, 16-byte chunks XOR-swizzled by the row (ldmatrix reads 8 rows
    // of one 16-byte column at a time: without the swizzle all 8 would hit the same banks)
    {
        const T* qa = static_cast<const T*>(p.q_abs) + ((int64_t)seq * p.num_heads + hg * kHg) * kLat;
        const T* qp = static_cast<const T*>(p.q_pe) + ((int64_t)seq * p.num_heads + hg * kHg) * kRope;
        for (int i = threadIdx.x; i < kHg * (kKd / 8); i += blockDim.x) {
            const int h = i / (kKd / 8), c = i % (kKd / 8);           // 72 chunks of 8 elements per row
            int4 v = make_int4(0, 0, 0, 0);
            if (h < heads) v = c < kLat / 8 ? *reinterpret_cast<const int4*>(qa + (int64_t)h * kLat + c * 8) : *reinterpret_cast<const int4*>(qp + (int64_t)h * kRope + (c - kLat / 8) * 8);
            *reinterpret_cast<int4*>(smem + kSmQ + h * (kKd * 2) + (((c & ~7) | ((c & 7) ^ (h & 7))) << 4)) = v;
        }
    }
    __syncthreads();
    const uint32_t* table = p.block_tables + (int64_t)seq * p.max_blocks;
    auto issue = [&](int tile, int stage) {
        const int tok = tok_begin + tile * kTl;
        const int blk = (int)table[tok / kPg], off = tok % kPg;
        const uint32_t dst = sb + kSmStage + stage * kStageB, bar = bars + stage * 8;
        mbar_expect_tx(bar, kStageB);
#pragma unroll
        for (int j = 0; j < 8; ++j) tma_load_3d(dst + j * kSubB, &cmap, bar, j * 64, off, blk);
        tma_load_3d(dst + 8 * kSubB, &pmap, bar, 0, off, blk);
    };
    if (threadIdx.x == 0) issue(0, 0);

    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;       // rows g and g + 8; l = this warp's tokens only

    for (int tl = 0; tl < ntiles; ++tl) {
        const int stage = tl & 1;
        if (threadIdx.x == 0 && tl + 1 < ntiles) issue(tl + 1, stage ^ 1);
        mbar_wait(bars + stage * 8, (tl >> 1) & 1);
        const uint32_t kt = sb + kSmStage + stage * kStageB;
        const int valid = min(kTl, ntok - tl * kTl);
        // ---- S for this warp's 8 tokens: 36 k-steps over [ckv | kpe] -------------------------------------------------------------
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const int row = warp * 8 + (lane & 7);                   // token row of the tile
#pragma unroll
            for (int kp = 0; kp < 18; ++kp) {                        // 32 dims (2 k-steps) per step
                const int chunk16 = kp * 4 + (lane >> 3);            // 16-byte chunk 0..71 along the 576 dims
                uint32_t kb[4];
                ldmatrix_x4(kb, kt + (chunk16 >> 3) * kSubB + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4));
                // Q fragments (A, row-major 16 x 16): ldmatrix x4 = {rows 0-7 | k lo, rows 8-15 | k lo, rows 0-7 | k hi, rows 8-15 | k hi}
                uint32_t a0[4], a1[4];
                {
                    const int qr = (lane & 7) + ((lane >> 3) & 1) * 8, qc = kp * 4 + (lane >> 4);           // k-step 2 kp: chunks 4 kp + {0, 1}
                    ldmatrix_x4(a0, sb + kSmQ + qr * (kKd * 2) + (((qc & ~7) | ((qc & 7) ^ (qr & 7))) << 4));
                    const int qc1 = qc + 2;                                                                  // k-step 2 kp + 1: chunks 4 kp + {2, 3}
                    ldmatrix_x4(a1, sb + kSmQ + qr * (kKd * 2) + (((qc1 & ~7) | ((qc1 & 7) ^ (qr & 7))) << 4));
                }
                mma_16816<T>(s, a0, kb[0], kb[1]);
                mma_16816<T>(s, a1, kb[2], kb[3]);
            }
        }
        // ---- mask, row maxima agreed across the four warps --------------------------------------------------------------------------
        const int tok = warp * 8 + 2 * t;
        s[0] = tok < valid ? s[0] * p.scale_log2 : -INFINITY; s[1] = tok + 1 < valid ? s[1] * p.scale_log2 : -INFINITY;
        s[2] = tok < valid ? s[2] * p.scale_log2 : -INFINITY; s[3] = tok + 1 < valid ? s[3] * p.scale_log2 : -INFINITY;
        float mx_lo = fmaxf(s[0], s[1]), mx_hi = fmaxf(s[2], s[3]);
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        if (t == 0) { wmax[warp * kHg + g] = mx_lo; wmax[warp * kHg + g + 8] = mx_hi; }
        __syncthreads();
        float mn_lo = m_lo, mn_hi = m_hi;
#pragma unroll
        for (int w = 0; w < 4; ++w) { mn_lo = fmaxf(mn_lo, wmax[w * kHg + g]); mn_hi = fmaxf(mn_hi, wmax[w * kHg + g + 8]); }
        // every tile holds >= 1 valid token (token 0 of the tile, in warp 0), so the new maxima are finite
        const float c_lo = fast_exp2(m_lo - mn_lo), c_hi = fast_exp2(m_hi - mn_hi);
        m_lo = mn_lo; m_hi = mn_hi;
        const float p0 = fast_exp2(s[0] - mn_lo), p1 = fast_exp2(s[1] - mn_lo), p2 = fast_exp2(s[2] - mn_hi), p3 = fast_exp2(s[3] - mn_hi);
        l_lo = l_lo * c_lo + p0 + p1; l_hi = l_hi * c_hi + p2 + p3;
        // P (16 heads x 32 tokens, 16-bit, 64-byte rows) -> shared memory for every warp's PV
        {
            T* ps = reinterpret_cast<T*>(smem + kSmP);
            *reinterpret_cast<uint32_t*>(ps + g * kTl + tok) = pack2<T>(p0, p1);
            *reinterpret_cast<uint32_t*>(ps + (g + 8) * kTl + tok) = pack2<T>(p2, p3);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { o[i][0] *= c_lo; o[i][1] *= c_lo; o[i][2] *= c_hi; o[i][3] *= c_hi; }
        // latent rows past the context may hold anything: zero them (their P is 0, but 0 * NaN would poison O); each warp cleans the
        // 128 dims it is about to read
        if (valid < kTl) {
            for (int r = valid + (lane >> 4); r < kTl; r += 2) {
                const int ch = lane & 15;                             // 16 chunks of 16 bytes = this warp's 128 dims = sub-tiles 2 warp, 2 warp + 1
                *reinterpret_cast<int4*>(smem + kSmStage + stage * kStageB + (2 * warp + (ch >> 3)) * kSubB + r * 128 + (((ch & 7) ^ (r & 7)) << 4)) = make_int4(0, 0, 0, 0);
            }
        }
        __syncthreads();                                             // P complete (and the zero fill of the other warps' rows is not needed by us)
        // ---- O[:, 128 warp .. +128) += P V : 2 k-steps (16 tokens) x 16 n-tiles -------------------------------------------------------
#pragma unroll
        for (int ktk = 0; ktk < 2; ++ktk) {
            uint32_t a[4];
            {
                // P is [16][32] 16-bit, 64-byte rows (4 chunks): A fragment of k-step ktk = chunks 2 ktk, 2 ktk + 1
                const int pr = (lane & 7) + ((lane >> 3) & 1) * 8, pc = 2 * ktk + (lane >> 4);
                ldmatrix_x4(a, sb + kSmP + pr * (kTl * 2) + pc * 16);
            }
            const int row = ktk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
            for (int np = 0; np < 8; ++np) {
                const int chunk16 = warp * 16 + np * 2 + (lane >> 4);   // 16-byte chunk along the 512 latent dims
                uint32_t vb[4];
                ldmatrix_x4_trans(vb, kt + (chunk16 >> 3) * kSubB + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4));
                mma_16816<T>(o[2 * np], a, vb[0], vb[1]);
                mma_16816<T>(o[2 * np + 1], a, vb[2], vb[3]);
            }
        }
        __syncthreads();                                             // stage, P and wmax are free again
    }

    // ---- partial of this (sequence, chunk): l summed over the warps through shared memory, O straight from registers --------------
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    if (t == 0) { wmax[warp * kHg + g] = l_lo; wmax[warp * kHg + g + 8] = l_hi; }
    __syncthreads();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int h = g + half * 8;
        if (h >= heads) continue;
        const int64_t slot = (((int64_t)seq * p.num_heads + hg * kHg + h) * p.max_chunks + chunk);
        float* orow = p.part_o + slot * kLat + warp * 128;
#pragma unroll
        for (int i = 0; i < 16; ++i) *reinterpret_cast<float2*>(orow + i * 8 + 2 * t) = make_float2(o[i][half * 2], o[i][half * 2 + 1]);
        if (warp == 0 && t == 0) {
            p.part_ml[slot * 2] = half ? m_hi : m_lo;
            p.part_ml[slot * 2 + 1] = wmax[h] + wmax[kHg + h] + wmax[2 * kHg + h] + wmax[3 * kHg + h];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
mla_merge_kernel(T* __restrict__ out, const float* __restrict__ part_o, const float* __restrict__ part_ml, const uint32_t* __restrict__ context_lens,
                 int num_heads, int max_chunks) {
    const int head = blockIdx.x, seq = blockIdx.y;
    const int ctx = (int)context_lens[seq];
    const int nchunks = (ctx + kChunkTok - 1) / kChunkTok;
    const int64_t base = ((int64_t)seq * num_heads + head) * max_chunks;
    float M = -INFINITY;
    for (int c = 0; c < nchunks; ++c) M = fmaxf(M, part_ml[(base + c) * 2]);
    float L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < nchunks; ++c) {
        const float w = exp2f(part_ml[(base + c) * 2] - M);
        L += w * part_ml[(base + c) * 2 + 1];
        const float4 v = *reinterpret_cast<const float4*>(part_o + (base + c) * kLat + threadIdx.x * 4);
        acc[0] += w * v.x; acc[1] += w * v.y; acc[2] += w * v.z; acc[3] += w * v.w;
    }
    const float inv = nchunks > 0 && L > 0.f ? 1.f / L : 0.f;
    T* o = out + ((int64_t)seq * num_heads + head) * kLat + threadIdx.x * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = from_f32<T>(acc[i] * inv);
}

// shape-generic MLA attention (prefill, odd latent sizes): one CTA per (query row, head), warps over the context, lanes over the dims
template <typename T>
__global__ void __launch_bounds__(128)
mla_generic_kernel(T* __restrict__ out, const T* __restrict__ q_abs, const T* __restrict__ q_pe, const T* __restrict__ ckv, const T* __restrict__ kpe,
                   const uint32_t* __restrict__ tables, const uint32_t* __restrict__ context_lens, const uint32_t* __restrict__ cu_q, int num_seqs,
                   int num_heads, int lat, int rope, int block_size, int max_blocks, float scale) {
    const int h = blockIdx.x, row = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int seq = row, L = 0;
    if (cu_q) {                    // prefill: sequence i owns rows cu_q[i] .. cu_q[i+1] = the last q_len positions of its context_lens[i] keys
        int lo = 0, hi = num_seqs;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cu_q[mid] <= (uint32_t)row) lo = mid; else hi = mid; }
        seq = lo;
        const int qlen = (int)(cu_q[seq + 1] - cu_q[seq]), klen = (int)context_lens[seq];
        L = klen - qlen + (row - (int)cu_q[seq]) + 1;
    } else L = (int)context_lens[seq];
    constexpr int kMaxPer = 16;                  // lat <= 512
    float qa[kMaxPer], acc[kMaxPer], qp[2];
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) { const int d = lane + 32 * i; qa[i] = d < lat ? to_f32(q_abs[((int64_t)row * num_heads + h) * lat + d]) : 0.f; acc[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 2; ++i) { const int d = lane + 32 * i; qp[i] = d < rope ? to_f32(q_pe[((int64_t)row * num_heads + h) * rope + d]) : 0.f; }
    float m = -INFINITY, l = 0.f;
    const uint32_t* table = tables + (int64_t)seq * max_blocks;
    for (int tk = warp; tk < L; tk += 4) {
        const int64_t slot = (int64_t)table[tk / block_size] * block_size + tk % block_size;
        float cv[kMaxPer], s = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i) { const int d = lane + 32 * i; cv[i] = d < lat ? to_f32(ckv[slot * lat + d]) : 0.f; s += qa[i] * cv[i]; }
#pragma unroll
        for (int i = 0; i < 2; ++i) { const int d = lane + 32 * i; if (d < rope) s += qp[i] * to_f32(kpe[slot * rope + d]); }
        s = warp_sum(s) * scale;
        const float mn = fmaxf(m, s), corr = __expf(m - mn), pw = __expf(s - mn);
#pragma unroll
        for (int i = 0; i < kMaxPer; ++i) acc[i] = acc[i] * corr + pw * cv[i];
        l = l * corr + pw; m = mn;
    }
    __shared__ float sm_m[4], sm_l[4], sm_acc[4][512];
    if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) sm_acc[warp][lane + 32 * i] = acc[i];
    __syncthreads();
    float gm = fmaxf(fmaxf(sm_m[0], sm_m[1]), fmaxf(sm_m[2], sm_m[3])), gl = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) gl += sm_m[w] == -INFINITY ? 0.f : sm_l[w] * __expf(sm_m[w] - gm);
    for (int d = threadIdx.x; d < lat; d += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) v += sm_m[w] == -INFINITY ? 0.f : sm_acc[w][d] * __expf(sm_m[w] - gm);
        out[((int64_t)row * num_heads + h) * lat + d] = from_f32<T>(gl > 0.f ? v / gl : 0.f);
    }
}

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
bool make_lat_map(CUtensorMap* map, const void* cache, int64_t num_blocks, int width, int dtype) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)kPg, (cuuint64_t)num_blocks};
    const cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)kPg * width * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)kTl, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(cache), dims, strides,
               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

}  // namespace b200

using namespace b200;

extern "C" {

void concat_and_cache_mla(const void* ckv, const void* k_pe, void* ckv_cache, void* kpe_cache, const int64_t* slot_mapping, int32_t num_tokens,
                          int32_t kv_lora_rank, int32_t qk_rope_head_dim, int32_t dtype, int64_t stream) {
    if (num_tokens == 0) return;
    B200_REQUIRE(ckv && k_pe && ckv_cache && kpe_cache && slot_mapping, kErrBadArg, "concat_and_cache_mla: null pointer");
    B200_REQUIRE(dtype == B200_BF16 || dtype == B200_F16, kErrUnsupported, "concat_and_cache_mla: dtype %d (f16 / bf16)", dtype);
    B200_REQUIRE(kv_lora_rank % 8 == 0 && qk_rope_head_dim % 8 == 0 && kv_lora_rank > 0 && qk_rope_head_dim > 0, kErrBadArg,
                 "concat_and_cache_mla: widths must be multiples of 8 (16-byte rows)");
    B200_REQUIRE((((uintptr_t)ckv | (uintptr_t)k_pe | (uintptr_t)ckv_cache | (uintptr_t)kpe_cache) & 15) == 0, kErrBadArg, "concat_and_cache_mla: 16-byte alignment");
    concat_and_cache_mla_kernel<__half><<<num_tokens, 128, 0, as_stream(stream)>>>((const __half*)ckv, (const __half*)k_pe, (__half*)ckv_cache, (__half*)kpe_cache,
                                                                                   slot_mapping, kv_lora_rank, qk_rope_head_dim);     // 16-bit bit copy
    count_launch();
    check_launch("concat_and_cache_mla");
}

size_t mla_paged_decode_workspace_bytes(int32_t num_seqs, int32_t num_heads, int32_t max_blocks_per_seq, int32_t block_size) {
    const size_t chunks = ((size_t)max_blocks_per_seq * block_size + kChunkTok - 1) / kChunkTok;
    return (size_t)num_seqs * num_heads * chunks * (kLat * 4 + 8) + 256;
}

void mla_paged_attention(void* out, const void* q_absorbed, const void* q_pe, const void* ckv_cache, const void* kpe_cache, const uint32_t* block_tables,
                         const uint32_t* context_lens, const uint32_t* cu_seqlens_q, int32_t num_seqs, int32_t num_rows, int32_t num_heads,
                         int32_t kv_lora_rank, int32_t qk_rope_head_dim, int32_t block_size, int32_t max_blocks_per_seq, int64_t num_blocks, float sm_scale,
                         int32_t dtype, void* workspace, size_t workspace_bytes, int64_t stream) {
    if (num_rows == 0 || num_seqs == 0) return;
    B200_REQUIRE(out && q_absorbed && q_pe && ckv_cache && kpe_cache && block_tables && context_lens, kErrBadArg, "mla_paged_attention: null pointer");
    B200_REQUIRE(dtype == B200_BF16 || dtype == B200_F16, kErrUnsupported, "mla_paged_attention: dtype %d (f16 / bf16)", dtype);
    B200_REQUIRE(num_heads > 0 && kv_lora_rank > 0 && kv_lora_rank <= 512 && qk_rope_head_dim > 0 && qk_rope_head_dim <= 64 && block_size > 0, kErrUnsupported,
                 "mla_paged_attention: latent %d (<= 512), rope %d (<= 64)", kv_lora_rank, qk_rope_head_dim);
    cudaStream_t st = as_stream(stream);
    const bool decode = cu_seqlens_q == nullptr;
    B200_REQUIRE(!decode || num_rows == num_seqs, kErrBadArg, "mla_paged_attention: decode takes one row per sequence");
    const size_t need = mla_paged_decode_workspace_bytes(num_seqs, num_heads, max_blocks_per_seq, block_size);
    if (decode && kv_lora_rank == kLat && qk_rope_head_dim == kRope && block_size == kPg && num_blocks > 0 && workspace && workspace_bytes >= need &&
        num_seqs <= 65535 && (((uintptr_t)q_absorbed | (uintptr_t)q_pe | (uintptr_t)ckv_cache | (uintptr_t)kpe_cache) & 15) == 0) {
        CUtensorMap cm, pm;
        if (make_lat_map(&cm, ckv_cache, num_blocks, kLat, dtype) && make_lat_map(&pm, kpe_cache, num_blocks, kRope, dtype)) {
            const int max_chunks = (int)(((size_t)max_blocks_per_seq * block_size + kChunkTok - 1) / kChunkTok);
            MlaParams p{q_absorbed, q_pe, block_tables, context_lens, static_cast<float*>(workspace),
                        static_cast<float*>(workspace) + (size_t)num_seqs * num_heads * max_chunks * kLat, num_heads, max_blocks_per_seq, max_chunks,
                        sm_scale * 1.4426950408889634f};
            const dim3 grid(max_chunks, num_seqs, ceil_div(num_heads, kHg));
            if (dtype == B200_BF16) {
                ensure_dynamic_smem(reinterpret_cast<const void*>(mla_decode_kernel<__nv_bfloat16>), kMlaSmem);
                mla_decode_kernel<__nv_bfloat16><<<grid, 128, kMlaSmem, st>>>(cm, pm, p);
                mla_merge_kernel<__nv_bfloat16><<<dim3(num_heads, num_seqs), 128, 0, st>>>((__nv_bfloat16*)out, p.part_o, p.part_ml, context_lens, num_heads, max_chunks);
            } else {
                ensure_dynamic_smem(reinterpret_cast<const void*>(mla_decode_kernel<__half>), kMlaSmem);
                mla_decode_kernel<__half><<<grid, 128, kMlaSmem, st>>>(cm, pm, p);
                mla_merge_kernel<__half><<<dim3(num_heads, num_seqs), 128, 0, st>>>((__half*)out, p.part_o, p.part_ml, context_lens, num_heads, max_chunks);
            }
            count_launch(2);
            check_launch("mla_paged_attention");
            return;
        }
    }
    const dim3 grid(num_heads, num_rows);
    if (dtype == B200_BF16)
        mla_generic_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((__nv_bfloat16*)out, (const __nv_bfloat16*)q_absorbed, (const __nv_bfloat16*)q_pe, (const __nv_bfloat16*)ckv_cache,
                                                               (const __nv_bfloat16*)kpe_cache, block_tables, context_lens, cu_seqlens_q, num_seqs, num_heads, kv_lora_rank,
                                                               qk_rope_head_dim, block_size, max_blocks_per_seq, sm_scale);
    else
        mla_generic_kernel<__half><<<grid, 128, 0, st>>>((__half*)out, (const __half*)q_absorbed, (const __half*)q_pe, (const __half*)ckv_cache, (const __half*)kpe_cache,
                                                        block_tables, context_lens, cu_seqlens_q, num_seqs, num_heads, kv_lora_rank, qk_rope_head_dim, block_size,
                                                        max_blocks_per_seq, sm_scale);
    count_launch();
    check_launch("mla_paged_attention");
}

}  // extern "C"
