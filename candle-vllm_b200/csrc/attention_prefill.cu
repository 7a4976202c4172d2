// attention_prefill.cu -- (chunked) prefill attention over the paged KV cache on tensor cores.
//
// PagedAttention::forward with is_prefill = true (call sites /root/reference/src/openai/models/layers/attention.rs:707-718, :983-994;
// metadata /root/reference/src/openai/pipelines/inputs.rs:133-148, :351-367; chunk size 8192, llm_engine.rs:95): varlen causal
// attention where sequence i owns q rows cu_seqlens_q[i] .. cu_seqlens_q[i+1] = the LAST q_len positions of its k_len context, and
// ALL keys / values (cached prefix + this chunk) are read from the paged cache the caller has already written.
//
// The generic kernel (attention_generic.cu) gives one CTA to every (query row, head) and re-reads the whole context from global
// memory for each: O(T^2) HBM / L2 traffic, unusable at an 8 K chunk.  Here a CTA owns a 64-row query tile of one head (4 warps x 16
// rows, Q fragments in registers) and walks the context page by page: each 64-token KV page of the kv head lands ONCE per tile in
// shared memory by TMA (the decode kernel's 4-D tensor map: box {64 dims, 1 head, 32 tokens, 1 block}, 128-byte swizzle, 2 stages),
// S = Q K^T and O += P V run on mma.sync m16n8k16 (bf16 / f16, fp32 accumulate) with the online softmax in registers, causal
// bottom-right mask (+ optional sliding window) applied only on the pages that touch the diagonal.  FlashAttention-2 dataflow; the
// arithmetic is ~4 * T_q * T_k * 128 flop per head against T_k * 512 B of page reads per 64 rows, i.e. tensor-bound.
// Covers: flash layout, head_dim 128, block 64, 16-bit cache of the model dtype; everything else stays on the generic kernel.
#include <cuda.h>

#include <type_traits>

#include "attention.cuh"

namespace b200 {

namespace {

constexpr int kHd = 128, kPg = 64, kHalf = 32;
constexpr int kQTile = 64, kPfWarps = 4, kPfThreads = kPfWarps * 32;
constexpr int kSub = kHalf * 64 * 2;                    // one swizzled sub-tile: 32 tokens x 64 dims x 2 B = 4 KB
constexpr int kStage = 2 * 4 * kSub;                    // 64 tokens: 2 halves x (K lo, K hi, V lo, V hi) = 32 KB
constexpr int kPfSmem = 2 * kStage + 64;

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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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

struct PrefillParams {
    const void* q; void* out;
    const uint32_t* block_tables; const uint32_t* cu_q; const uint32_t* cu_k;
    int num_seqs, num_heads, num_kv_heads, max_blocks, window;
    float scale_log2;
};

template <typename T>
__global__ void __launch_bounds__(kPfThreads, 2)
paged_attn_prefill_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap, const PrefillParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bars = smem_base + 2 * kStage;

    // which (sequence, 64-row query tile) is this CTA?  tiles are numbered sequence by sequence
    int seq = -1, tile = 0;
    {
        int acc = 0;
        for (int s = 0; s < p.num_seqs; ++s) {
            const int nt = ((int)(p.cu_q[s + 1] - p.cu_q[s]) + kQTile - 1) / kQTile;
            if ((int)blockIdx.x < acc + nt) { seq = s; tile = (int)blockIdx.x - acc; break; }
            acc += nt;
        }
    }
    if (seq < 0) return;
    const int head = blockIdx.y, kvh = head / (p.num_heads / p.num_kv_heads);
    const int q_begin = (int)p.cu_q[seq], qlen = (int)p.cu_q[seq + 1] - q_begin, klen = (int)(p.cu_k[seq + 1] - p.cu_k[seq]);
    const int q0 = tile * kQTile;                                   // first query row of the tile within the sequence
    const int rows = min(kQTile, qlen - q0);
    const int pos0 = klen - qlen + q0;                              // context position of tile row 0 (bottom-right aligned causal mask)
    const int last_pos = pos0 + rows - 1;
    const int first_key = p.window > 0 ? max(0, pos0 - p.window + 1) : 0;
    const int page_lo = first_key / kPg, page_hi = last_pos / kPg;  // pages [page_lo, page_hi] hold every visible key
    const uint32_t* table = p.block_tables + (int64_t)seq * p.max_blocks;

    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int page, int stage) {            // one elected thread: 8 boxes = one 64-token page of this kv head
        const uint32_t dst = smem_base + stage * kStage, bar = bars + stage * 8;
        const int blk = (int)table[page];
        mbar_expect_tx(bar, kStage);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            tma_load_4d(dst + h * 4 * kSub, &kmap, bar, 0, kvh, h * kHalf, blk);
            tma_load_4d(dst + h * 4 * kSub + kSub, &kmap, bar, 64, kvh, h * kHalf, blk);
            tma_load_4d(dst + h * 4 * kSub + 2 * kSub, &vmap, bar, 0, kvh, h * kHalf, blk);
            tma_load_4d(dst + h * 4 * kSub + 3 * kSub, &vmap, bar, 64, kvh, h * kHalf, blk);
        }
    };
    if (threadIdx.x == 0) issue(page_lo, 0);

    // ---- Q fragments of this warp's 16 rows (rows r0 = 16 warp + g and r0 + 8), zero beyond the sequence ------------------------
    const T* qbase = static_cast<const T*>(p.q);
    const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
    uint32_t qa[8][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int c = ks * 16 + 2 * t;
        const T* ql = qbase + ((int64_t)(q_begin + q0 + r_lo) * p.num_heads + head) * kHd + c;
        const T* qh = qbase + ((int64_t)(q_begin + q0 + r_hi) * p.num_heads + head) * kHd + c;
        qa[ks][0] = r_lo < rows ? *reinterpret_cast<const uint32_t*>(ql) : 0u;
        qa[ks][1] = r_hi < rows ? *reinterpret_cast<const uint32_t*>(qh) : 0u;
        qa[ks][2] = r_lo < rows ? *reinterpret_cast<const uint32_t*>(ql + 8) : 0u;
        qa[ks][3] = r_hi < rows ? *reinterpret_cast<const uint32_t*>(qh + 8) : 0u;
    }
    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    const int pos_lo = pos0 + r_lo, pos_hi = pos0 + r_hi;          // last visible key of each of this thread's two rows
    const int warp_min_pos = pos0 + warp * 16, warp_max_pos = pos0 + warp * 16 + 15;

    for (int page = page_lo, it = 0; page <= page_hi; ++page, ++it) {
        const int stage = it & 1;
        if (threadIdx.x == 0 && page + 1 <= page_hi) issue(page + 1, stage ^ 1);      // that stage was drained before the last __syncthreads
        mbar_wait(bars + stage * 8, (it >> 1) & 1);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int tok0 = page * kPg + h * kHalf;                 // first token of this 32-token half
            if (tok0 > warp_max_pos) break;                          // entirely in the future of every row of this warp
            if (p.window > 0 && tok0 + kHalf - 1 < warp_min_pos - p.window + 1) continue;      // entirely before every row's window
            const uint32_t kt = smem_base + stage * kStage + h * 4 * kSub, vt = kt + 2 * kSub;
            // ---- S = Q K^T : 4 n-tiles (8 tokens) x 8 k-steps -----------------------------------------------------------------
            float sacc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int row = nt * 8 + (lane & 7);
#pragma unroll
                for (int kp = 0; kp < 4; ++kp) {
                    const int chunk = kp * 4 + (lane >> 3);
                    uint32_t kb[4];
                    ldmatrix_x4(kb, kt + (chunk >> 3) * kSub + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
                    mma_16816<T>(sacc[nt], qa[2 * kp], kb[0], kb[1]);
                    mma_16816<T>(sacc[nt], qa[2 * kp + 1], kb[2], kb[3]);
                }
            }
            // ---- mask (only where the half touches the diagonal / window edge / context end) + online softmax --------------------
            const bool need_mask = tok0 + kHalf - 1 > warp_min_pos || (p.window > 0 && tok0 < warp_max_pos - p.window + 1);
            float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int tok = tok0 + nt * 8 + 2 * t + e;
                    bool ok_lo = true, ok_hi = true;
                    if (need_mask) {
                        ok_lo = tok <= pos_lo && (p.window <= 0 || tok > pos_lo - p.window);
                        ok_hi = tok <= pos_hi && (p.window <= 0 || tok > pos_hi - p.window);
                    }
                    sacc[nt][e] = ok_lo ? sacc[nt][e] * p.scale_log2 : -INFINITY;
                    sacc[nt][2 + e] = ok_hi ? sacc[nt][2 + e] * p.scale_log2 : -INFINITY;
                    mx_lo = fmaxf(mx_lo, sacc[nt][e]);
                    mx_hi = fmaxf(mx_hi, sacc[nt][2 + e]);
                }
            }
            mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
            mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
            const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
            // a row that has seen no visible key yet keeps m = -inf: use 0 as the subtrahend so that exp2(-inf - 0) = 0, never NaN
            const float sub_lo = mn_lo == -INFINITY ? 0.f : mn_lo, sub_hi = mn_hi == -INFINITY ? 0.f : mn_hi;
            const float c_lo = fast_exp2(m_lo - sub_lo), c_hi = fast_exp2(m_hi - sub_hi);
            m_lo = mn_lo; m_hi = mn_hi;
            l_lo *= c_lo; l_hi *= c_hi;
            uint32_t pa[4][2];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float p0 = fast_exp2(sacc[nt][0] - sub_lo), p1 = fast_exp2(sacc[nt][1] - sub_lo);
                const float p2 = fast_exp2(sacc[nt][2] - sub_hi), p3 = fast_exp2(sacc[nt][3] - sub_hi);
                l_lo += p0 + p1; l_hi += p2 + p3;
                pa[nt][0] = pack2<T>(p0, p1); pa[nt][1] = pack2<T>(p2, p3);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) { o[i][0] *= c_lo; o[i][1] *= c_lo; o[i][2] *= c_hi; o[i][3] *= c_hi; }
            // ---- O += P V : 2 k-steps (16 tokens) x 16 n-tiles (8 dims).  V rows past the context hold whatever the block contains:
            // their P is exactly 0 only if V is finite, so rows beyond klen are zeroed in shared memory first (as in the decode kernel)
            if (tok0 + kHalf > klen) {
                __syncwarp();
                for (int r = max(klen - tok0, 0) + (lane >> 4); r < kHalf; r += 2) {
                    const int ch = lane & 15;
                    *reinterpret_cast<int4*>(smem + (vt - smem_base) + (ch >> 3) * kSub + r * 128 + (((ch & 7) ^ (r & 7)) << 4)) = make_int4(0, 0, 0, 0);
                }
                __syncwarp();
            }
#pragma unroll
            for (int ktk = 0; ktk < 2; ++ktk) {
                const uint32_t a[4] = {pa[2 * ktk][0], pa[2 * ktk][1], pa[2 * ktk + 1][0], pa[2 * ktk + 1][1]};
                const int row = ktk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
#pragma unroll
                for (int np = 0; np < 8; ++np) {
                    const int chunk = np * 2 + (lane >> 4);
                    uint32_t vb[4];
                    ldmatrix_x4_trans(vb, vt + (chunk >> 3) * kSub + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
                    mma_16816<T>(o[2 * np], a, vb[0], vb[1]);
                    mma_16816<T>(o[2 * np + 1], a, vb[2], vb[3]);
                }
            }
        }
        __syncthreads();                                             // every warp is done with this stage before it is refilled
    }

    // ---- normalise and store ------------------------------------------------------------------------------------------------------
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float inv_lo = l_lo > 0.f ? 1.f / l_lo : 0.f, inv_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
    T* obase = static_cast<T*>(p.out);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (r_lo < rows)
            *reinterpret_cast<uint32_t*>(obase + ((int64_t)(q_begin + q0 + r_lo) * p.num_heads + head) * kHd + i * 8 + 2 * t) = pack2<T>(o[i][0] * inv_lo, o[i][1] * inv_lo);
        if (r_hi < rows)
            *reinterpret_cast<uint32_t*>(obase + ((int64_t)(q_begin + q0 + r_hi) * p.num_heads + head) * kHd + i * 8 + 2 * t) = pack2<T>(o[i][2] * inv_hi, o[i][3] * inv_hi);
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
bool make_kv_map(CUtensorMap* map, const void* cache, int64_t num_blocks, int kvh, int dtype) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)kHd, (cuuint64_t)kvh, (cuuint64_t)kPg, (cuuint64_t)num_blocks};
    const cuuint64_t strides[3] = {(cuuint64_t)kHd * 2, (cuuint64_t)kvh * kHd * 2, (cuuint64_t)kPg * kvh * kHd * 2};
    const cuuint32_t box[4] = {64, 1, (cuuint32_t)kHalf, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(cache), dims, strides,
               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool paged_attention_prefill_tc_supported(int head_dim, int block_size, int dtype, int cache_dtype, int layout, float softcap, int64_t num_blocks,
                                          const void* q, const void* kc, const void* vc) {
    return head_dim == kHd && block_size == kPg && (dtype == B200_BF16 || dtype == B200_F16) && cache_dtype == dtype && layout == B200_KV_FLASH &&
           softcap <= 0.f && num_blocks > 0 && (((uintptr_t)q & 3) == 0) && (((uintptr_t)kc | (uintptr_t)vc) & 15) == 0 && encode_fn() != nullptr;
}

// q / out [total_q, H, 128] of `dtype`; total_q_tiles_max = upper bound on sum_i ceil(q_len_i / 64) (the kernel finds its tile from the
// DEVICE-side cu_seqlens, so a graph captured with padded metadata replays correctly)
void paged_attention_prefill_tc(void* out, const void* q, const void* kc, const void* vc, const uint32_t* block_tables, const uint32_t* cu_q,
                                const uint32_t* cu_k, int num_seqs, int total_q, int num_heads, int num_kv_heads, int max_blocks, int64_t num_blocks,
                                float scale, int window, int dtype, cudaStream_t st) {
    CUtensorMap km, vm;
    if (!make_kv_map(&km, kc, num_blocks, num_kv_heads, dtype) || !make_kv_map(&vm, vc, num_blocks, num_kv_heads, dtype)) {
        set_error(kErrCuda, "paged_attention_prefill: cuTensorMapEncodeTiled failed");
        return;
    }
    PrefillParams p{q, out, block_tables, cu_q, cu_k, num_seqs, num_heads, num_kv_heads, max_blocks, window, scale * 1.4426950408889634f};
    const int tiles = total_q / kQTile + num_seqs;                   // >= sum of ceil(q_len / 64)
    if (dtype == B200_BF16) {
        ensure_dynamic_smem(reinterpret_cast<const void*>(paged_attn_prefill_kernel<__nv_bfloat16>), kPfSmem);
        paged_attn_prefill_kernel<__nv_bfloat16><<<dim3(tiles, num_heads), kPfThreads, kPfSmem, st>>>(km, vm, p);
    } else {
        ensure_dynamic_smem(reinterpret_cast<const void*>(paged_attn_prefill_kernel<__half>), kPfSmem);
        paged_attn_prefill_kernel<__half><<<dim3(tiles, num_heads), kPfThreads, kPfSmem, st>>>(km, vm, p);
    }
    count_launch();
    check_launch("paged_attention_prefill");
}

}  // namespace b200
