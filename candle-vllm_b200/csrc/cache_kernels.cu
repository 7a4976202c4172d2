// cache_kernels.cu -- KV-cache block ops: copy_blocks (K4), swap_blocks (K5), reshape_and_cache (K3).
//
// All three are HBM-bound byte movers: 16-byte vectorised, coalesced accesses, grids sized in
// multiples of the SM count.  Bit-exact by construction (plain copies / RNE casts).
//
// Reference call sites: /root/reference/src/backend/cache.rs:15-165 (copy_blocks),
// /root/reference/src/scheduler/cache_engine.rs:345-399,527-535 (swap / copy),
// /root/reference/src/openai/models/layers/attention.rs:707-718,983-994 (cache write inside
// PagedAttention::forward); slot arithmetic /root/reference/src/openai/pipelines/inputs.rs:410-423.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------
// copy_blocks: grid.y = layer*2 + {K,V}; grid.x strides over (pair, 16-byte chunk).
// The (src,dst) table and the per-layer base pointers arrive as HOST arrays (cache.rs:112-114);
// they travel as a kernel-parameter struct (<= 192 pairs x 128 layers per launch; larger calls are cut into several
// launches).  No allocation, no device-side table, no host sync -> capture-safe.
// Semantics = the reference kernel's: every pair is an independent block copy and all pairs of a call run concurrently,
// so a block may not be both a source and a destination in one call (CoW never produces such chains,
// block_engine.rs).  Because a multi-launch call would silently turn that race into an ordering, a call whose pairs
// chain ACROSS launches is rejected instead.
// ------------------------------------------------------------------------------------------
constexpr int kMaxParamLayers = 128;
constexpr int kMaxParamPairs = 192;

struct CopyBlocksParams {
    uint64_t kptr[kMaxParamLayers];
    uint64_t vptr[kMaxParamLayers];
    int32_t src[kMaxParamPairs];
    int32_t dst[kMaxParamPairs];
};

__global__ void __launch_bounds__(256)
copy_blocks_kernel(const __grid_constant__ CopyBlocksParams p, int num_pairs, int64_t bytes_per_block) {
    const int layer = blockIdx.y >> 1;
    char* base = reinterpret_cast<char*>((blockIdx.y & 1) ? p.vptr[layer] : p.kptr[layer]);
    const int64_t vec_per_block = bytes_per_block >> 4;
    const int64_t total = vec_per_block * num_pairs;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int pair = (int)(i / vec_per_block);
        const int64_t off = i - (int64_t)pair * vec_per_block;
        const int4* s = reinterpret_cast<const int4*>(base + (int64_t)p.src[pair] * bytes_per_block) + off;
        int4* d = reinterpret_cast<int4*>(base + (int64_t)p.dst[pair] * bytes_per_block) + off;
        *d = __ldg(s);
    }
}

// bytewise path when blocks are not 16-byte aligned / sized (never for real KV shapes; kept for exactness)
__global__ void copy_blocks_tail_kernel(const __grid_constant__ CopyBlocksParams p, int num_pairs,
                                        int64_t bytes_per_block, int64_t tail_start) {
    const int layer = blockIdx.y >> 1;
    char* base = reinterpret_cast<char*>((blockIdx.y & 1) ? p.vptr[layer] : p.kptr[layer]);
    const int64_t tail = bytes_per_block - tail_start;
    for (int64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < tail * num_pairs; i += (int64_t)gridDim.x * blockDim.x) {
        const int pair = (int)(i / tail);
        const int64_t off = tail_start + (i - pair * tail);
        base[(int64_t)p.dst[pair] * bytes_per_block + off] = base[(int64_t)p.src[pair] * bytes_per_block + off];
    }
}

static void copy_blocks_impl(void* key_cache_ptrs, void* value_cache_ptrs, const void* block_mapping,
                             int32_t num_layers, int32_t num_pairs, int32_t numel_per_block,
                             int elem_size, int64_t stream) {
    if (num_layers == 0 || num_pairs == 0) return;          // cache.rs:41-44: nothing to do
    B200_REQUIRE(key_cache_ptrs && value_cache_ptrs && block_mapping, kErrBadArg, "copy_blocks: null pointer");
    B200_REQUIRE(num_layers > 0 && num_pairs > 0 && numel_per_block > 0, kErrBadArg,
                 "copy_blocks: negative size (layers=%d pairs=%d numel=%d)", num_layers, num_pairs, numel_per_block);
    const uint64_t* kp = static_cast<const uint64_t*>(key_cache_ptrs);
    const uint64_t* vp = static_cast<const uint64_t*>(value_cache_ptrs);
    const int64_t* map = static_cast<const int64_t*>(block_mapping);
    const int64_t bytes = (int64_t)numel_per_block * elem_size;
    for (int i = 0; i < num_pairs; ++i)
        B200_REQUIRE(map[2 * i] >= 0 && map[2 * i + 1] >= 0 && map[2 * i] <= INT32_MAX && map[2 * i + 1] <= INT32_MAX,
                     kErrBadArg, "copy_blocks: block id out of range in pair %d", i);
    if (num_pairs > kMaxParamPairs) {
        // several launches: a destination of one launch that is a source (or destination) of another would make the result
        // depend on how we cut the call -- refuse (see the header comment)
        std::vector<int64_t> srcs, dsts;
        srcs.reserve(num_pairs); dsts.reserve(num_pairs);
        for (int i = 0; i < num_pairs; ++i) { srcs.push_back(map[2 * i]); dsts.push_back(map[2 * i + 1]); }
        std::sort(srcs.begin(), srcs.end());
        std::sort(dsts.begin(), dsts.end());
        B200_REQUIRE(std::adjacent_find(dsts.begin(), dsts.end()) == dsts.end(), kErrBadArg, "copy_blocks: a block is the destination of two pairs");
        for (int64_t d : dsts)
            B200_REQUIRE(!std::binary_search(srcs.begin(), srcs.end(), d), kErrBadArg,
                         "copy_blocks: block %lld is both a source and a destination in one call of %d pairs", (long long)d, num_pairs);
    }
    for (int l0 = 0; l0 < num_layers; l0 += kMaxParamLayers) {
        const int nl = num_layers - l0 < kMaxParamLayers ? num_layers - l0 : kMaxParamLayers;
        for (int p0 = 0; p0 < num_pairs; p0 += kMaxParamPairs) {
            const int np = num_pairs - p0 < kMaxParamPairs ? num_pairs - p0 : kMaxParamPairs;
            CopyBlocksParams prm;
            for (int l = 0; l < nl; ++l) { prm.kptr[l] = kp[l0 + l]; prm.vptr[l] = vp[l0 + l]; }
            for (int i = 0; i < np; ++i) { prm.src[i] = (int32_t)map[2 * (p0 + i)]; prm.dst[i] = (int32_t)map[2 * (p0 + i) + 1]; }
            // 16-byte vector path needs every block start 16-byte aligned; otherwise copy bytewise
            bool aligned = (bytes & 15) == 0;
            for (int l = 0; l < nl && aligned; ++l) aligned = ((prm.kptr[l] | prm.vptr[l]) & 15) == 0;
            const int64_t vecs = aligned ? (bytes >> 4) * np : 0;
            if (vecs > 0) {
                int gx = (int)((vecs + 255) / 256);
                const int cap = 8 * sm_count() / (2 * nl) + 1;     // ~8 CTAs/SM in total
                if (gx > cap) gx = cap;
                copy_blocks_kernel<<<dim3(gx, 2 * nl), 256, 0, as_stream(stream)>>>(prm, np, bytes);
                count_launch();
            }
            if (!aligned) {
                copy_blocks_tail_kernel<<<dim3(8, 2 * nl), 256, 0, as_stream(stream)>>>(prm, np, bytes, 0);
                count_launch();
            }
        }
    }
    check_launch("copy_blocks");
}

// ------------------------------------------------------------------------------------------
// reshape_and_cache: one CTA per token; threads cover kvh*hd elements, 8 elements per thread.
// ------------------------------------------------------------------------------------------
template <typename TIn, typename TCache, bool kFp8>
__global__ void __launch_bounds__(256)
reshape_and_cache_kernel(const TIn* __restrict__ key, const TIn* __restrict__ value,
                         TCache* __restrict__ kc, TCache* __restrict__ vc,
                         const int64_t* __restrict__ slot_mapping, int num_kv_heads, int head_dim,
                         int block_size, int64_t key_stride, int64_t value_stride, int layout) {
    const int t = blockIdx.x;
    const int64_t slot = slot_mapping[t];
    if (slot < 0) return;                                   // _PAD_SLOT_ID (llm_engine.rs:94)
    const int n = num_kv_heads * head_dim;
    const int64_t blk = slot / block_size, off = slot % block_size;
    constexpr int x = 16 / (int)sizeof(TCache);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float kf = to_f32(key[t * key_stride + i]);
        const float vf = to_f32(value[t * value_stride + i]);
        TCache kq, vq;
        if constexpr (kFp8) { kq = f32_to_e4m3(kf); vq = f32_to_e4m3(vf); }
        else { kq = from_f32<TCache>(kf); vq = from_f32<TCache>(vf); }
        if (layout == B200_KV_FLASH) {
            kc[slot * n + i] = kq;
            vc[slot * n + i] = vq;
        } else {
            const int h = i / head_dim, d = i % head_dim;
            // K [nb, kvh, hd/x, bs, x]; V [nb, kvh, hd, bs]
            kc[(((blk * num_kv_heads + h) * (head_dim / x) + d / x) * block_size + off) * x + d % x] = kq;
            vc[((blk * num_kv_heads + h) * head_dim + d) * block_size + off] = vq;
        }
    }
}

// vectorised flash-layout fast path: same 16-bit dtype in and out, 16-byte loads/stores.
__global__ void __launch_bounds__(128)
reshape_and_cache_flash_vec_kernel(const int4* __restrict__ key, const int4* __restrict__ value,
                                   int4* __restrict__ kc, int4* __restrict__ vc,
                                   const int64_t* __restrict__ slot_mapping, int row_vecs,
                                   int64_t key_stride_vecs, int64_t value_stride_vecs) {
    const int t = blockIdx.x;
    const int64_t slot = slot_mapping[t];
    if (slot < 0) return;
    for (int i = threadIdx.x; i < row_vecs; i += blockDim.x) {
        kc[slot * row_vecs + i] = __ldg(key + t * key_stride_vecs + i);
        vc[slot * row_vecs + i] = __ldg(value + t * value_stride_vecs + i);
    }
}

// vectorised FP8 (e4m3) flash-layout path: one thread converts 16 consecutive elements of a token row and writes them with ONE
// 16-byte store (K and V), instead of 16 single-byte stores.  Same cast as the scalar kernel (RNE, saturating, scale 1.0).
template <typename TIn>
__global__ void __launch_bounds__(128)
reshape_and_cache_flash_fp8_vec_kernel(const TIn* __restrict__ key, const TIn* __restrict__ value, uint8_t* __restrict__ kc,
                                       uint8_t* __restrict__ vc, const int64_t* __restrict__ slot_mapping, int n,
                                       int64_t key_stride, int64_t value_stride) {
    const int t = blockIdx.x;
    const int64_t slot = slot_mapping[t];
    if (slot < 0) return;
    for (int i = threadIdx.x * 16; i < n; i += blockDim.x * 16) {
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const TIn* src = (which ? value + t * value_stride : key + t * key_stride) + i;
            float f[16];
            if constexpr (sizeof(TIn) == 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float4 q = reinterpret_cast<const float4*>(src)[j]; f[4 * j] = q.x; f[4 * j + 1] = q.y; f[4 * j + 2] = q.z; f[4 * j + 3] = q.w; }
            } else {
                TIn h[16];
                reinterpret_cast<uint4*>(h)[0] = reinterpret_cast<const uint4*>(src)[0];
                reinterpret_cast<uint4*>(h)[1] = reinterpret_cast<const uint4*>(src)[1];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = to_f32(h[j]);
            }
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t lo = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[4 * j], f[4 * j + 1]), __NV_SATFINITE, __NV_E4M3);
                const uint32_t hi = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(f[4 * j + 2], f[4 * j + 3]), __NV_SATFINITE, __NV_E4M3);
                w[j] = lo | (hi << 16);
            }
            *reinterpret_cast<uint4*>((which ? vc : kc) + slot * n + i) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

template <typename TIn>
static void launch_rac(const void* key, const void* value, void* kc, void* vc, const int64_t* slots,
                       int T, int kvh, int hd, int bs, int64_t ks, int64_t vs, int cache_dtype,
                       int in_dtype, int layout, cudaStream_t st) {
    const int n = kvh * hd;
    const int threads = n >= 256 ? 256 : ((n + 31) / 32) * 32;
    if (cache_dtype == B200_FP8_E4M3 || cache_dtype == B200_U8) {
        reshape_and_cache_kernel<TIn, uint8_t, true><<<T, threads, 0, st>>>(
            (const TIn*)key, (const TIn*)value, (uint8_t*)kc, (uint8_t*)vc, slots, kvh, hd, bs, ks, vs, layout);
    } else if (cache_dtype == B200_BF16) {
        reshape_and_cache_kernel<TIn, __nv_bfloat16, false><<<T, threads, 0, st>>>(
            (const TIn*)key, (const TIn*)value, (__nv_bfloat16*)kc, (__nv_bfloat16*)vc, slots, kvh, hd, bs, ks, vs, layout);
    } else if (cache_dtype == B200_F16) {
        reshape_and_cache_kernel<TIn, __half, false><<<T, threads, 0, st>>>(
            (const TIn*)key, (const TIn*)value, (__half*)kc, (__half*)vc, slots, kvh, hd, bs, ks, vs, layout);
    } else {
        reshape_and_cache_kernel<TIn, float, false><<<T, threads, 0, st>>>(
            (const TIn*)key, (const TIn*)value, (float*)kc, (float*)vc, slots, kvh, hd, bs, ks, vs, layout);
    }
}

// FlashInfer-style CSR page tables (indptr / indices / last_len, /root/reference/src/openai/pipelines/inputs.rs:477-506) -> the padded block
// table + context lengths every attention entry point of this library reads.  Device side, no host sync: graph-replay safe.
__global__ void csr_to_paged_kernel(const uint32_t* __restrict__ indptr, const uint32_t* __restrict__ indices, const uint32_t* __restrict__ last_len,
                                    uint32_t* __restrict__ tables, uint32_t* __restrict__ ctx, int num_seqs, int width, int block_size) {
    const int b = blockIdx.x;
    if (b >= num_seqs) return;
    const uint32_t p0 = indptr[b], pages = indptr[b + 1] - p0;
    for (int j = threadIdx.x; j < width; j += blockDim.x) tables[(int64_t)b * width + j] = (uint32_t)j < pages ? indices[p0 + j] : 0u;
    if (threadIdx.x == 0) ctx[b] = pages == 0 ? 0u : (pages - 1) * (uint32_t)block_size + last_len[b];      // kv_len (inputs.rs:523-531)
}

}  // namespace b200

using namespace b200;

extern "C" {

void copy_blocks_bf16(void* k, void* v, const void* m, int32_t nl, int32_t np, int32_t numel, int64_t s) { copy_blocks_impl(k, v, m, nl, np, numel, 2, s); }
void copy_blocks_f16(void* k, void* v, const void* m, int32_t nl, int32_t np, int32_t numel, int64_t s) { copy_blocks_impl(k, v, m, nl, np, numel, 2, s); }
void copy_blocks_f32(void* k, void* v, const void* m, int32_t nl, int32_t np, int32_t numel, int64_t s) { copy_blocks_impl(k, v, m, nl, np, numel, 4, s); }
void copy_blocks_u8(void* k, void* v, const void* m, int32_t nl, int32_t np, int32_t numel, int64_t s) { copy_blocks_impl(k, v, m, nl, np, numel, 1, s); }

void flashinfer_csr_to_paged(const uint32_t* indptr, const uint32_t* indices, const uint32_t* last_len, uint32_t* block_tables, uint32_t* context_lens,
                             int32_t num_seqs, int32_t max_blocks_per_seq, int32_t block_size, int64_t stream) {
    if (num_seqs == 0) return;
    B200_REQUIRE(indptr && indices && last_len && block_tables && context_lens && max_blocks_per_seq > 0 && block_size > 0, kErrBadArg, "flashinfer_csr_to_paged: bad arguments");
    csr_to_paged_kernel<<<num_seqs, 128, 0, as_stream(stream)>>>(indptr, indices, last_len, block_tables, context_lens, num_seqs, max_blocks_per_seq, block_size);
    count_launch();
    check_launch("flashinfer_csr_to_paged");
}

void swap_blocks(const void* src, void* dst, const int64_t* mapping, int32_t num_pairs,
                 int64_t bytes_per_block, int64_t stream) {
    if (num_pairs == 0) return;
    B200_REQUIRE(src && dst && mapping, kErrBadArg, "swap_blocks: null pointer");
    B200_REQUIRE(num_pairs > 0 && bytes_per_block > 0, kErrBadArg, "swap_blocks: bad sizes");
    // Coalesce runs where both src and dst advance by one block: a swap of a whole sequence is
    // usually a few long runs, i.e. a few large DMA transfers instead of one per block.
    int i = 0;
    while (i < num_pairs) {
        const int64_t s0 = mapping[2 * i], d0 = mapping[2 * i + 1];
        B200_REQUIRE(s0 >= 0 && d0 >= 0, kErrBadArg, "swap_blocks: negative block id");
        int run = 1;
        while (i + run < num_pairs && mapping[2 * (i + run)] == s0 + run && mapping[2 * (i + run) + 1] == d0 + run) ++run;
        cudaError_t e = cudaMemcpyAsync(static_cast<char*>(dst) + d0 * bytes_per_block,
                                        static_cast<const char*>(src) + s0 * bytes_per_block,
                                        (size_t)run * bytes_per_block, cudaMemcpyDefault, as_stream(stream));
        if (e != cudaSuccess) { set_error(kErrCuda, "swap_blocks: %s", cudaGetErrorString(e)); return; }
        i += run;
    }
}

void reshape_and_cache(const void* key, const void* value, void* key_cache, void* value_cache,
                       const int64_t* slot_mapping, int32_t num_tokens, int32_t num_kv_heads,
                       int32_t head_dim, int32_t block_size, int64_t key_stride, int64_t value_stride,
                       int32_t in_dtype, int32_t cache_dtype, int32_t layout, int64_t stream) {
    if (num_tokens == 0) return;
    B200_REQUIRE(key && value && key_cache && value_cache && slot_mapping, kErrBadArg, "reshape_and_cache: null pointer");
    B200_REQUIRE(num_tokens > 0 && num_kv_heads > 0 && head_dim > 0 && block_size > 0, kErrBadArg, "reshape_and_cache: bad sizes");
    B200_REQUIRE(layout == B200_KV_FLASH || layout == B200_KV_PAGED, kErrBadArg, "reshape_and_cache: bad layout %d", layout);
    const bool fp8 = cache_dtype == B200_FP8_E4M3 || cache_dtype == B200_U8;
    B200_REQUIRE(fp8 || cache_dtype == in_dtype || in_dtype == B200_F32, kErrUnsupported,
                 "reshape_and_cache: cache dtype %d incompatible with input dtype %d", cache_dtype, in_dtype);
    const int esz = fp8 ? 1 : (cache_dtype == B200_F32 ? 4 : 2);
    B200_REQUIRE(layout == B200_KV_FLASH || head_dim % (16 / esz) == 0, kErrBadArg, "reshape_and_cache: head_dim %% x != 0");
    cudaStream_t st = as_stream(stream);
    const int n = num_kv_heads * head_dim;
    if (layout == B200_KV_FLASH && !fp8 && cache_dtype == in_dtype && in_dtype != B200_F32 && n % 8 == 0 &&
        key_stride % 8 == 0 && value_stride % 8 == 0 && ((uintptr_t)key & 15) == 0 && ((uintptr_t)value & 15) == 0 &&
        ((uintptr_t)key_cache & 15) == 0 && ((uintptr_t)value_cache & 15) == 0) {
        reshape_and_cache_flash_vec_kernel<<<num_tokens, 128, 0, st>>>(
            (const int4*)key, (const int4*)value, (int4*)key_cache, (int4*)value_cache, slot_mapping, n / 8,
            key_stride / 8, value_stride / 8);
    } else if (layout == B200_KV_FLASH && fp8 && n % 16 == 0 && (key_stride * (in_dtype == B200_F32 ? 4 : 2)) % 16 == 0 &&
               (value_stride * (in_dtype == B200_F32 ? 4 : 2)) % 16 == 0 &&
               ((((uintptr_t)key | (uintptr_t)value | (uintptr_t)key_cache | (uintptr_t)value_cache)) & 15) == 0 &&
               (in_dtype == B200_F32 || in_dtype == B200_BF16 || in_dtype == B200_F16)) {
        const int threads = n / 16 >= 128 ? 128 : ((n / 16 + 31) / 32) * 32;
        if (in_dtype == B200_F32)
            reshape_and_cache_flash_fp8_vec_kernel<float><<<num_tokens, threads, 0, st>>>((const float*)key, (const float*)value, (uint8_t*)key_cache, (uint8_t*)value_cache, slot_mapping, n, key_stride, value_stride);
        else if (in_dtype == B200_BF16)
            reshape_and_cache_flash_fp8_vec_kernel<__nv_bfloat16><<<num_tokens, threads, 0, st>>>((const __nv_bfloat16*)key, (const __nv_bfloat16*)value, (uint8_t*)key_cache, (uint8_t*)value_cache, slot_mapping, n, key_stride, value_stride);
        else
            reshape_and_cache_flash_fp8_vec_kernel<__half><<<num_tokens, threads, 0, st>>>((const __half*)key, (const __half*)value, (uint8_t*)key_cache, (uint8_t*)value_cache, slot_mapping, n, key_stride, value_stride);
    } else if (in_dtype == B200_F32) {
        launch_rac<float>(key, value, key_cache, value_cache, slot_mapping, num_tokens, num_kv_heads, head_dim, block_size, key_stride, value_stride, cache_dtype, in_dtype, layout, st);
    } else if (in_dtype == B200_BF16) {
        launch_rac<__nv_bfloat16>(key, value, key_cache, value_cache, slot_mapping, num_tokens, num_kv_heads, head_dim, block_size, key_stride, value_stride, cache_dtype, in_dtype, layout, st);
    } else if (in_dtype == B200_F16) {
        launch_rac<__half>(key, value, key_cache, value_cache, slot_mapping, num_tokens, num_kv_heads, head_dim, block_size, key_stride, value_stride, cache_dtype, in_dtype, layout, st);
    } else {
        set_error(kErrUnsupported, "reshape_and_cache: unsupported input dtype %d", in_dtype);
        return;
    }
    count_launch();
    check_launch("reshape_and_cache");
}

}  // extern "C"
