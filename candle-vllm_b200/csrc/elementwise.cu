// elementwise.cu -- the small ops between the big kernels of a decoder layer (K15, K21) and the
// fused rope + q-cast + cache-write used by the decode engine.
//
// Reference semantics: candle_nn::ops::rms_norm (layers/qrmsnorm.rs:28-31), FusedRope::apply_inplace
// (layers/rotary_emb.rs:52-69, interleaved = rope_i for GGUF llama quantized_llama.rs:313-318),
// silu(w1 x) * (w3 x) (quantized_llama.rs:32-37), to_dtype casts (layers/attention.rs:971-975),
// tok_embeddings lookup (quantized_llama.rs:449), argmax sampling (pipeline.rs:2338).
#include <type_traits>

#include "common.cuh"
#include "mega.cuh"

namespace b200 {

// ---- rms_norm: one CTA per row; the row stays in registers (float4 per thread per 1024 columns) ----
template <typename TOut, bool kK4>
__device__ __forceinline__ void store4(TOut* o, int i, float a, float b, float c, float d) {
    // K4 order swaps the middle two of every aligned group of four
    if constexpr (kK4) { const float t = b; b = c; c = t; }
    if constexpr (sizeof(TOut) == 2) {
        uint2 v;
        if constexpr (std::is_same<TOut, __half>::value) {
            __half2 p0 = __halves2half2(from_f32<__half>(a), from_f32<__half>(b)), p1 = __halves2half2(from_f32<__half>(c), from_f32<__half>(d));
            v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
        } else {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
            v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
        }
        *reinterpret_cast<uint2*>(o + i) = v;
    } else {
        *reinterpret_cast<float4*>(o + i) = make_float4(a, b, c, d);
    }
}

template <typename TOut, bool kK4 = false>
__global__ void __launch_bounds__(256)
rms_norm_kernel(const float* __restrict__ x, const float* __restrict__ w, TOut* __restrict__ out, int n, float eps) {
    pdl_wait();
    pdl_trigger();
    constexpr int kMaxIt = 8;                             // rows up to 8192 columns stay in registers
    const int row = blockIdx.x;
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * n);
    const float4* wr = reinterpret_cast<const float4*>(w);
    const int nv = n >> 2;
    __shared__ float red[8];
    float4 v[kMaxIt];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < kMaxIt; ++it) {
        const int i = threadIdx.x + it * 256;
        if (i < nv) { v[it] = xr[i]; ss += v[it].x * v[it].x + v[it].y * v[it].y + v[it].z * v[it].z + v[it].w * v[it].w; }
    }
    for (int i = threadIdx.x + kMaxIt * 256; i < nv; i += 256) { const float4 t = xr[i]; ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w; }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float sc = rsqrtf(tot / (float)n + eps);
    TOut* o = out + (int64_t)row * n;
#pragma unroll
    for (int it = 0; it < kMaxIt; ++it) {
        const int i = threadIdx.x + it * 256;
        if (i < nv) { const float4 g = __ldg(wr + i); store4<TOut, kK4>(o, 4 * i, v[it].x * sc * g.x, v[it].y * sc * g.y, v[it].z * sc * g.z, v[it].w * sc * g.w); }
    }
    for (int i = threadIdx.x + kMaxIt * 256; i < nv; i += 256) {
        const float4 t = xr[i], g = __ldg(wr + i);
        store4<TOut, kK4>(o, 4 * i, t.x * sc * g.x, t.y * sc * g.y, t.z * sc * g.z, t.w * sc * g.w);
    }
}

// ---- RoPE in place on f32 q,k --------------------------------------------------------------
__global__ void fused_rope_f32_kernel(float* __restrict__ q, float* __restrict__ k,
                                      const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                      const int64_t* __restrict__ positions, int num_heads, int num_kv_heads,
                                      int head_dim, int interleaved) {
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.x;
    const int half = head_dim >> 1;
    const int64_t pos = positions[t];
    const int total = (num_heads + num_kv_heads) * half;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int h = i / half, j = i % half;
        float* p = h < num_heads ? q + ((int64_t)t * num_heads + h) * head_dim
                                 : k + ((int64_t)t * num_kv_heads + (h - num_heads)) * head_dim;
        const float c = cos_t[pos * half + j], s = sin_t[pos * half + j];
        const int i0 = interleaved ? 2 * j : j, i1 = interleaved ? 2 * j + 1 : j + half;
        const float a = p[i0], b = p[i1];
        p[i0] = a * c - b * s;
        p[i1] = a * s + b * c;
    }
}

// ---- fused: rope(q,k) + q -> 16-bit + k,v -> cache (flash layout) ----------------------------
// qkv f32 [T, (h + 2 kvh) * hd] is the packed output of the fused QKV projection.
template <typename T16, bool kFp8, bool kZeroSrc = false>
__global__ void __launch_bounds__(128)
rope_and_cache_kernel(float* __restrict__ qkv, T16* __restrict__ q_out, void* __restrict__ kc_, void* __restrict__ vc_,
                      const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                      const int64_t* __restrict__ positions, const int64_t* __restrict__ slot_mapping,
                      int num_heads, int num_kv_heads, int head_dim, int interleaved) {
    pdl_wait();
    pdl_trigger();
    // grid = (tokens, heads + 2 kv heads): one CTA per (token, head slot) so that even a 32-token decode batch fills the GPU
    const int t = blockIdx.x, hs = blockIdx.y;
    const int half = head_dim >> 1;
    const int64_t pos = positions[t];
    const int64_t slot = slot_mapping[t];
    const int row = (num_heads + 2 * num_kv_heads) * head_dim;
    float* src = qkv + (int64_t)t * row + (int64_t)hs * head_dim;
    const int kvn = num_kv_heads * head_dim;
    if (hs < num_heads + num_kv_heads) {
        // q or k head: rotate
        for (int j = threadIdx.x; j < half; j += blockDim.x) {
            const float c = cos_t[pos * half + j], s = sin_t[pos * half + j];
            const int i0 = interleaved ? 2 * j : j, i1 = interleaved ? 2 * j + 1 : j + half;
            const float a = src[i0], b = src[i1];
            const float r0 = a * c - b * s, r1 = a * s + b * c;
            if (hs < num_heads) {
                T16* o = q_out + ((int64_t)t * num_heads + hs) * head_dim;
                o[i0] = from_f32<T16>(r0);
                o[i1] = from_f32<T16>(r1);
            } else if (slot >= 0) {
                const int64_t base = slot * kvn + (int64_t)(hs - num_heads) * head_dim;
                // reference: k -> model dtype first (attention.rs:971-975), then the cache write casts again
                if constexpr (kFp8) {
                    static_cast<uint8_t*>(kc_)[base + i0] = f32_to_e4m3(to_f32(from_f32<T16>(r0)));
                    static_cast<uint8_t*>(kc_)[base + i1] = f32_to_e4m3(to_f32(from_f32<T16>(r1)));
                } else {
                    static_cast<T16*>(kc_)[base + i0] = from_f32<T16>(r0);
                    static_cast<T16*>(kc_)[base + i1] = from_f32<T16>(r1);
                }
            }
        }
    } else if (slot >= 0) {
        // v head: cast + cache write
        const int64_t base = slot * kvn + (int64_t)(hs - num_heads - num_kv_heads) * head_dim;
        for (int i = threadIdx.x; i < head_dim; i += blockDim.x) {
            if constexpr (kFp8) static_cast<uint8_t*>(vc_)[base + i] = f32_to_e4m3(to_f32(from_f32<T16>(src[i])));
            else static_cast<T16*>(vc_)[base + i] = from_f32<T16>(src[i]);
        }
    }
    if constexpr (kZeroSrc) {          // leave the split-K accumulator zeroed for the next layer's QKV GEMM
        __syncthreads();
        for (int i = threadIdx.x; i < head_dim; i += blockDim.x) src[i] = 0.f;
    }
}

// Vectorised form of the above for head_dim % 32 == 0 (every Llama / Qwen / DeepSeek head size): one thread owns 16 rotation
// pairs = 32 elements of one (token, head slot) -- interleaved: 32 consecutive elements; NeoX: [16u, 16u+16) and
// [half+16u, half+16u+16) -- so q and the bf16 / f16 cache get 16-byte stores and an FP8 (e4m3) cache gets 16-byte stores of 16
// values (north_star: "reshape_and_cache writes FP8 KV as coalesced vectorised HBM stores").  Same arithmetic and the same two
// roundings (f32 -> model dtype -> e4m3, attention.rs:971-975 then the cache write) as the scalar kernel.
template <typename T16> __device__ __forceinline__ uint32_t pack16(float a, float b);
template <> __device__ __forceinline__ uint32_t pack16<__half>(float a, float b) { const __half2 v = __halves2half2(from_f32<__half>(a), from_f32<__half>(b)); return *reinterpret_cast<const uint32_t*>(&v); }
template <> __device__ __forceinline__ uint32_t pack16<__nv_bfloat16>(float a, float b) { const __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&v); }
template <typename T16> __device__ __forceinline__ float round16(float a) { return to_f32(from_f32<T16>(a)); }
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
    const uint32_t lo = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}
// store 16 f32 values as 16 T16 (two 16-byte stores) or, kFp8, as 16 e4m3 bytes (one 16-byte store) after rounding to T16
template <typename T16, bool kFp8>
__device__ __forceinline__ void store16(void* base, int64_t elem_off, const float (&v)[16]) {
    if constexpr (kFp8) {
        uint4 o;
        o.x = pack_e4m3x4(round16<T16>(v[0]), round16<T16>(v[1]), round16<T16>(v[2]), round16<T16>(v[3]));
        o.y = pack_e4m3x4(round16<T16>(v[4]), round16<T16>(v[5]), round16<T16>(v[6]), round16<T16>(v[7]));
        o.z = pack_e4m3x4(round16<T16>(v[8]), round16<T16>(v[9]), round16<T16>(v[10]), round16<T16>(v[11]));
        o.w = pack_e4m3x4(round16<T16>(v[12]), round16<T16>(v[13]), round16<T16>(v[14]), round16<T16>(v[15]));
        *reinterpret_cast<uint4*>(static_cast<uint8_t*>(base) + elem_off) = o;
    } else {
        uint4 a, b;
        a.x = pack16<T16>(v[0], v[1]); a.y = pack16<T16>(v[2], v[3]); a.z = pack16<T16>(v[4], v[5]); a.w = pack16<T16>(v[6], v[7]);
        b.x = pack16<T16>(v[8], v[9]); b.y = pack16<T16>(v[10], v[11]); b.z = pack16<T16>(v[12], v[13]); b.w = pack16<T16>(v[14], v[15]);
        uint4* o = reinterpret_cast<uint4*>(static_cast<T16*>(base) + elem_off);
        o[0] = a; o[1] = b;
    }
}
__device__ __forceinline__ void load16(const float* p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float4 t = reinterpret_cast<const float4*>(p)[i]; v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
}
__device__ __forceinline__ void zero16(float* p) {
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// one work item = (token t, head slot hs, unit u): callable from the stand-alone kernel below and from the fused layer kernel
// 16 consecutive f32 of row `p` summed over the slabs of their tile (layer_mega.cu: split-K partial sums, added in slab order)
__device__ __forceinline__ void load16_slabs(const float* p, int64_t slab_stride, int slabs, float (&v)[16]) {
    // all loads first (independent L2 round trips), then the adds in slab order
    float4 t[kMegaMaxSlabs][4];
#pragma unroll
    for (int s = 0; s < kMegaMaxSlabs; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            t[s][i] = s < slabs ? __ldcg(reinterpret_cast<const float4*>(p + (int64_t)s * slab_stride) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 a = t[0][i];
#pragma unroll
        for (int s = 1; s < kMegaMaxSlabs; ++s) { a.x += t[s][i].x; a.y += t[s][i].y; a.z += t[s][i].z; a.w += t[s][i].w; }
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
}

template <typename T16, bool kFp8, bool kZeroSrc>
__device__ __forceinline__ void rope_cache_item(int item, const SlabInfo& si, float* __restrict__ qkv, T16* __restrict__ q_out, void* __restrict__ kc_, void* __restrict__ vc_,
                                                const float* __restrict__ cos_t, const float* __restrict__ sin_t, const int64_t* __restrict__ positions,
                                                const int64_t* __restrict__ slot_mapping, int num_heads, int num_kv_heads, int head_dim, int interleaved) {
    const int units = head_dim >> 5, slots_per_tok = num_heads + 2 * num_kv_heads;
    const int u = item % units, hs = (item / units) % slots_per_tok, t = item / (units * slots_per_tok);
    const int half = head_dim >> 1;
    const int64_t pos = positions[t], slot = slot_mapping[t];
    float* src = qkv + ((int64_t)t * slots_per_tok + hs) * head_dim;
    const int64_t kvn = (int64_t)num_kv_heads * head_dim;
    // element ranges of this unit: A = [a0, a0 + 16), B = [b0, b0 + 16)
    const bool rot = hs < num_heads + num_kv_heads;
    const int a0 = (rot && !interleaved) ? 16 * u : 32 * u, b0 = (rot && !interleaved) ? half + 16 * u : 32 * u + 16;
    float a[16], b[16];
    if (si.slab_stride == 0) {
        load16(src + a0, a);
        load16(src + b0, b);
    } else {
        // which 128-row tile of the fused QKV GEMM these columns belong to -> how many CTAs split it -> that many slabs
        const int sg = hs < num_heads ? 0 : (hs < num_heads + num_kv_heads ? 1 : 2);
        const int hl = hs - (sg == 0 ? 0 : (sg == 1 ? num_heads : num_heads + num_kv_heads));
        const uint32_t total = (uint32_t)si.n_tiles * (uint32_t)si.nsb;
        const uint32_t ta = (uint32_t)(si.seg_tile0[sg] + (hl * head_dim + a0) / 128), tb = (uint32_t)(si.seg_tile0[sg] + (hl * head_dim + b0) / 128);
        load16_slabs(src + a0, si.slab_stride, tile_slabs(ta, (uint32_t)si.nsb, total, (uint32_t)si.grid), a);
        load16_slabs(src + b0, si.slab_stride, tile_slabs(tb, (uint32_t)si.nsb, total, (uint32_t)si.grid), b);
    }
    if (rot) {
        float c[16], sn[16];
        load16(cos_t + pos * half + 16 * u, c);
        load16(sin_t + pos * half + 16 * u, sn);
        if (interleaved) {              // pair j = 16u + i lives at elements (2j, 2j + 1): i < 8 in A, i >= 8 in B
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float x0 = a[2 * i], x1 = a[2 * i + 1], y0 = b[2 * i], y1 = b[2 * i + 1];
                a[2 * i] = x0 * c[i] - x1 * sn[i]; a[2 * i + 1] = x0 * sn[i] + x1 * c[i];
                b[2 * i] = y0 * c[8 + i] - y1 * sn[8 + i]; b[2 * i + 1] = y0 * sn[8 + i] + y1 * c[8 + i];
            }
        } else {                        // pair j = 16u + i lives at (j, j + half) = (A[i], B[i])
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float x0 = a[i], x1 = b[i];
                a[i] = x0 * c[i] - x1 * sn[i]; b[i] = x0 * sn[i] + x1 * c[i];
            }
        }
    }
    if (hs < num_heads) {
        T16* o = q_out + ((int64_t)t * num_heads + hs) * head_dim;
        store16<T16, false>(o, a0, a);
        store16<T16, false>(o, b0, b);
    } else if (slot >= 0) {
        const bool is_k = hs < num_heads + num_kv_heads;
        const int64_t base = slot * kvn + (int64_t)(hs - num_heads - (is_k ? 0 : num_kv_heads)) * head_dim;
        void* cache = is_k ? kc_ : vc_;
        store16<T16, kFp8>(cache, base + a0, a);
        store16<T16, kFp8>(cache, base + b0, b);
    }
    if constexpr (kZeroSrc) { zero16(src + a0); zero16(src + b0); }     // leave the split-K accumulator zeroed for the next QKV GEMM
}

template <typename T16, bool kFp8, bool kZeroSrc>
__global__ void __launch_bounds__(128)
rope_and_cache_vec_kernel(const SlabInfo si, float* __restrict__ qkv, T16* __restrict__ q_out, void* __restrict__ kc_, void* __restrict__ vc_,
                          const float* __restrict__ cos_t, const float* __restrict__ sin_t, const int64_t* __restrict__ positions,
                          const int64_t* __restrict__ slot_mapping, int num_tokens, int num_heads, int num_kv_heads, int head_dim, int interleaved) {
    pdl_wait();
    pdl_trigger();
    const int items = num_tokens * (num_heads + 2 * num_kv_heads) * (head_dim >> 5);
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < items)
        rope_cache_item<T16, kFp8, kZeroSrc>(item, si, qkv, q_out, kc_, vc_, cos_t, sin_t, positions, slot_mapping, num_heads, num_kv_heads, head_dim, interleaved);
}

template <typename TOut, bool kK4 = false, bool kZeroSrc = false>
__global__ void silu_mul_kernel(float* __restrict__ g, float* __restrict__ u, TOut* __restrict__ out, int64_t n) {
    pdl_wait();
    pdl_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = g[i];
        out[kK4 ? k4_index(i) : i] = from_f32<TOut>(a / (1.f + __expf(-a)) * u[i]);
        if (kZeroSrc) { g[i] = 0.f; u[i] = 0.f; }           // leave the split-K accumulators zeroed for the next GEMM
    }
}

__global__ void add_f32_kernel(float* __restrict__ x, const float* __restrict__ y, int64_t n) {
    pdl_wait();
    pdl_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] += y[i];
}

template <typename TS, typename TD, bool kK4 = false>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
    pdl_wait();
    pdl_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        d[kK4 ? k4_index(i) : i] = from_f32<TD>(to_f32(s[i]));
}

__global__ void embedding_f32_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids,
                                     float* __restrict__ out, int dim) {
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.x;
    const float4* src = reinterpret_cast<const float4*>(table + ids[t] * (int64_t)dim);
    float4* dst = reinterpret_cast<float4*>(out + (int64_t)t * dim);
    for (int i = threadIdx.x; i < dim / 4; i += blockDim.x) dst[i] = __ldg(src + i);
}

// argmax per row (first maximal index, like candle's argmax); one CTA per row.
__global__ void __launch_bounds__(1024)
argmax_f32_kernel(const float* __restrict__ x, int32_t* __restrict__ out, int n) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x;
    const float* xr = x + (int64_t)row * n;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = xr[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    __shared__ float sv[32];
    __shared__ int si[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) out[row] = bi == 0x7fffffff ? 0 : bi;
    }
}

// stage 1 of greedy sampling: pairs[chunk][row] = (max logit, global index) over column chunk blockIdx.y of this rank's vocab
// shard.  Stage 2 (argmax_reduce_pairs_kernel) reduces over chunks -- and, tensor-parallel, over the ranks' gathered pairs.
__global__ void __launch_bounds__(256)
argmax_pair_kernel(const float* __restrict__ x, float2* __restrict__ pairs, int ld, int n, int index_offset) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x, rows = gridDim.x;
    const int per = (((n + (int)gridDim.y - 1) / (int)gridDim.y) + 3) & ~3;
    const int c0 = blockIdx.y * per, c1 = min(n, c0 + per);
    const float* xr = x + (int64_t)row * ld;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
        const float v = xr[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    __shared__ float sv[32];
    __shared__ int si[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        // an empty chunk reports (-inf, INT_MAX): it loses every comparison in stage 2
        if (threadIdx.x == 0) pairs[(int64_t)blockIdx.y * rows + row] = make_float2(best, __int_as_float(bi == 0x7fffffff ? bi : bi + index_offset));
    }
}

// gathered [world][rows] pairs -> token ids (first maximal global index wins ties, like a full argmax)
__global__ void argmax_reduce_pairs_kernel(const float2* __restrict__ gathered, int32_t* __restrict__ out, int rows, int world) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int r = 0; r < world; ++r) {
        const float2 pr = gathered[(int64_t)r * rows + row];
        const int idx = __float_as_int(pr.y);
        if (pr.x > best || (pr.x == best && idx < bi)) { best = pr.x; bi = idx; }
    }
    out[row] = bi == 0x7fffffff ? 0 : bi;
}

// logits [rows, ld]; only the first n columns of a row take part (vocab-parallel padding columns are never sampled)
void argmax_pairs(const float* logits, void* pairs, int rows, int ld, int n, int chunks, int index_offset, cudaStream_t st) {
    launch_pdl(argmax_pair_kernel, dim3(rows, chunks), dim3(256), 0, st, logits, static_cast<float2*>(pairs), ld, n, index_offset);
    count_launch();
    check_launch("argmax_pairs");
}
void argmax_reduce_pairs(const void* gathered, int32_t* out, int rows, int world, cudaStream_t st) {
    launch_pdl(argmax_reduce_pairs_kernel, dim3(ceil_div(rows, 128)), dim3(128), 0, st, static_cast<const float2*>(gathered), out, rows, world);
    count_launch();
    check_launch("argmax_reduce_pairs");
}

// VocabParallelLinear's gather epilogue (distributed.rs:1648-1664): gathered [world, rows, vocab_l] -> out [rows, vocab]
// (transpose(0, 1), reshape, narrow to the original vocab)
__global__ void gather_logits_transpose_kernel(const float* __restrict__ g, float* __restrict__ out, int world, int rows, int vocab_l, int vocab) {
    pdl_wait();
    pdl_trigger();
    const int64_t total = (int64_t)rows * vocab;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / vocab), col = (int)(i - (int64_t)r * vocab);
        const int w = col / vocab_l, c = col - w * vocab_l;
        out[i] = g[((int64_t)w * rows + r) * vocab_l + c];
    }
}
void gather_logits_transpose(const float* gathered, float* out, int world, int rows, int vocab_l, int vocab, cudaStream_t st) {
    launch_pdl(gather_logits_transpose_kernel, dim3(sm_count() * 4), dim3(256), 0, st, gathered, out, world, rows, vocab_l, vocab);
    count_launch();
    check_launch("gather_logits");
}

static inline int ew_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace b200

using namespace b200;

extern "C" {

void rms_norm(const float* x, const float* weight, void* out, int32_t rows, int32_t n, float eps,
              int32_t out_dtype, int64_t stream) {
    if (rows == 0) return;
    B200_REQUIRE(x && weight && out && rows > 0 && n > 0, kErrBadArg, "rms_norm: bad arguments");
    B200_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)weight & 15) == 0 && ((uintptr_t)out & 15) == 0 && n % 4 == 0, kErrBadArg,
                 "rms_norm: x, weight, out must be 16-byte aligned and n %% 4 == 0");
    if (out_dtype == B200_F32) launch_pdl(rms_norm_kernel<float, false>, dim3(rows), dim3(256), 0, as_stream(stream), x, weight, (float*)out, n, eps);
    else if (out_dtype == B200_F16) launch_pdl(rms_norm_kernel<__half, false>, dim3(rows), dim3(256), 0, as_stream(stream), x, weight, (__half*)out, n, eps);
    else if (out_dtype == B200_F16_K4) launch_pdl(rms_norm_kernel<__half, true>, dim3(rows), dim3(256), 0, as_stream(stream), x, weight, (__half*)out, n, eps);
    else if (out_dtype == B200_BF16) launch_pdl(rms_norm_kernel<__nv_bfloat16, false>, dim3(rows), dim3(256), 0, as_stream(stream), x, weight, (__nv_bfloat16*)out, n, eps);
    else { set_error(kErrUnsupported, "rms_norm: out dtype %d", out_dtype); return; }
    count_launch();
    check_launch("rms_norm");
}

void fused_rope_f32(float* q, float* k, const float* cos_t, const float* sin_t, const int64_t* positions,
                    int32_t num_tokens, int32_t num_heads, int32_t num_kv_heads, int32_t head_dim,
                    int32_t interleaved, int64_t stream) {
    if (num_tokens == 0) return;
    B200_REQUIRE(q && k && cos_t && sin_t && positions && head_dim % 2 == 0, kErrBadArg, "fused_rope: bad arguments");
    fused_rope_f32_kernel<<<num_tokens, 256, 0, as_stream(stream)>>>(q, k, cos_t, sin_t, positions, num_heads, num_kv_heads, head_dim, interleaved);
    count_launch();
    check_launch("fused_rope");
}

}  // extern "C"

namespace b200 {
static void rope_and_cache_any(float* qkv, const SlabInfo& si, void* q_out, void* key_cache, void* value_cache,
                    const float* cos_t, const float* sin_t, const int64_t* positions,
                    const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads,
                    int32_t num_kv_heads, int32_t head_dim, int32_t interleaved,
                    int32_t dtype, int32_t cache_dtype, bool zero_src, int64_t stream) {
    if (num_tokens == 0) return;
    B200_REQUIRE(qkv && q_out && key_cache && value_cache && cos_t && sin_t && positions && slot_mapping, kErrBadArg, "rope_and_cache: null pointer");
    const bool fp8 = cache_dtype == B200_FP8_E4M3 || cache_dtype == B200_U8;
    B200_REQUIRE(fp8 || cache_dtype == dtype, kErrUnsupported, "rope_and_cache: cache dtype %d vs dtype %d", cache_dtype, dtype);
    cudaStream_t st = as_stream(stream);
    // 16-byte loads / stores need 16-byte aligned rows: head_dim % 32 == 0 and aligned bases (cudaMalloc gives 256)
    B200_REQUIRE(si.slab_stride == 0 || head_dim % 32 == 0, kErrUnsupported, "rope_and_cache: slab input needs head_dim %% 32 == 0");
    const bool vec = head_dim % 32 == 0 && ((((uintptr_t)qkv | (uintptr_t)q_out | (uintptr_t)key_cache | (uintptr_t)value_cache | (uintptr_t)cos_t | (uintptr_t)sin_t) & 15) == 0);
    const int items = num_tokens * (num_heads + 2 * num_kv_heads) * (head_dim >> 5);
#define LAUNCH(T16, F8, Z) do { if (vec) launch_pdl(rope_and_cache_vec_kernel<T16, F8, Z>, dim3(ceil_div(items, 128)), dim3(128), 0, st, si, qkv, (T16*)q_out, key_cache, value_cache, cos_t, sin_t, positions, slot_mapping, (int)num_tokens, (int)num_heads, (int)num_kv_heads, (int)head_dim, (int)interleaved); \
        else launch_pdl(rope_and_cache_kernel<T16, F8, Z>, dim3(num_tokens, num_heads + 2 * num_kv_heads), dim3(64), 0, st, qkv, (T16*)q_out, key_cache, value_cache, cos_t, sin_t, positions, slot_mapping, num_heads, num_kv_heads, head_dim, interleaved); } while (0)
#define LAUNCH2(T16, F8) do { if (zero_src) LAUNCH(T16, F8, true); else LAUNCH(T16, F8, false); } while (0)
    if (dtype == B200_BF16) { if (fp8) LAUNCH2(__nv_bfloat16, true); else LAUNCH2(__nv_bfloat16, false); }
    else if (dtype == B200_F16) { if (fp8) LAUNCH2(__half, true); else LAUNCH2(__half, false); }
    else { set_error(kErrUnsupported, "rope_and_cache: dtype %d", dtype); return; }
#undef LAUNCH2
#undef LAUNCH
    count_launch();
    check_launch("rope_and_cache");
}

void rope_and_cache_impl(float* qkv, void* q_out, void* key_cache, void* value_cache, const float* cos_t, const float* sin_t,
                         const int64_t* positions, const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads, int32_t num_kv_heads,
                         int32_t head_dim, int32_t interleaved, int32_t dtype, int32_t cache_dtype, bool zero_src, int64_t stream) {
    SlabInfo si{};
    rope_and_cache_any(qkv, si, q_out, key_cache, value_cache, cos_t, sin_t, positions, slot_mapping, num_tokens, num_heads, num_kv_heads, head_dim,
                       interleaved, dtype, cache_dtype, zero_src, stream);
}
// the packed QKV row arrives as split-K slabs of the persistent layer kernel (nothing to re-zero)
void rope_and_cache_slabs(float* qkv, const SlabInfo& si, void* q_out, void* key_cache, void* value_cache, const float* cos_t, const float* sin_t,
                          const int64_t* positions, const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads, int32_t num_kv_heads,
                          int32_t head_dim, int32_t interleaved, int32_t dtype, int32_t cache_dtype, int64_t stream) {
    rope_and_cache_any(qkv, si, q_out, key_cache, value_cache, cos_t, sin_t, positions, slot_mapping, num_tokens, num_heads, num_kv_heads, head_dim,
                       interleaved, dtype, cache_dtype, false, stream);
}

// act = silu(gate) * up in the activation format the next linear reads (fp16 K4, or natural f16 / bf16); gate / up re-zeroed
void silu_mul_zero_src_fmt(float* gate, float* up, void* out, int64_t numel, int fmt, int64_t stream) {
    if (fmt == B200_F16_K4) launch_pdl(silu_mul_kernel<__half, true, true>, dim3(ew_grid(numel)), dim3(256), 0, as_stream(stream), gate, up, (__half*)out, numel);
    else if (fmt == B200_F16) launch_pdl(silu_mul_kernel<__half, false, true>, dim3(ew_grid(numel)), dim3(256), 0, as_stream(stream), gate, up, (__half*)out, numel);
    else if (fmt == B200_BF16) launch_pdl(silu_mul_kernel<__nv_bfloat16, false, true>, dim3(ew_grid(numel)), dim3(256), 0, as_stream(stream), gate, up, (__nv_bfloat16*)out, numel);
    else { set_error(kErrUnsupported, "silu_mul: activation format %d", fmt); return; }
    count_launch();
    check_launch("silu_mul");
}

void silu_mul_zero_src(float* gate, float* up, void* out_f16_k4, int64_t numel, int64_t stream) {
    launch_pdl(silu_mul_kernel<__half, true, true>, dim3(ew_grid(numel)), dim3(256), 0, as_stream(stream), gate, up, (__half*)out_f16_k4, numel);
    count_launch();
    check_launch("silu_mul");
}
}  // namespace b200

extern "C" {

void rope_and_cache(const float* qkv, void* q_out, void* key_cache, void* value_cache,
                    const float* cos_t, const float* sin_t, const int64_t* positions,
                    const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads,
                    int32_t num_kv_heads, int32_t head_dim, int32_t block_size, int32_t interleaved,
                    int32_t dtype, int32_t cache_dtype, int64_t stream) {
    (void)block_size;
    rope_and_cache_impl(const_cast<float*>(qkv), q_out, key_cache, value_cache, cos_t, sin_t, positions, slot_mapping, num_tokens,
                        num_heads, num_kv_heads, head_dim, interleaved, dtype, cache_dtype, false, stream);
}

void silu_mul(const float* gate, const float* up, void* out, int64_t numel, int32_t out_dtype, int64_t stream) {
    if (numel == 0) return;
    B200_REQUIRE(gate && up && out && numel > 0, kErrBadArg, "silu_mul: bad arguments");
    float* g = const_cast<float*>(gate); float* u = const_cast<float*>(up);
    if (out_dtype == B200_F32) silu_mul_kernel<float><<<ew_grid(numel), 256, 0, as_stream(stream)>>>(g, u, (float*)out, numel);
    else if (out_dtype == B200_F16) silu_mul_kernel<__half><<<ew_grid(numel), 256, 0, as_stream(stream)>>>(g, u, (__half*)out, numel);
    else if (out_dtype == B200_F16_K4) silu_mul_kernel<__half, true><<<ew_grid(numel), 256, 0, as_stream(stream)>>>(g, u, (__half*)out, numel);
    else { set_error(kErrUnsupported, "silu_mul: out dtype %d", out_dtype); return; }
    count_launch();
    check_launch("silu_mul");
}

void add_f32(float* x, const float* y, int64_t numel, int64_t stream) {
    if (numel == 0) return;
    B200_REQUIRE(x && y && numel > 0, kErrBadArg, "add_f32: bad arguments");
    launch_pdl(add_f32_kernel, dim3(ew_grid(numel)), dim3(256), 0, as_stream(stream), x, y, numel);
    count_launch();
    check_launch("add_f32");
}

void cast(const void* src, void* dst, int64_t numel, int32_t sd, int32_t dd, int64_t stream) {
    if (numel == 0) return;
    B200_REQUIRE(src && dst && numel > 0, kErrBadArg, "cast: bad arguments");
    cudaStream_t st = as_stream(stream);
    const int g = ew_grid(numel);
#define C(SD, ST, DD, DT) if (sd == SD && dd == DD) { cast_kernel<ST, DT><<<g, 256, 0, st>>>((const ST*)src, (DT*)dst, numel); count_launch(); check_launch("cast"); return; }
    C(B200_F32, float, B200_F16, __half) C(B200_F32, float, B200_BF16, __nv_bfloat16) C(B200_F32, float, B200_F32, float)
    C(B200_F16, __half, B200_F32, float) C(B200_BF16, __nv_bfloat16, B200_F32, float)
    C(B200_BF16, __nv_bfloat16, B200_F16, __half) C(B200_F16, __half, B200_BF16, __nv_bfloat16)
#undef C
#define C4(SD, ST) if (sd == SD && dd == B200_F16_K4) { cast_kernel<ST, __half, true><<<g, 256, 0, st>>>((const ST*)src, (__half*)dst, numel); count_launch(); check_launch("cast"); return; }
    C4(B200_F32, float) C4(B200_F16, __half) C4(B200_BF16, __nv_bfloat16)
#undef C4
    set_error(kErrUnsupported, "cast: %d -> %d unsupported", sd, dd);
}

void embedding_f32(const float* table, const int64_t* ids, float* out, int32_t num_tokens, int32_t dim, int64_t stream) {
    if (num_tokens == 0) return;
    B200_REQUIRE(table && ids && out && dim % 4 == 0, kErrBadArg, "embedding: bad arguments");
    launch_pdl(embedding_f32_kernel, dim3(num_tokens), dim3(256), 0, as_stream(stream), table, ids, out, (int)dim);
    count_launch();
    check_launch("embedding");
}

void argmax_f32(const float* logits, int32_t* out, int32_t rows, int32_t n, int64_t stream) {
    if (rows == 0) return;
    B200_REQUIRE(logits && out && n > 0, kErrBadArg, "argmax: bad arguments");
    launch_pdl(argmax_f32_kernel, dim3(rows), dim3(1024), 0, as_stream(stream), logits, out, (int)n);
    count_launch();
    check_launch("argmax");
}

}  // extern "C"
