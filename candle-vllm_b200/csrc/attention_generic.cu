// attention_generic.cu -- shape-generic paged attention (decode + varlen causal prefill).
//
// One CTA per (query row, head); warps stride over the context, lanes over head_dim; online softmax
// in fp32.  This is the catch-all path: both KV layouts (flash and legacy paged,
// /root/reference/src/scheduler/cache_engine.rs:298-341), FP8(e4m3) caches, softcap, sliding
// window, any head_dim <= 256.  The TMA-staged split-KV kernel in attention_decode.cu takes over
// for the flash-layout bf16/f16 decode shapes that dominate the metric.
//
// Semantics: NaiveAttention::forward (/root/reference/src/openai/models/mod.rs:1268-1307) over
// K/V gathered through the block table; metadata /root/reference/src/openai/pipelines/inputs.rs
// :351-367 (prefill), :552-568 (decode).
#include "attention.cuh"

namespace b200 {

template <typename TC, bool kFp8>
__device__ __forceinline__ float load_cache(const TC* p) {
    if constexpr (kFp8) return e4m3_to_f32(*reinterpret_cast<const uint8_t*>(p));
    else return to_f32(*p);
}

constexpr int kGenWarps = 4;
constexpr int kMaxDimPerLane = 8;   // head_dim <= 256

template <typename T, typename TC, bool kFp8, typename TOut>
__global__ void __launch_bounds__(kGenWarps * 32)
paged_attention_generic_kernel(TOut* __restrict__ out, const T* __restrict__ q, const TC* __restrict__ kc,
                               const TC* __restrict__ vc, const GenericAttnArgs a, const int k4) {
    const int h = blockIdx.x, row = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = a.head_dim;
    int seq, L;   // L = number of visible keys [0, L)
    if (a.prefill) {
        int lo = 0, hi = a.num_seqs;              // largest s with cu_q[s] <= row
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.cu_q[mid] <= (uint32_t)row) lo = mid; else hi = mid; }
        seq = lo;
        const int qlen = (int)(a.cu_q[seq + 1] - a.cu_q[seq]);
        const int klen = (int)(a.cu_k[seq + 1] - a.cu_k[seq]);
        L = klen - qlen + (row - (int)a.cu_q[seq]) + 1;
    } else {
        seq = row;
        L = (int)a.context_lens[seq];
    }
    const int start = (a.window > 0 && L > a.window) ? L - a.window : 0;
    const int kvh = h / (a.num_heads / a.num_kv_heads);
    const uint32_t* table = a.block_tables + (int64_t)seq * a.max_blocks;

    float qr[kMaxDimPerLane];
#pragma unroll
    for (int i = 0; i < kMaxDimPerLane; ++i) {
        const int d = lane + 32 * i;
        qr[i] = d < hd ? to_f32(q[((int64_t)row * a.num_heads + h) * hd + d]) : 0.f;
    }
    float m = -INFINITY, l = 0.f, acc[kMaxDimPerLane];
#pragma unroll
    for (int i = 0; i < kMaxDimPerLane; ++i) acc[i] = 0.f;
    constexpr int x = 16 / (int)sizeof(TC);

    for (int t = start + warp; t < L; t += kGenWarps) {
        const int64_t blk = table[t / a.block_size];
        const int off = t % a.block_size;
        float kv[kMaxDimPerLane], vv[kMaxDimPerLane];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxDimPerLane; ++i) {
            const int d = lane + 32 * i;
            if (d < hd) {
                int64_t ki, vi;
                if (a.layout == B200_KV_FLASH) {
                    ki = vi = ((blk * a.block_size + off) * a.num_kv_heads + kvh) * hd + d;
                } else {
                    ki = (((blk * a.num_kv_heads + kvh) * (hd / x) + d / x) * a.block_size + off) * x + d % x;
                    vi = ((blk * a.num_kv_heads + kvh) * hd + d) * a.block_size + off;
                }
                kv[i] = load_cache<TC, kFp8>(kc + ki);
                vv[i] = load_cache<TC, kFp8>(vc + vi);
                s += qr[i] * kv[i];
            } else { kv[i] = vv[i] = 0.f; }
        }
        s = warp_sum(s) * a.scale;
        if (a.softcap > 0.f) s = tanhf(s / a.softcap) * a.softcap;
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn), p = __expf(s - mn);
#pragma unroll
        for (int i = 0; i < kMaxDimPerLane; ++i) acc[i] = acc[i] * corr + p * vv[i];
        l = l * corr + p;
        m = mn;
    }

    __shared__ float sm_m[kGenWarps], sm_l[kGenWarps];
    __shared__ float sm_acc[kGenWarps][256];
    if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
#pragma unroll
    for (int i = 0; i < kMaxDimPerLane; ++i) sm_acc[warp][lane + 32 * i] = acc[i];
    __syncthreads();
    float gm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kGenWarps; ++w) gm = fmaxf(gm, sm_m[w]);
    float gl = 0.f;
#pragma unroll
    for (int w = 0; w < kGenWarps; ++w) gl += sm_m[w] == -INFINITY ? 0.f : sm_l[w] * __expf(sm_m[w] - gm);
    for (int d = threadIdx.x; d < hd; d += blockDim.x) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < kGenWarps; ++w) o += sm_m[w] == -INFINITY ? 0.f : sm_acc[w][d] * __expf(sm_m[w] - gm);
        const float r = gl > 0.f ? o / gl : 0.f;
        // round to the model dtype first (the reference returns `dtype`), then to the requested out type
        out[((int64_t)row * a.num_heads + h) * hd + (k4 ? (int)k4_index(d) : d)] = from_f32<TOut>(to_f32(from_f32<T>(r)));
    }
}

template <typename T, typename TOut>
static void launch_generic(void* out, const void* q, const void* kc, const void* vc, const GenericAttnArgs& a,
                           int rows, int cache_dtype, int k4, cudaStream_t st) {
    dim3 grid(a.num_heads, rows);
    if (cache_dtype == B200_FP8_E4M3 || cache_dtype == B200_U8)
        paged_attention_generic_kernel<T, uint8_t, true, TOut><<<grid, kGenWarps * 32, 0, st>>>((TOut*)out, (const T*)q, (const uint8_t*)kc, (const uint8_t*)vc, a, k4);
    else
        paged_attention_generic_kernel<T, T, false, TOut><<<grid, kGenWarps * 32, 0, st>>>((TOut*)out, (const T*)q, (const T*)kc, (const T*)vc, a, k4);
    count_launch();
}

void paged_attention_generic(void* out, const void* q, const void* kc, const void* vc, const GenericAttnArgs& a,
                             int rows, int dtype, int cache_dtype, int out_dtype, cudaStream_t st) {
    const int k4 = out_dtype == B200_F16_K4;
    if (dtype == B200_BF16) {
        if (out_dtype == B200_F16 || k4) launch_generic<__nv_bfloat16, __half>(out, q, kc, vc, a, rows, cache_dtype, k4, st);
        else launch_generic<__nv_bfloat16, __nv_bfloat16>(out, q, kc, vc, a, rows, cache_dtype, 0, st);
    } else {
        launch_generic<__half, __half>(out, q, kc, vc, a, rows, cache_dtype, k4, st);
    }
    check_launch("paged_attention_generic");
}

}  // namespace b200
