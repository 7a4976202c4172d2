// moe.cu -- fused mixture-of-experts path on GGUF expert tensors: top-k routing, expert sort and the grouped
// dequant-GEMM.  Mirrors the reference's FusedMoe::forward over attention_rs::{topk::topk_softmax, moe::moe_gemm_gguf}
// (/root/reference/src/openai/models/layers/moe.rs:35-45, :425-480, :1429-1482; quantized_qwen3_moe.rs:70-141; BASELINE config 4:
// Qwen3-30B-A3B Q4_K, 128 experts, top-8).
//
//   topk_softmax            softmax over the router logits of a token, the k largest probabilities and their expert ids
//   sort_expert_assignments the flattened [T * k] expert ids sorted ascending + the permutation (pair index = token * k + slot): a
//                           STABLE counting sort on the device (the reference sorts with candle's sort; order inside an expert does not
//                           change any result)
//   moe_gemm_gguf           out[pair, :] = W[expert(pair)] . x[row(pair)] (* topk_weight[pair]); row(pair) = pair / k when x has one row
//                           per token (gate / up), pair when it has one per pair (down).
// Decode is a weight stream: per layer only the experts that were hit are read, each ONCE -- the routed rows of an expert are gathered
// (fp16, K4 order) next to each other, and the tcgen05 dequant-GEMM of qmatmul_tc.cu runs in "grouped" mode over a device-side list of
// (expert, 128-row weight tile, chunk of <= 32 rows) items: weight tile = UMMA A (dequantised straight into TMEM), the expert's rows =
// UMMA B through one TMA box, results scattered back to their pair rows with the routing weight folded in.  Nothing on this path
// synchronises with the host or allocates: the item list is built on the device from this step's routing (graph-capture safe).
#include <cuda_fp8.h>

#include <type_traits>

#include "moe.cuh"
#include "qmatmul.cuh"

namespace b200 {

namespace {

// ---- routing ------------------------------------------------------------------------------------------------------------------
// one warp per token; E <= 1024 experts (32 per lane), k <= 32
__global__ void __launch_bounds__(128)
topk_softmax_kernel(const float* __restrict__ logits, float* __restrict__ w_out, uint32_t* __restrict__ id_out, int tokens, int E, int k) {
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 4 + warp;
    if (t >= tokens) return;
    constexpr int kPer = 32;
    float v[kPer];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int e = lane + 32 * i;
        v[i] = e < E ? logits[(int64_t)t * E + e] : -INFINITY;
        mx = fmaxf(mx, v[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) { v[i] = lane + 32 * i < E ? __expf(v[i] - mx) : 0.f; sum += v[i]; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = 0; j < k; ++j) {                      // k rounds of arg-max; ties -> the smallest expert id
        float best = -1.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int e = lane + 32 * i;
            if (e < E && (v[i] > best || (v[i] == best && e < bi))) { best = v[i]; bi = e; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { w_out[(int64_t)t * k + j] = best * inv; id_out[(int64_t)t * k + j] = (uint32_t)bi; }
        if ((bi & 31) == lane) v[bi >> 5] = -1.f;      // remove the winner (probabilities are >= 0)
    }
}

// stable counting sort of the flattened expert ids (one CTA; pairs <= 64 Ki, experts <= 1024) and the grouped-GEMM item list
__global__ void __launch_bounds__(1024)
moe_sort_kernel(const uint32_t* __restrict__ ids, uint32_t* __restrict__ sorted_expert, uint32_t* __restrict__ sorted_pair, int pairs, int E) {
    pdl_wait();
    pdl_trigger();
    __shared__ int count[1024], start[1025];
    for (int e = threadIdx.x; e < E; e += blockDim.x) count[e] = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) atomicAdd(&count[min(ids[p], (uint32_t)E - 1)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int e = 0; e < E; ++e) { start[e] = acc; acc += count[e]; }
        start[E] = acc;
    }
    __syncthreads();
    // stable placement: thread e walks the pairs in order for its expert(s) -- pairs are few (decode) or the walk is short per expert
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        if (count[e] == 0) continue;
        int pos = start[e];
        for (int p = 0; p < pairs; ++p)
            if (min(ids[p], (uint32_t)E - 1) == (uint32_t)e) { sorted_expert[pos] = (uint32_t)e; sorted_pair[pos] = (uint32_t)p; ++pos; }
    }
}

// items of the grouped GEMM from the SORTED expert ids: for every expert segment, every 128-row weight tile, every chunk of 32 rows
__global__ void __launch_bounds__(256)
moe_items_kernel(const uint32_t* __restrict__ sorted_expert, int pairs, int E, int n, MoeItem* __restrict__ items, int* __restrict__ num_items,
                 int max_items) {
    pdl_wait();
    pdl_trigger();
    __shared__ int seg_start[1025], item_base[1025];
    // segment starts by binary search (sorted input)
    for (int e = threadIdx.x; e <= E; e += blockDim.x) {
        int lo = 0, hi = pairs;                          // first position with expert >= e
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_expert[mid] < (uint32_t)e) lo = mid + 1; else hi = mid; }
        seg_start[e] = lo;
    }
    __syncthreads();
    const int tiles = (n + 127) / 128;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int e = 0; e < E; ++e) { item_base[e] = acc; acc += ((seg_start[e + 1] - seg_start[e] + 31) / 32) * tiles; }
        item_base[E] = acc;
        *num_items = acc < max_items ? acc : max_items;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const int cnt = seg_start[e + 1] - seg_start[e];
        const int chunks = (cnt + 31) / 32;
        int w = item_base[e];
        for (int t = 0; t < tiles; ++t)                  // chunks of one (expert, tile) are neighbours: the weight tile stays in L2
            for (int c = 0; c < chunks; ++c, ++w)
                if (w < max_items) items[w] = MoeItem{e * n + t * 128, seg_start[e] + c * 32, min(32, cnt - c * 32), t * 128};
    }
}

// expert-sorted fp16 (K4 order) copy of the routed activation rows: xs[j] = x[row(sorted_pair[j])], row = pair / topk when x has one row
// per token.  Rows [pairs, pairs + 32) are zero-filled (a chunk's TMA box may run past the last routed row).
__global__ void __launch_bounds__(256)
moe_gather_kernel(const float* __restrict__ x, const uint32_t* __restrict__ sorted_pair, __half* __restrict__ xs, int pairs, int k, int per_token,
                  int topk) {
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x;
    __half* dst = xs + (int64_t)j * k;
    if (j >= pairs) {
        for (int i = threadIdx.x * 8; i < k; i += blockDim.x * 8) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint32_t pr = sorted_pair[j];
    const float* src = x + (int64_t)(per_token ? pr / (uint32_t)topk : pr) * k;
    for (int i = threadIdx.x * 4; i < k; i += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        const __half2 p0 = __halves2half2(from_f32<__half>(v.x), from_f32<__half>(v.z)), p1 = __halves2half2(from_f32<__half>(v.y), from_f32<__half>(v.w));
        *reinterpret_cast<uint2*>(dst + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
    }
}

// the same from 16-bit activations into fp16 in NATURAL order (moe_gemm_fp8)
template <typename T>
__global__ void __launch_bounds__(256)
moe_gather16_kernel(const T* __restrict__ x, const uint32_t* __restrict__ sorted_pair, __half* __restrict__ xs, int pairs, int k, int per_token, int topk) {
    pdl_wait();
    pdl_trigger();
    const int j = blockIdx.x;
    __half* dst = xs + (int64_t)j * k;
    if (j >= pairs) {
        for (int i = threadIdx.x * 8; i < k; i += blockDim.x * 8) *reinterpret_cast<uint4*>(dst + i) = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint32_t pr = sorted_pair[j];
    const T* src = x + (int64_t)(per_token ? pr / (uint32_t)topk : pr) * k;
    for (int i = threadIdx.x * 8; i < k; i += blockDim.x * 8) {
        const uint4 raw = *reinterpret_cast<const uint4*>(src + i);
        if constexpr (std::is_same<T, __half>::value) { *reinterpret_cast<uint4*>(dst + i) = raw; }
        else {
            const T* v = reinterpret_cast<const T*>(&raw);
            __half o[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = from_f32<__half>(to_f32(v[q]));
            *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(o);
        }
    }
}

// shape-generic FP8 fallback: one warp per (pair, output column); exact fp32 decode
template <typename T>
__global__ void __launch_bounds__(256)
moe_gemm_fp8_generic_kernel(const T* __restrict__ x, const uint8_t* __restrict__ w, const float* __restrict__ scale, const float* __restrict__ topk_w,
                            const uint32_t* __restrict__ sorted_pair, const uint32_t* __restrict__ sorted_expert, float* __restrict__ out, int pairs, int n, int k,
                            int by, int bx, int per_token, int topk) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.y, col = blockIdx.x * 8 + warp;
    if (j >= pairs || col >= n) return;
    const uint32_t pr = sorted_pair[j], e = sorted_expert[j];
    const T* xr = x + (int64_t)(per_token ? pr / (uint32_t)topk : pr) * k;
    const uint8_t* wr = w + ((int64_t)e * n + col) * k;
    const int sk = (k + bx - 1) / bx;
    const float* sr = scale + ((int64_t)e * ((n + by - 1) / by) + col / by) * sk;
    float acc = 0.f;
    for (int i = lane; i < k; i += 32) {
        const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)wr[i], __NV_E4M3);
        acc += __half2float(__half(h)) * sr[i / bx] * to_f32(xr[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(int64_t)pr * n + col] = acc * (topk_w ? topk_w[pr] : 1.f);
}

// shape-generic fallback (k not a multiple of 256, Q8_0, ...): one warp per (pair, output column chunk) -- correctness path
template <int kType>
__global__ void __launch_bounds__(256)
moe_gemm_generic_kernel(const float* __restrict__ x, const void* __restrict__ w_, const float* __restrict__ topk_w, const uint32_t* __restrict__ sorted_pair,
                        const uint32_t* __restrict__ sorted_expert, float* __restrict__ out, int pairs, int n, int k, int per_token, int topk) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.y;
    const int col = blockIdx.x * 8 + warp;
    if (j >= pairs || col >= n) return;
    const uint32_t pr = sorted_pair[j], e = sorted_expert[j];
    const float* xr = x + (int64_t)(per_token ? pr / (uint32_t)topk : pr) * k;
    float acc = 0.f;
    if constexpr (kType == B200_GGML_Q4_K) {
        const block_q4_K* wr = static_cast<const block_q4_K*>(w_) + ((int64_t)e * n + col) * (k / 256);
        for (int i = lane; i < k; i += 32) acc += to_f32(from_f32<__half>(xr[i])) * to_f32(from_f32<__half>(q4k_weight(wr + i / 256, i % 256)));
    } else if constexpr (kType == B200_GGML_Q6_K) {
        const block_q6_K* wr = static_cast<const block_q6_K*>(w_) + ((int64_t)e * n + col) * (k / 256);
        for (int i = lane; i < k; i += 32) acc += to_f32(from_f32<__half>(xr[i])) * to_f32(from_f32<__half>(q6k_weight(wr + i / 256, i % 256)));
    } else {
        const block_q8_0* wr = static_cast<const block_q8_0*>(w_) + ((int64_t)e * n + col) * (k / 32);
        for (int i = lane; i < k; i += 32) acc += to_f32(from_f32<__half>(xr[i])) * to_f32(from_f32<__half>(q8_0_weight(wr + i / 32, i % 32)));
    }
    acc = warp_sum(acc);
    if (lane == 0) out[(int64_t)pr * n + col] = acc * (topk_w ? topk_w[pr] : 1.f);
}

}  // namespace

}  // namespace b200

using namespace b200;

extern "C" {

void topk_softmax(const float* router_logits, float* topk_weights, uint32_t* topk_ids, int32_t num_tokens, int32_t num_experts, int32_t topk,
                  int64_t stream) {
    if (num_tokens == 0) return;
    B200_REQUIRE(router_logits && topk_weights && topk_ids, kErrBadArg, "topk_softmax: null pointer");
    B200_REQUIRE(num_experts >= 1 && num_experts <= 1024 && topk >= 1 && topk <= num_experts && topk <= 32, kErrUnsupported,
                 "topk_softmax: %d experts (<= 1024), top-%d (<= 32)", num_experts, topk);
    launch_pdl(topk_softmax_kernel, dim3(ceil_div(num_tokens, 4)), dim3(128), 0, as_stream(stream), router_logits, topk_weights, topk_ids, (int)num_tokens,
               (int)num_experts, (int)topk);
    count_launch();
    check_launch("topk_softmax");
}

void sort_expert_assignments(const uint32_t* topk_ids, uint32_t* expert_ids, uint32_t* sorted_token_ids, int32_t num_pairs, int32_t num_experts,
                             int64_t stream) {
    if (num_pairs == 0) return;
    B200_REQUIRE(topk_ids && expert_ids && sorted_token_ids, kErrBadArg, "sort_expert_assignments: null pointer");
    B200_REQUIRE(num_experts >= 1 && num_experts <= 1024 && num_pairs > 0, kErrUnsupported, "sort_expert_assignments: %d experts (<= 1024)", num_experts);
    launch_pdl(moe_sort_kernel, dim3(1), dim3(1024), 0, as_stream(stream), topk_ids, expert_ids, sorted_token_ids, (int)num_pairs, (int)num_experts);
    count_launch();
    check_launch("sort_expert_assignments");
}

size_t moe_gemm_workspace_bytes(int32_t num_pairs, int32_t n, int32_t k, int32_t num_experts) {
    const size_t max_items = ((size_t)num_pairs / 32 + (size_t)num_experts) * (size_t)((n + 127) / 128);
    return ((size_t)(num_pairs + 32) * (size_t)k * 2 + 255) / 256 * 256 + max_items * sizeof(MoeItem) + 256;
}

void moe_gemm_gguf(const float* x, const void* experts, const float* topk_weights, const uint32_t* sorted_token_ids, const uint32_t* expert_ids,
                   float* out, int32_t num_experts, int32_t topk, int32_t size_m, int32_t num_pairs, int32_t n, int32_t k, int32_t ggml_type,
                   int32_t is_prefill, void* workspace, size_t workspace_bytes, int64_t stream) {
    (void)is_prefill;
    if (num_pairs == 0 || n == 0) return;
    B200_REQUIRE(x && experts && sorted_token_ids && expert_ids && out, kErrBadArg, "moe_gemm_gguf: null pointer");
    B200_REQUIRE(num_experts >= 1 && num_experts <= 1024 && topk >= 1 && n > 0 && k > 0, kErrBadArg, "moe_gemm_gguf: bad sizes");
    B200_REQUIRE(size_m == num_pairs || (int64_t)size_m * topk == num_pairs, kErrBadArg,
                 "moe_gemm_gguf: x has %d rows, expected %d (one per pair) or %d (one per token)", size_m, num_pairs, num_pairs / topk);
    B200_REQUIRE(ggml_type == B200_GGML_Q4_K || ggml_type == B200_GGML_Q6_K || ggml_type == B200_GGML_Q8_0, kErrUnsupported, "moe_gemm_gguf: ggml type %d", ggml_type);
    B200_REQUIRE(k % (ggml_type == B200_GGML_Q8_0 ? 32 : 256) == 0, kErrBadArg, "moe_gemm_gguf: k = %d is not a multiple of the block size", k);
    cudaStream_t st = as_stream(stream);
    const int per_token = size_m != num_pairs;
    if (qmatmul_tc_moe_supported(n, k, ggml_type) && k % 8 == 0 && workspace && workspace_bytes >= moe_gemm_workspace_bytes(num_pairs, n, k, num_experts) &&
        ((uintptr_t)workspace & 255) == 0) {
        __half* xs = static_cast<__half*>(workspace);
        const size_t xs_bytes = ((size_t)(num_pairs + 32) * (size_t)k * 2 + 255) / 256 * 256;
        MoeItem* items = reinterpret_cast<MoeItem*>(static_cast<char*>(workspace) + xs_bytes);
        const int max_items = (int)(((size_t)num_pairs / 32 + (size_t)num_experts) * (size_t)((n + 127) / 128));
        int* num_items = reinterpret_cast<int*>(items + max_items);
        launch_pdl(moe_gather_kernel, dim3(num_pairs + 32), dim3(256), 0, st, x, sorted_token_ids, xs, (int)num_pairs, (int)k, per_token, (int)topk);
        launch_pdl(moe_items_kernel, dim3(1), dim3(256), 0, st, expert_ids, (int)num_pairs, (int)num_experts, (int)n, items, num_items, max_items);
        count_launch(2);
        if (!check_launch("moe_gemm_gguf")) return;
        qmatmul_tc_moe(xs, num_pairs + 32, experts, num_experts, out, n, n, k, ggml_type, items, num_items, max_items, sorted_token_ids, topk_weights, st);
        return;
    }
    const dim3 grid(ceil_div(n, 8), num_pairs);
    if (ggml_type == B200_GGML_Q4_K) moe_gemm_generic_kernel<B200_GGML_Q4_K><<<grid, 256, 0, st>>>(x, experts, topk_weights, sorted_token_ids, expert_ids, out, num_pairs, n, k, per_token, topk);
    else if (ggml_type == B200_GGML_Q6_K) moe_gemm_generic_kernel<B200_GGML_Q6_K><<<grid, 256, 0, st>>>(x, experts, topk_weights, sorted_token_ids, expert_ids, out, num_pairs, n, k, per_token, topk);
    else moe_gemm_generic_kernel<B200_GGML_Q8_0><<<grid, 256, 0, st>>>(x, experts, topk_weights, sorted_token_ids, expert_ids, out, num_pairs, n, k, per_token, topk);
    count_launch();
    check_launch("moe_gemm_gguf");
}

size_t moe_gemm_fp8_workspace_bytes(int32_t num_pairs, int32_t n, int32_t k, int32_t num_experts) {
    return moe_gemm_workspace_bytes(num_pairs, n, k, num_experts) + ((size_t)num_pairs * (size_t)n * 4 + 255) / 256 * 256 + 256;
}

// attention_rs::moe::moe_gemm_fp8 (call sites /root/reference/src/openai/models/layers/moe.rs:1447-1473): x [size_m, k] and out [num_pairs, n] of
// `dtype` (f16 / bf16); experts e4m3 [E, n, k]; scale f32 [E, ceil(n / by), ceil(k / bx)]; the rest as moe_gemm_gguf.
void moe_gemm_fp8(const void* x, const void* experts, const float* scale, const float* topk_weights, const uint32_t* sorted_token_ids,
                  const uint32_t* expert_ids, void* out, int32_t num_experts, int32_t topk, int32_t size_m, int32_t num_pairs, int32_t n, int32_t k,
                  int32_t block_y, int32_t block_x, int32_t dtype, int32_t is_prefill, void* workspace, size_t workspace_bytes, int64_t stream) {
    (void)is_prefill;
    if (num_pairs == 0 || n == 0) return;
    B200_REQUIRE(x && experts && scale && sorted_token_ids && expert_ids && out, kErrBadArg, "moe_gemm_fp8: null pointer");
    B200_REQUIRE(num_experts >= 1 && num_experts <= 1024 && topk >= 1 && n > 0 && k > 0 && block_y > 0 && block_x > 0, kErrBadArg, "moe_gemm_fp8: bad sizes");
    B200_REQUIRE(size_m == num_pairs || (int64_t)size_m * topk == num_pairs, kErrBadArg,
                 "moe_gemm_fp8: x has %d rows, expected %d (one per pair) or %d (one per token)", size_m, num_pairs, num_pairs / topk);
    B200_REQUIRE(dtype == B200_F16 || dtype == B200_BF16, kErrUnsupported, "moe_gemm_fp8: dtype %d (f16 / bf16)", dtype);
    B200_REQUIRE(workspace && ((uintptr_t)workspace & 255) == 0 && workspace_bytes >= moe_gemm_fp8_workspace_bytes(num_pairs, n, k, num_experts), kErrBadArg,
                 "moe_gemm_fp8: workspace of moe_gemm_fp8_workspace_bytes() bytes, 256-byte aligned, required");
    cudaStream_t st = as_stream(stream);
    const int per_token = size_m != num_pairs;
    char* ws = static_cast<char*>(workspace);
    const size_t xs_bytes = ((size_t)(num_pairs + 32) * (size_t)k * 2 + 255) / 256 * 256;
    const int max_items = (int)(((size_t)num_pairs / 32 + (size_t)num_experts) * (size_t)((n + 127) / 128));
    const size_t base = moe_gemm_workspace_bytes(num_pairs, n, k, num_experts);
    float* y32 = reinterpret_cast<float*>(ws + base);
    float* norm = reinterpret_cast<float*>(ws + base + ((size_t)num_pairs * (size_t)n * 4 + 255) / 256 * 256);
    if (qmatmul_tc_moe_fp8_supported(n, k, block_y, block_x) && k % 8 == 0 && (((uintptr_t)x | (uintptr_t)experts) & 15) == 0) {
        __half* xs = reinterpret_cast<__half*>(ws);
        MoeItem* items = reinterpret_cast<MoeItem*>(ws + xs_bytes);
        int* num_items = reinterpret_cast<int*>(items + max_items);
        if (dtype == B200_BF16) launch_pdl(moe_gather16_kernel<__nv_bfloat16>, dim3(num_pairs + 32), dim3(256), 0, st, (const __nv_bfloat16*)x, sorted_token_ids, xs, (int)num_pairs, (int)k, per_token, (int)topk);
        else launch_pdl(moe_gather16_kernel<__half>, dim3(num_pairs + 32), dim3(256), 0, st, (const __half*)x, sorted_token_ids, xs, (int)num_pairs, (int)k, per_token, (int)topk);
        launch_pdl(moe_items_kernel, dim3(1), dim3(256), 0, st, expert_ids, (int)num_pairs, (int)num_experts, (int)n, items, num_items, max_items);
        count_launch(2);
        if (!check_launch("moe_gemm_fp8")) return;
        qmatmul_tc_moe_fp8(xs, num_pairs + 32, experts, scale, num_experts, y32, n, n, k, block_y, block_x, items, num_items, max_items, sorted_token_ids,
                           topk_weights, norm, st);
    } else {
        const dim3 grid(ceil_div(n, 8), num_pairs);
        if (dtype == B200_BF16)
            moe_gemm_fp8_generic_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const uint8_t*)experts, scale, topk_weights, sorted_token_ids, expert_ids, y32, num_pairs, n, k, block_y, block_x, per_token, topk);
        else
            moe_gemm_fp8_generic_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, (const uint8_t*)experts, scale, topk_weights, sorted_token_ids, expert_ids, y32, num_pairs, n, k, block_y, block_x, per_token, topk);
        count_launch();
        if (!check_launch("moe_gemm_fp8")) return;
    }
    cast(y32, out, (int64_t)num_pairs * n, B200_F32, dtype, stream);
}

}  // extern "C"
