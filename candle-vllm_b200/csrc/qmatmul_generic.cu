// qmatmul_generic.cu -- shape-generic quantised mat-mul and dequantise for GGML Q4_K / Q6_K / Q8_0.
//
// y[m,n] = x[m,k] . dequant(W[n,k])^T.  One warp per weight row, lanes over k (8 consecutive
// weights per lane per 256-wide super-block), 8 activation rows per pass, fp32 accumulate.  This is
// the catch-all / bring-up path (odd shapes, large m); the tcgen05 kernel in qmatmul_tc.cu is the
// decode hot path.  Activations are rounded to fp16 first so both paths share one numerical
// contract (see include/b200_backend.h).
//
// Reference call sites: QMatMul::forward /root/reference/src/openai/models/linear.rs:765-806;
// QTensor::dequantize linear.rs:808-842.
#include "qmatmul.cuh"

namespace b200 {

template <int kType> struct QType;
template <> struct QType<B200_GGML_Q4_K> { using Block = block_q4_K; static constexpr int kElems = 256;
    static __device__ __forceinline__ float w(const Block* b, int i) { return q4k_weight(b, i); } };
template <> struct QType<B200_GGML_Q6_K> { using Block = block_q6_K; static constexpr int kElems = 256;
    static __device__ __forceinline__ float w(const Block* b, int i) { return q6k_weight(b, i); } };
template <> struct QType<B200_GGML_Q8_0> { using Block = block_q8_0; static constexpr int kElems = 32;
    static __device__ __forceinline__ float w(const Block* b, int i) { return q8_0_weight(b, i); } };

constexpr int kRowsPerCta = 8;   // warps
constexpr int kMTile = 8;

template <int kType, typename TX>
__global__ void __launch_bounds__(kRowsPerCta * 32)
qmatmul_generic_kernel(const TX* __restrict__ x, const void* __restrict__ w_, float* __restrict__ y,
                       int64_t ldy, int m, int n, int k, int accumulate) {
    using Q = QType<kType>;
    using Block = typename Q::Block;
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowsPerCta + warp;
    if (row >= n) return;
    const int m0 = blockIdx.y * kMTile;
    const int blocks_per_row = k / Q::kElems;
    const Block* wrow = static_cast<const Block*>(w_) + (int64_t)row * blocks_per_row;
    float acc[kMTile];
#pragma unroll
    for (int i = 0; i < kMTile; ++i) acc[i] = 0.f;
    for (int k0 = lane * 8; k0 < k; k0 += 256) {
        float wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kk = k0 + j;
            wv[j] = Q::w(wrow + kk / Q::kElems, kk % Q::kElems);
        }
#pragma unroll
        for (int i = 0; i < kMTile; ++i) {
            if (m0 + i < m) {
                const TX* xr = x + (int64_t)(m0 + i) * k + k0;
                // fp16 activations arrive in K4 order (middle two of every 4 swapped); f32 ones are natural
                constexpr bool k4 = sizeof(TX) == 2;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i] += to_f32(from_f32<__half>(to_f32(xr[k4 ? (int)k4_index(j) : j]))) * wv[j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kMTile; ++i) {
        const float s = warp_sum(acc[i]);
        if (lane == 0 && m0 + i < m) {
            float* o = y + (int64_t)(m0 + i) * ldy + row;
            if (accumulate) atomicAdd(o, s); else *o = s;
        }
    }
}

template <int kType>
__global__ void dequantize_kernel(const void* __restrict__ w_, float* __restrict__ out, int64_t total) {
    using Q = QType<kType>;
    const typename Q::Block* w = static_cast<const typename Q::Block*>(w_);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = Q::w(w + i / Q::kElems, (int)(i % Q::kElems));
}

// W -> fp16 [n, k] (natural k order), 8 weights per thread and one 16-byte store: the "dequantise once" half of the large-m path
// (dense_gemm.cu reuses the result for every m-tile).  One rounding to fp16, like the decode kernel's dequant-into-TMEM.
template <int kType>
__global__ void __launch_bounds__(256)
dequantize_f16_kernel(const void* __restrict__ w_, __half* __restrict__ out, int64_t total8) {
    using Q = QType<kType>;
    const typename Q::Block* w = static_cast<const typename Q::Block*>(w_);
    pdl_wait();
    pdl_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * 8;
        const typename Q::Block* b = w + e / Q::kElems;
        const int o = (int)(e % Q::kElems);
        __half h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = from_f32<__half>(Q::w(b, o + j));
        *reinterpret_cast<uint4*>(out + e) = *reinterpret_cast<const uint4*>(h);
    }
}

bool dequantize_f16(const void* w, void* out_f16, int64_t n, int64_t k, int ggml_type, cudaStream_t st) {
    const int64_t total8 = n * k / 8;
    int64_t g = (total8 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    switch (ggml_type) {
        case B200_GGML_Q4_K: launch_pdl(dequantize_f16_kernel<B200_GGML_Q4_K>, dim3((int)g), dim3(256), 0, st, w, (__half*)out_f16, total8); break;
        case B200_GGML_Q6_K: launch_pdl(dequantize_f16_kernel<B200_GGML_Q6_K>, dim3((int)g), dim3(256), 0, st, w, (__half*)out_f16, total8); break;
        case B200_GGML_Q8_0: launch_pdl(dequantize_f16_kernel<B200_GGML_Q8_0>, dim3((int)g), dim3(256), 0, st, w, (__half*)out_f16, total8); break;
        default: set_error(kErrUnsupported, "dequantize: ggml type %d unsupported", ggml_type); return false;
    }
    count_launch();
    return check_launch("dequantize_f16");
}

template <int kType>
static void launch(const void* x, bool f16, const void* w, float* y, int64_t ldy, int m, int n, int k, int acc, cudaStream_t st) {
    dim3 grid(ceil_div(n, kRowsPerCta), ceil_div(m, kMTile));
    if (f16) qmatmul_generic_kernel<kType, __half><<<grid, kRowsPerCta * 32, 0, st>>>((const __half*)x, w, y, ldy, m, n, k, acc);
    else qmatmul_generic_kernel<kType, float><<<grid, kRowsPerCta * 32, 0, st>>>((const float*)x, w, y, ldy, m, n, k, acc);
    count_launch();
}

void qmatmul_generic(const void* x, bool x_is_f16, const void* w, float* y, int64_t ldy, int m, int n, int k,
                     int ggml_type, int accumulate, cudaStream_t st) {
    switch (ggml_type) {
        case B200_GGML_Q4_K: launch<B200_GGML_Q4_K>(x, x_is_f16, w, y, ldy, m, n, k, accumulate, st); break;
        case B200_GGML_Q6_K: launch<B200_GGML_Q6_K>(x, x_is_f16, w, y, ldy, m, n, k, accumulate, st); break;
        case B200_GGML_Q8_0: launch<B200_GGML_Q8_0>(x, x_is_f16, w, y, ldy, m, n, k, accumulate, st); break;
        default: set_error(kErrUnsupported, "qmatmul: ggml type %d unsupported", ggml_type); return;
    }
    check_launch("qmatmul_generic");
}

// rows beyond the tensor-core kernel's 64 (prefill chunks) go through it 64 at a time: every pass re-streams the weights, which at
// >= 64 rows per pass is within ~3x of a dense GEMM and two orders faster than the SIMT fallback
bool qmatmul_tc_usable(int m, int n, int k, int ggml_type) { return m >= 1 && qmatmul_tc_supported(m > 64 ? 64 : m, n, k, ggml_type); }

void qmatmul_dispatch(const void* x_f16, const void* w, float* y, int64_t ldy, int m, int n, int k,
                      int ggml_type, int accumulate, cudaStream_t st) {
    if (qmatmul_tc_usable(m, n, k, ggml_type)) {
        for (int m0 = 0; m0 < m; m0 += 64)
            qmatmul_tc(static_cast<const __half*>(x_f16) + (int64_t)m0 * k, w, y + (int64_t)m0 * ldy, ldy, m - m0 < 64 ? m - m0 : 64, n, k, ggml_type,
                       accumulate, st);
    } else {
        qmatmul_generic(x_f16, true, w, y, ldy, m, n, k, ggml_type, accumulate, st);
    }
}

static bool can_fuse(int nseg, const int* types, const int* n, int m, int k) {
    bool fuse = nseg >= 1 && nseg <= 3;
    for (int i = 0; i < nseg && fuse; ++i) fuse = types[i] == types[0] && qmatmul_tc_supported(m, n[i], k, types[i]);
    return fuse;
}

void qmatmul_dispatch_multi(const void* x_f16, int nseg, const void* const* w, const int* types, float* const* y, const int* n,
                            int64_t ldy, int m, int k, int accumulate, cudaStream_t st) {
    if (can_fuse(nseg, types, n, m, k)) { qmatmul_tc_multi(x_f16, nseg, w, y, n, ldy, m, k, types[0], accumulate, 0, 0, st); return; }
    for (int i = 0; i < nseg; ++i) qmatmul_dispatch(x_f16, w[i], y[i], ldy, m, n[i], k, types[i], accumulate, st);
}

int qmatmul_slabs_needed(int nseg, const int* n, const int* types, int m, int k) {
    if (!can_fuse(nseg, types, n, m, k)) return 1;
    int64_t tiles = 0;
    for (int i = 0; i < nseg; ++i) tiles += (n[i] + 127) / 128;
    return qmatmul_tc_slab_count(tiles, k / 256);
}

int qmatmul_dispatch_slabs(const void* x_f16, int nseg, const void* const* w, const int* types, float* const* y, const int* n,
                           int64_t ldy, int m, int k, int slabs_avail, int64_t slab_stride, cudaStream_t st) {
    if (can_fuse(nseg, types, n, m, k)) return qmatmul_tc_multi(x_f16, nseg, w, y, n, ldy, m, k, types[0], 0, slabs_avail, slab_stride, st);
    for (int i = 0; i < nseg; ++i) qmatmul_generic(x_f16, true, w[i], y[i], ldy, m, n[i], k, types[i], 0, st);   // whole product -> slab 0
    return 1;
}

}  // namespace b200

using namespace b200;

static bool qmm_check(const char* who, const void* x, const void* w, const float* y, int m, int n, int k, int t) {
    if (!x || !w || !y) { set_error(kErrBadArg, "%s: null pointer", who); return false; }
    if (m <= 0 || n <= 0 || k <= 0) { set_error(kErrBadArg, "%s: bad sizes m=%d n=%d k=%d", who, m, n, k); return false; }
    const int be = t == B200_GGML_Q8_0 ? 32 : 256;
    if (t != B200_GGML_Q4_K && t != B200_GGML_Q6_K && t != B200_GGML_Q8_0) { set_error(kErrUnsupported, "%s: ggml type %d unsupported", who, t); return false; }
    if (k % be) { set_error(kErrBadArg, "%s: k=%d not a multiple of the block size %d", who, k, be); return false; }
    return true;
}

extern "C" {

// Rows at or above this count take the "dequantise once + dense tensor-core GEMM" path: the 16-bit copy of W costs ~4.6 bytes of
// traffic per weight (0.56 read, 2 written, 2 read back), a 64-row pass of the decode kernel 0.56 -- break-even at ~8 passes.
static constexpr int kDenseRows = 512;

size_t qmatmul_workspace_bytes(int32_t m, int32_t n, int32_t k) {
    size_t b = (size_t)(m > 0 ? m : 0) * (size_t)(k > 0 ? k : 0) * 2 + 256;    // fp16 copy of the activations
    if (m >= kDenseRows) b += (size_t)(n > 0 ? n : 0) * (size_t)(k > 0 ? k : 0) * 2 + 256;   // + fp16 copy of the weights
    return b;
}

void qmatmul_f16act(const void* x_f16, const void* w, float* y, int32_t m, int32_t n, int32_t k,
                    int32_t ggml_type, int32_t accumulate, int64_t stream) {
    if (m == 0 || n == 0) return;
    if (!qmm_check("qmatmul_f16act", x_f16, w, y, m, n, k, ggml_type)) return;
    if (!accumulate && qmatmul_tc_usable(m, n, k, ggml_type) && qmatmul_tc_needs_zeroed_output(n, k))
        cudaMemsetAsync(y, 0, (size_t)m * n * sizeof(float), as_stream(stream));
    qmatmul_dispatch(x_f16, w, y, n, m, n, k, ggml_type, accumulate, as_stream(stream));
}

int32_t qmatmul_slab_count(int32_t m, int32_t n, int32_t k, int32_t ggml_type) {
    if (m <= 0 || n <= 0 || k <= 0) return 0;
    return qmatmul_slabs_needed(1, &n, &ggml_type, m, k);
}

int32_t qmatmul_f16act_slabs(const void* x_f16, const void* w, float* y_slabs, int32_t slabs_avail, int32_t m, int32_t n,
                             int32_t k, int32_t ggml_type, int64_t stream) {
    if (m == 0 || n == 0) return 0;
    if (!qmm_check("qmatmul_f16act_slabs", x_f16, w, y_slabs, m, n, k, ggml_type)) return 0;
    if (slabs_avail < 1) { set_error(kErrBadArg, "qmatmul_f16act_slabs: slabs_avail = %d", slabs_avail); return 0; }
    return qmatmul_dispatch_slabs(x_f16, 1, &w, &ggml_type, &y_slabs, &n, n, m, k, slabs_avail, (int64_t)m * n, as_stream(stream));
}

void qmatmul_f32(const float* x, const void* w, float* y, int32_t m, int32_t n, int32_t k,
                 int32_t ggml_type, int32_t accumulate, void* workspace, size_t workspace_bytes, int64_t stream) {
    if (m == 0 || n == 0) return;
    if (!qmm_check("qmatmul_f32", x, w, y, m, n, k, ggml_type)) return;
    if (m >= kDenseRows && !accumulate && k % 8 == 0 && ((uintptr_t)workspace & 15) == 0 && workspace &&
        workspace_bytes >= qmatmul_workspace_bytes(m, n, k)) {
        // prefill chunk: weights -> fp16 once, activations -> fp16, dense tcgen05 GEMM with fp32 output (dense_gemm.cu)
        char* xs = static_cast<char*>(workspace);
        char* ws = xs + (((size_t)m * k * 2 + 255) & ~(size_t)255);
        cast(x, xs, (int64_t)m * k, B200_F32, B200_F16, stream);
        if (!dequantize_f16(w, ws, n, k, ggml_type, as_stream(stream))) return;
        dense_gemm_16(xs, ws, nullptr, y, m, n, k, k, k, n, B200_F16, B200_F32, as_stream(stream));
        return;
    }
    if (qmatmul_tc_usable(m, n, k, ggml_type)) {
        B200_REQUIRE(workspace && workspace_bytes >= qmatmul_workspace_bytes(m, n, k), kErrBadArg,
                     "qmatmul_f32: workspace too small (%zu < %zu)", workspace_bytes, qmatmul_workspace_bytes(m, n, k));
        cast(x, workspace, (int64_t)m * k, B200_F32, B200_F16_K4, stream);
        if (!accumulate && qmatmul_tc_needs_zeroed_output(n, k)) cudaMemsetAsync(y, 0, (size_t)m * n * sizeof(float), as_stream(stream));
        qmatmul_dispatch(workspace, w, y, n, m, n, k, ggml_type, accumulate, as_stream(stream));      // 64 rows per tensor-core pass
    } else {
        qmatmul_generic(x, false, w, y, n, m, n, k, ggml_type, accumulate, as_stream(stream));
    }
}

void dequantize_f32(const void* w, float* out, int64_t n, int64_t k, int32_t ggml_type, int64_t stream) {
    if (n == 0 || k == 0) return;
    B200_REQUIRE(w && out && n > 0 && k > 0, kErrBadArg, "dequantize: bad arguments");
    const int64_t total = n * k;
    int64_t g = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    cudaStream_t st = as_stream(stream);
    switch (ggml_type) {
        case B200_GGML_Q4_K: dequantize_kernel<B200_GGML_Q4_K><<<(int)g, 256, 0, st>>>(w, out, total); break;
        case B200_GGML_Q6_K: dequantize_kernel<B200_GGML_Q6_K><<<(int)g, 256, 0, st>>>(w, out, total); break;
        case B200_GGML_Q8_0: dequantize_kernel<B200_GGML_Q8_0><<<(int)g, 256, 0, st>>>(w, out, total); break;
        default: set_error(kErrUnsupported, "dequantize: ggml type %d unsupported", ggml_type); return;
    }
    count_launch();
    check_launch("dequantize");
}

}  // extern "C"
