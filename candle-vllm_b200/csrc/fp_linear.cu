// fp_linear.cu -- weight-only low-precision float linears: block-scaled FP8 (e4m3), NVFP4 and MXFP4.
//
// Reference call sites (the kernels themselves live in attention-rs, not in the reference tree):
//   LnFp8::forward   -> attention_rs::fp8_linear::fp8_matmul      /root/reference/src/openai/models/linear.rs:1190-1221
//       weight e4m3 [N,K], weight_scale f32 [ceil(N/by), ceil(K/bx)] multiplies the tile (linear.rs:944-973)
//   LnNvfp4::forward -> attention_rs::nvfp4_linear::nvfp4_matmul  linear.rs:1913-1943
//       blocks u8 [N,K/2] (two e2m1 per byte, low nibble = even k), scales e4m3 [N,K/16], global_scale f32 (:1821-1853)
//   LnMxfp4::forward -> attention_rs::mxfp4_linear::mxfp4_matmul  linear.rs:1717-1757
//       blocks u8 [N,K/2], scales e8m0 [N,K/32] (weight = e2m1 * 2^(scale - 127))
// Contract here: out[m,n] = sum_k x[m,k] * w[n,k] (+ bias[n]) with the weight decoded exactly in fp32, fp32 accumulation,
// one rounding to the activation dtype at the end.  NVFP4's input_scale (activation quantisation of the reference's
// native-FP4 tensor-core path) does not apply to a weight-only product and is ignored.
//
// One warp per output row, lanes over K in 16-byte weight chunks, 8 activation rows per pass: a shape-generic SIMT kernel
// (any m / n / k the formats allow).  The m <= 64 decode shapes of block-FP8 run on the tcgen05 pipeline instead
// (qmatmul_tc.cu: F8Quarter, fp8_tc_run).  B200_FP8_GENERIC=1 forces the SIMT kernel (tests compare the two).
#include <cstdlib>

#include "qmatmul.cuh"

namespace b200 {

namespace {

constexpr int kRows = 8;      // warps per CTA = output rows per CTA
constexpr int kMT = 8;        // activation rows per pass

__device__ __forceinline__ float e4m3_to_f32(uint32_t b) {
    // exact decode (finite values; 0x7f / 0xff are NaN in e4m3fn): sign | 4-bit exponent (bias 7) | 3-bit mantissa
    const uint32_t e = (b >> 3) & 0xf, mnt = b & 7;
    float v;
    if (e == 0) v = (float)mnt * 0.001953125f;                                   // subnormal: m * 2^-9
    else if (e == 15 && mnt == 7) v = __int_as_float(0x7fc00000);
    else v = __int_as_float(((e + 120) << 23) | (mnt << 20));                     // 2^(e-7) * (1 + m/8)
    return (b & 0x80) ? -v : v;
}
__device__ __forceinline__ float e2m1_to_f32(uint32_t nib) {
    // {0, .5, 1, 1.5, 2, 3, 4, 6} with sign bit 3
    const uint32_t mag = nib & 7;
    const float v = mag < 2 ? 0.5f * (float)mag : __int_as_float((((mag >> 1) + 126) << 23) | ((mag & 1) << 22));
    return (nib & 8) ? -v : v;
}
__device__ __forceinline__ float e8m0_to_f32(uint32_t b) {
    return b == 0 ? __int_as_float(0x00400000) /* 2^-127 */ : (b == 255 ? __int_as_float(0x7fc00000) : __int_as_float(b << 23));
}

struct Fp8Dec {            // 16 weights per lane step
    const uint8_t* w; const float* scale; int k, by, bx, sk;     // sk = scale columns
    static constexpr int kPer = 16;
    __device__ __forceinline__ void load(int row, int k0, float (&wv)[16]) const {
        const uint4 raw = *reinterpret_cast<const uint4*>(w + (int64_t)row * k + k0);
        const uint32_t words[4] = {raw.x, raw.y, raw.z, raw.w};
        const float* srow = scale + (int64_t)(row / by) * sk;
#pragma unroll
        for (int j = 0; j < 16; ++j) wv[j] = e4m3_to_f32((words[j >> 2] >> (8 * (j & 3))) & 0xff) * srow[(k0 + j) / bx];
    }
};
struct Nvfp4Dec {          // 32 weights per lane step
    const uint8_t* w; const uint8_t* scales; float global; int k;
    static constexpr int kPer = 32;
    __device__ __forceinline__ void load(int row, int k0, float (&wv)[32]) const {
        const uint4 raw = *reinterpret_cast<const uint4*>(w + ((int64_t)row * k + k0) / 2);
        const uint32_t words[4] = {raw.x, raw.y, raw.z, raw.w};
        const uint8_t* srow = scales + (int64_t)row * (k / 16) + k0 / 16;
        const float s0 = e4m3_to_f32(srow[0]) * global, s1 = e4m3_to_f32(srow[1]) * global;
#pragma unroll
        for (int j = 0; j < 32; ++j) wv[j] = e2m1_to_f32((words[j >> 3] >> (4 * (j & 7))) & 0xf) * (j < 16 ? s0 : s1);
    }
};
struct Mxfp4Dec {
    const uint8_t* w; const uint8_t* scales; int k;
    static constexpr int kPer = 32;
    __device__ __forceinline__ void load(int row, int k0, float (&wv)[32]) const {
        const uint4 raw = *reinterpret_cast<const uint4*>(w + ((int64_t)row * k + k0) / 2);
        const uint32_t words[4] = {raw.x, raw.y, raw.z, raw.w};
        const float s = e8m0_to_f32(scales[(int64_t)row * (k / 32) + k0 / 32]);
#pragma unroll
        for (int j = 0; j < 32; ++j) wv[j] = e2m1_to_f32((words[j >> 3] >> (4 * (j & 7))) & 0xf) * s;
    }
};

template <typename Dec, typename T>
__global__ void __launch_bounds__(kRows * 32)
fp_linear_kernel(const T* __restrict__ x, const Dec dec, const T* __restrict__ bias, T* __restrict__ out, int m, int n, int k) {
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRows + warp;
    if (row >= n) return;
    const int m0 = blockIdx.y * kMT;
    constexpr int P = Dec::kPer;
    float acc[kMT];
#pragma unroll
    for (int i = 0; i < kMT; ++i) acc[i] = 0.f;
    for (int k0 = lane * P; k0 < k; k0 += 32 * P) {
        float wv[P];
        dec.load(row, k0, wv);
#pragma unroll
        for (int i = 0; i < kMT; ++i) {
            if (m0 + i < m) {
                const T* xr = x + (int64_t)(m0 + i) * k + k0;
#pragma unroll
                for (int j = 0; j < P; j += 8) {
                    const uint4 xv = *reinterpret_cast<const uint4*>(xr + j);
                    const T* xe = reinterpret_cast<const T*>(&xv);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[i] += to_f32(xe[e]) * wv[j + e];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kMT; ++i) {
        const float s = warp_sum(acc[i]);
        if (lane == 0 && m0 + i < m) out[(int64_t)(m0 + i) * n + row] = from_f32<T>(s + (bias ? to_f32(bias[row]) : 0.f));
    }
}

// W -> 16-bit [n, k] in the activation dtype (one rounding of the exactly decoded weight): the "dequantise once" half of the prefill path
template <typename Dec, typename T>
__global__ void __launch_bounds__(256)
fp_dequant_kernel(const Dec dec, T* __restrict__ out, int n, int k) {
    pdl_wait();
    pdl_trigger();
    constexpr int P = Dec::kPer;
    const int64_t per_row = k / P, total = (int64_t)n * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / per_row), k0 = (int)(i - (int64_t)row * per_row) * P;
        float wv[P];
        dec.load(row, k0, wv);
        T o[P];
#pragma unroll
        for (int j = 0; j < P; ++j) o[j] = from_f32<T>(wv[j]);
#pragma unroll
        for (int j = 0; j < P / 8; ++j) reinterpret_cast<uint4*>(out + (int64_t)row * k + k0)[j] = reinterpret_cast<const uint4*>(o)[j];
    }
}

// Rows at or above this count (prefill chunks) dequantise the weights once into the library scratch and run the dense tcgen05 GEMM
// (dense_gemm.cu) with the bias in its epilogue: every m-tile reuses the 16-bit copy instead of re-decoding the weights per 8 rows.
constexpr int kDenseRows = 512;

template <typename Dec>
bool prefill_dense(const Dec& dec, const void* x, const void* bias, void* out, int m, int n, int k, int dtype, cudaStream_t st, const char* who) {
    if (m < kDenseRows || k % Dec::kPer || k % 8 || n % 8) return false;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    void* w16 = get_scratch((size_t)n * k * 2 + 256, st);
    if (!w16) { if (cs == cudaStreamCaptureStatusNone) return false; b200_last_error(); return false; }     // capture with a cold scratch: SIMT path
    int64_t g = ((int64_t)n * (k / Dec::kPer) + 255) / 256;
    if (g > (int64_t)sm_count() * 16) g = (int64_t)sm_count() * 16;
    if (dtype == B200_BF16) launch_pdl(fp_dequant_kernel<Dec, __nv_bfloat16>, dim3((int)g), dim3(256), 0, st, dec, (__nv_bfloat16*)w16, n, k);
    else launch_pdl(fp_dequant_kernel<Dec, __half>, dim3((int)g), dim3(256), 0, st, dec, (__half*)w16, n, k);
    count_launch();
    if (!check_launch(who)) return true;
    dense_gemm_16(x, w16, bias, out, m, n, k, k, k, n, dtype, dtype, st);
    return true;
}

template <typename Dec>
void launch(const Dec& dec, const void* x, const void* bias, void* out, int m, int n, int k, int dtype, cudaStream_t st, const char* who) {
    if (prefill_dense(dec, x, bias, out, m, n, k, dtype, st, who)) return;
    const dim3 grid(ceil_div(n, kRows), ceil_div(m, kMT));
    if (dtype == B200_BF16)
        launch_pdl(fp_linear_kernel<Dec, __nv_bfloat16>, grid, dim3(kRows * 32), 0, st, (const __nv_bfloat16*)x, dec, (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, m, n, k);
    else
        launch_pdl(fp_linear_kernel<Dec, __half>, grid, dim3(kRows * 32), 0, st, (const __half*)x, dec, (const __half*)bias, (__half*)out, m, n, k);
    count_launch();
    check_launch(who);
}

// x (f16 / bf16) -> fp16 in "K8" order for the FP4 GEMM on the tcgen05 pipeline: within every aligned group of 8 along k the
// positions 0..7 hold k = 0,4,1,5,2,6,3,7 (qmatmul_tc.cu, F4Quarter)
template <typename T>
__global__ void cast_f16_k8_kernel(const T* __restrict__ x, __half* __restrict__ out, int64_t groups) {
    pdl_wait();
    pdl_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 raw = *reinterpret_cast<const uint4*>(x + i * 8);
        const T* v = reinterpret_cast<const T*>(&raw);
        __half o[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { o[2 * q] = __float2half_rn(to_f32(v[q])); o[2 * q + 1] = __float2half_rn(to_f32(v[4 + q])); }
        *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(o);
    }
}

// decode sizes of the FP4 linears on the tcgen05 pipeline (qmatmul_tc.cu), 64 rows per pass.  Scratch = fp16 K8 copy of x + fp32
// partial-sum slabs + the post-scale pair.  false: shape / alignment not supported there (the caller falls back to the SIMT kernel)
bool fp4_tc_route(bool mx, const void* x, const void* blocks, const void* scales, float global_scale, const void* bias, void* out, int m, int n, int k,
                  int dtype, cudaStream_t st) {
    static const bool force_generic = [] { const char* e = getenv("B200_FP4_GENERIC"); return e && atoi(e) != 0; }();
    if (force_generic || !fp4_tc_supported(m > 64 ? 64 : m, n, k) || (((uintptr_t)out) & 7) || (((uintptr_t)scales) & 3)) return false;
    const int mc = m < 64 ? m : 64;
    const size_t x_bytes = ((size_t)mc * k * 2 + 255) & ~(size_t)255;
    const size_t slab_bytes = (size_t)wq16_slabs(n, k) * mc * n * 4;
    char* xs = static_cast<char*>(get_scratch(x_bytes + slab_bytes + 256, st));
    if (!xs) { b200_last_error(); return false; }
    for (int m0 = 0; m0 < m; m0 += 64) {
        const int mm = m - m0 < 64 ? m - m0 : 64;
        const int64_t groups = (int64_t)mm * k / 8;
        int64_t g = (groups + 255) / 256;
        if (g > (int64_t)sm_count() * 8) g = (int64_t)sm_count() * 8;
        const char* xr = static_cast<const char*>(x) + (size_t)m0 * k * 2;
        if (dtype == B200_BF16) launch_pdl(cast_f16_k8_kernel<__nv_bfloat16>, dim3((int)g), dim3(256), 0, st, (const __nv_bfloat16*)xr, (__half*)xs, groups);
        else launch_pdl(cast_f16_k8_kernel<__half>, dim3((int)g), dim3(256), 0, st, (const __half*)xr, (__half*)xs, groups);
        count_launch();
        fp4_tc_run(mx, xs, blocks, scales, global_scale, bias, static_cast<char*>(out) + (size_t)m0 * n * 2, dtype, mm, n, k,
                   reinterpret_cast<float*>(xs + x_bytes), reinterpret_cast<float*>(xs + x_bytes + slab_bytes), st);
    }
    return true;
}

bool common_check(const char* who, const void* x, const void* w, const void* s, const void* out, int m, int n, int k, int dtype, int kmul) {
    if (!x || !w || !s || !out) { set_error(kErrBadArg, "%s: null pointer", who); return false; }
    if (m <= 0 || n <= 0 || k <= 0 || k % kmul) { set_error(kErrBadArg, "%s: bad sizes m=%d n=%d k=%d (k %% %d)", who, m, n, k, kmul); return false; }
    if (dtype != B200_F16 && dtype != B200_BF16) { set_error(kErrUnsupported, "%s: dtype %d (f16 / bf16 only)", who, dtype); return false; }
    if (((uintptr_t)x | (uintptr_t)w) & 15) { set_error(kErrBadArg, "%s: x and weights must be 16-byte aligned", who); return false; }
    return true;
}

}  // namespace

}  // namespace b200

using namespace b200;

extern "C" {

void fp8_matmul(const void* x, const void* weight, const float* weight_scale, const void* bias, void* out, int32_t m, int32_t n,
                int32_t k, int32_t block_y, int32_t block_x, int32_t dtype, int64_t stream) {
    if (m == 0 || n == 0) return;
    if (!common_check("fp8_matmul", x, weight, weight_scale, out, m, n, k, dtype, 16)) return;
    B200_REQUIRE(block_y > 0 && block_x > 0, kErrBadArg, "fp8_matmul: block sizes [%d, %d]", block_y, block_x);
    static const bool force_generic = [] { const char* e = getenv("B200_FP8_GENERIC"); return e && atoi(e) != 0; }();
    if (!force_generic && m >= kDenseRows) {
        const Fp8Dec dec{static_cast<const uint8_t*>(weight), weight_scale, k, block_y, block_x, (k + block_x - 1) / block_x};
        if (prefill_dense(dec, x, bias, out, m, n, k, dtype, as_stream(stream), "fp8_matmul")) return;
    }
    if (!force_generic && fp8_tc_supported(m > 64 ? 64 : m, n, k, block_y, block_x) && (((uintptr_t)out | (uintptr_t)weight_scale) & 7) == 0) {
        // tcgen05 pipeline (qmatmul_tc.cu), 64 rows per pass.  Scratch = fp16 copy of x (bf16 input) + fp32 partial-sum slabs + norm.
        cudaStream_t st = as_stream(stream);
        const int mc = m < 64 ? m : 64;
        const size_t x_bytes = ((size_t)mc * k * 2 + 255) & ~(size_t)255;
        const size_t slab_bytes = (size_t)wq16_slabs(n, k) * mc * n * 4;
        char* xs = static_cast<char*>(get_scratch(x_bytes + slab_bytes + 256, st));
        if (!xs) return;
        for (int m0 = 0; m0 < m; m0 += 64) {
            const int mm = m - m0 < 64 ? m - m0 : 64;
            const void* x16 = static_cast<const char*>(x) + (size_t)m0 * k * 2;                     // f16 activations feed the tensor map directly
            if (dtype != B200_F16) { cast(x16, xs, (int64_t)mm * k, dtype, B200_F16, stream); x16 = xs; }
            fp8_tc_run(x16, weight, weight_scale, bias, static_cast<char*>(out) + (size_t)m0 * n * 2, dtype, mm, n, k, block_y, block_x,
                       reinterpret_cast<float*>(xs + x_bytes), reinterpret_cast<float*>(xs + x_bytes + slab_bytes), st);
        }
        return;
    }
    const Fp8Dec dec{static_cast<const uint8_t*>(weight), weight_scale, k, block_y, block_x, (k + block_x - 1) / block_x};
    launch(dec, x, bias, out, m, n, k, dtype, as_stream(stream), "fp8_matmul");
}

void nvfp4_matmul(const void* x, const void* blocks, const void* scales, float global_scale, float input_scale, const void* bias,
                  void* out, int32_t m, int32_t n, int32_t k, int32_t dtype, int64_t stream) {
    (void)input_scale;      // activation-quantisation scale of the reference's native FP4 path: not used by a weight-only product
    if (m == 0 || n == 0) return;
    if (!common_check("nvfp4_matmul", x, blocks, scales, out, m, n, k, dtype, 32)) return;
    const Nvfp4Dec dec{static_cast<const uint8_t*>(blocks), static_cast<const uint8_t*>(scales), global_scale, k};
    if (m < kDenseRows && fp4_tc_route(false, x, blocks, scales, global_scale, bias, out, m, n, k, dtype, as_stream(stream))) return;
    launch(dec, x, bias, out, m, n, k, dtype, as_stream(stream), "nvfp4_matmul");
}

void mxfp4_matmul(const void* x, const void* blocks, const void* scales, const void* bias, void* out, int32_t m, int32_t n,
                  int32_t k, int32_t dtype, int64_t stream) {
    if (m == 0 || n == 0) return;
    if (!common_check("mxfp4_matmul", x, blocks, scales, out, m, n, k, dtype, 32)) return;
    const Mxfp4Dec dec{static_cast<const uint8_t*>(blocks), static_cast<const uint8_t*>(scales), k};
    if (m < kDenseRows && fp4_tc_route(true, x, blocks, scales, 1.f, bias, out, m, n, k, dtype, as_stream(stream))) return;
    launch(dec, x, bias, out, m, n, k, dtype, as_stream(stream), "mxfp4_matmul");
}

}  // extern "C"
