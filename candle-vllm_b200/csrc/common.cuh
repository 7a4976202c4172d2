// common.cuh -- shared helpers for libb200backend (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200_backend.h"

namespace b200 {

// ---- per-thread error record (b200_last_error) ---------------------------------------------
void set_error(int code, const char* fmt, ...);
bool check_launch(const char* what);   // records cudaGetLastError() if any; returns true when ok
void count_launch(int n = 1);          // library-wide launch counter (gpu_launches evidence)
int  sm_count();                        // SM count of the CURRENT device (cached per device)
// cudaFuncAttributeMaxDynamicSharedMemorySize, applied once per (device, kernel): the attribute is per device, so a
// per-process `static bool` would leave every device but the first without it (threaded.rs runs one thread per rank in
// one process, /root/reference/src/openai/pipelines/threaded.rs:33-134)
void ensure_dynamic_smem(const void* kernel, int bytes);

#define B200_REQUIRE(cond, code, ...)                  \
    do {                                               \
        if (!(cond)) {                                 \
            ::b200::set_error((code), __VA_ARGS__);    \
            return;                                    \
        }                                              \
    } while (0)

enum ErrorCode { kErrBadArg = 1, kErrUnsupported = 2, kErrCuda = 3, kErrNoDevice = 4 };

static inline cudaStream_t as_stream(int64_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel on the decode path is launched with programmatic stream serialization allowed: its grid may start
// (block scheduling, barrier init, TMEM alloc, even weight prefetch) while the previous kernel drains, and it
// calls pdl_wait() before touching anything the previous kernel produced.  griddepcontrol.wait returns only when
// the prerequisite grid has fully completed and flushed, so triggering early is always safe.  B200_PDL=0 disables.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- dtype conversion -------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) {
    // saturating (no inf): activations fed to fp16 tensor-core GEMMs must stay finite
    return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
}
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint8_t f32_to_e4m3(float v) {
    return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float e4m3_to_f32(uint8_t b) {
    __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3);
    return __half2float(__half(h));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// position of natural index k in the K4 activation order (swap the middle two of every 4; self-inverse)
__host__ __device__ __forceinline__ int64_t k4_index(int64_t k) { return (k & ~3ll) | ((0xD8 >> ((k & 3) * 2)) & 3); }

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- GGML block formats (GGUF spec) -----------------------------------
struct __align__(2) block_q4_K { __half d, dmin; uint8_t scales[12]; uint8_t qs[128]; };      // 144 B
struct __align__(2) block_q6_K { uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; __half d; }; // 210 B
struct __align__(2) block_q8_0 { __half d; int8_t qs[32]; };                                   // 34 B
static_assert(sizeof(block_q4_K) == 144 && sizeof(block_q6_K) == 210 && sizeof(block_q8_0) == 34, "ggml blocks");

__device__ __forceinline__ void q4k_scale_min(int j, const uint8_t* q, int& sc, int& m) {
    if (j < 4) { sc = q[j] & 63; m = q[j + 4] & 63; }
    else { sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

// weight element i (0..255) of a Q4_K block: w = d*sc*q - dmin*m (fp32, ggml operation order)
__device__ __forceinline__ float q4k_weight(const block_q4_K* b, int i) {
    const int j = i >> 5, c = i >> 6, l = i & 31;
    int sc, m; q4k_scale_min(j, b->scales, sc, m);
    const uint8_t byte = b->qs[c * 32 + l];
    const int q = (j & 1) ? (byte >> 4) : (byte & 0xF);
    const float d1 = __half2float(b->d) * (float)sc, m1 = __half2float(b->dmin) * (float)m;
    return d1 * (float)q - m1;
}
__device__ __forceinline__ float q6k_weight(const block_q6_K* b, int i) {
    const int half = i >> 7, r = i & 127, g = r >> 5, l = r & 31;
    const uint8_t* ql = b->ql + 64 * half; const uint8_t* qh = b->qh + 32 * half;
    const int lo = (g & 1) ? ql[l + 32] : ql[l];
    const int nib = (g >= 2) ? (lo >> 4) : (lo & 0xF);
    const int hi = (qh[l] >> (2 * g)) & 3;
    const int q = (nib | (hi << 4)) - 32;
    return (__half2float(b->d) * (float)b->scales[i >> 4]) * (float)q;
}
__device__ __forceinline__ float q8_0_weight(const block_q8_0* b, int i) {
    return __half2float(b->d) * (float)b->qs[i];
}

}  // namespace b200
