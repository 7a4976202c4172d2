// qmatmul.cuh -- internal interfaces between the quantised-matmul translation units.
#pragma once
#include "common.cuh"

namespace b200 {

// shape-generic SIMT path (qmatmul_generic.cu): any m, any n, k % block == 0
void qmatmul_generic(const void* x, bool x_is_f16, const void* w, float* y, int64_t ldy, int m, int n, int k,
                     int ggml_type, int accumulate, cudaStream_t st);

// tcgen05 path (qmatmul_tc.cu): fp16 activations [m,k], m <= 32 * n_mtiles; returns false when the shape is not covered
bool qmatmul_tc_supported(int m, int n, int k, int ggml_type);
// same, with m > 64 served 64 rows per pass (qmatmul_dispatch)
bool qmatmul_tc_usable(int m, int n, int k, int ggml_type);
// true when some output tile is split across CTAs: y must hold the addend (zeros for a plain product)
bool qmatmul_tc_needs_zeroed_output(int n, int k);

// elementwise.cu internals used by the engine
void rope_and_cache_impl(float* qkv, void* q_out, void* key_cache, void* value_cache, const float* cos_t, const float* sin_t,
                         const int64_t* positions, const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads,
                         int32_t num_kv_heads, int32_t head_dim, int32_t interleaved, int32_t dtype, int32_t cache_dtype,
                         bool zero_src, int64_t stream);
void silu_mul_zero_src(float* gate, float* up, void* out_f16_k4, int64_t numel, int64_t stream);
void silu_mul_zero_src_fmt(float* gate, float* up, void* out, int64_t numel, int fmt, int64_t stream);
void qmatmul_tc(const void* x_f16, const void* w, float* y, int64_t ldy, int m, int n, int k, int ggml_type,
                int accumulate, cudaStream_t st);

// several weight matrices (same type / k) on one activation in one launch.  slabs_avail > 0 selects slab mode: split tiles
// write partial sums to y[i] + s * slab_stride, s < (return value); the caller's consumer adds the slabs up.
int qmatmul_tc_multi(const void* x_f16, int nseg, const void* const* w, float* const* y, const int* n, int64_t ldy, int m, int k,
                     int ggml_type, int accumulate, int slabs_avail, int64_t slab_stride, cudaStream_t st);
int qmatmul_tc_slab_count(int64_t n_tiles, int nsb);
// legacy (accumulating) form; falls back to separate launches
void qmatmul_dispatch_multi(const void* x_f16, int nseg, const void* const* w, const int* types, float* const* y, const int* n,
                            int64_t ldy, int m, int k, int accumulate, cudaStream_t st);
// slab form: returns how many slabs hold partial sums (1 when the product went to slab 0 whole); never reads y
int qmatmul_dispatch_slabs(const void* x_f16, int nseg, const void* const* w, const int* types, float* const* y, const int* n,
                           int64_t ldy, int m, int k, int slabs_avail, int64_t slab_stride, cudaStream_t st);
// slabs qmatmul_dispatch_slabs may need for these shapes
int qmatmul_slabs_needed(int nseg, const int* n, const int* types, int m, int k);

// 16-bit-output weight-only GEMMs on the tcgen05 pipeline: fp32 partial-sum slabs ([wq16_slabs][m][n] f32, caller scratch)
// + a finishing pass.  m <= 64, k % 256 == 0.
int wq16_slabs(int n, int k);
// int4 (GPTQ symmetric / AWQ with zero points; repacked by gptq_repack / awq_repack) x fp16 activations in K4 order
void marlin_tc(const void* x_f16_k4, const void* w, const void* scales, const void* qzeros /* null: symmetric */, void* out, int out_dtype,
               int m, int n, int k, int group_size, float* slabs, cudaStream_t st);
// L2 prefetch hint: the weights the NEXT GEMM of this thread's chain will stream (consumed by the next tcgen05 GEMM launch)
void qmatmul_tc_prefetch_next(int n, const void* const* ptrs, const size_t* bytes);
// decode-engine form: up to 3 int4 matrices sharing the activations in one launch, f32 output, split tiles red.add into y
// (accumulate = 0: y zeroed by the caller)
void marlin_tc_f32_multi(const void* x_f16_k4, int nseg, const void* const* w, const void* const* scales, int scale_bf16, const void* const* qzeros,
                         float* const* y, const int* n, int64_t ldy, int m, int k, int group_size, int accumulate, cudaStream_t st);
// e4m3 [n,k] with f32 scale per [by, bx] tile x fp16 activations in natural order
bool fp8_tc_supported(int m, int n, int k, int by, int bx);
void fp8_tc_run(const void* x_f16, const void* w, const float* scale, const void* bias, void* out, int out_dtype, int m, int n, int k,
                int by, int bx, float* slabs, float* norm /* 2 floats, device */, cudaStream_t st);
// e2m1 [n, k/2] with e4m3 scale per 16 (+ global scale) or e8m0 scale per 32 x fp16 activations in K8 order
bool fp4_tc_supported(int m, int n, int k);
void fp4_tc_run(bool mx, const void* x_f16_k8, const void* blocks, const void* scales, float global_scale, const void* bias, void* out, int out_dtype,
                int m, int n, int k, float* slabs, float* norm /* 2 floats, device */, cudaStream_t st);
// library-owned scratch (marlin_api.cu): grown outside stream capture or set once with b200_set_scratch()
void* get_scratch(size_t bytes, cudaStream_t st);

// dense 16-bit GEMM on tcgen05 (dense_gemm.cu): y[m,n] = x[m,k] . w[n,k]^T (+ bias); operands f16 / bf16 (`dtype`), out f16 / bf16 / f32
// accumulate (f32 output only): y += result.  allow_split_k: decode-size calls with f32 output may cut k over several CTAs whose partial
// sums meet in y through fp32 atomics (y must then hold the addend: zeros, or the accumulate target)
bool dense_gemm_16(const void* x, const void* w, const void* bias, void* y, int m, int n, int k, int64_t ldx, int64_t ldw, int64_t ldy,
                   int dtype, int out_dtype, cudaStream_t st, int accumulate = 0, int allow_split_k = 0);
// GGML blocks -> fp16 [n, k] (natural order), one rounding
bool dequantize_f16(const void* w, void* out_f16, int64_t n, int64_t k, int ggml_type, cudaStream_t st);

// picks tc or generic; y row stride ldy (elements)
void qmatmul_dispatch(const void* x_f16, const void* w, float* y, int64_t ldy, int m, int n, int k,
                      int ggml_type, int accumulate, cudaStream_t st);

}  // namespace b200
