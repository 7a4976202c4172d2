// qmatmul.cuh -- internal interfaces between the quantised-matmul translation units.
#pragma once
#include "common.cuh"

namespace b200 {

// shape-generic SIMT path (qmatmul_generic.cu): any m, any n, k % block == 0
void qmatmul_generic(const void* x, bool x_is_f16, const void* w, float* y, int64_t ldy, int m, int n, int k,
                     int ggml_type, int accumulate, cudaStream_t st);

// tcgen05 path (qmatmul_tc.cu): fp16 activations [m,k], m <= 32 * n_mtiles; returns false when the shape is not covered
bool qmatmul_tc_supported(int m, int n, int k, int ggml_type);
void qmatmul_tc(const void* x_f16, const void* w, float* y, int64_t ldy, int m, int n, int k, int ggml_type,
                int accumulate, cudaStream_t st);

// picks tc or generic; y row stride ldy (elements)
void qmatmul_dispatch(const void* x_f16, const void* w, float* y, int64_t ldy, int m, int n, int k,
                      int ggml_type, int accumulate, cudaStream_t st);

}  // namespace b200
