// moe.cuh -- shared between moe.cu and the grouped mode of qmatmul_tc.cu
#pragma once
#include "common.cuh"

namespace b200 {

// one unit of grouped-GEMM work: a 128-row weight tile of one expert applied to a chunk of <= 32 rows routed to that expert
struct MoeItem {
    int w_row0;      // first row in the stacked [E * N, K] weight tensor (expert * N + n0)
    int x_row0;      // first row in the expert-sorted activation copy / position in sorted_token_ids
    int count;       // rows of this chunk (1 .. 32)
    int n0;          // first output column within the expert (multiple of 128)
};

// grouped dequant-GEMM on the tcgen05 pipeline (qmatmul_tc.cu): y[row_map[x_row0 + i]][n0 + r] = row_scale * sum_k xs[x_row0 + i][k] w[w_row0 + r][k]
// xs: fp16 K4 [rows_padded, k] (expert-sorted), w: stacked GGML blocks [E * n, k]; items / num_items on the device; max_items bounds the grid
bool qmatmul_tc_moe_supported(int n, int k, int ggml_type);
void qmatmul_tc_moe(const void* xs_f16_k4, int xs_rows, const void* w, int num_experts, float* y, int64_t ldy, int n, int k, int ggml_type,
                    const MoeItem* items, const int* num_items, int max_items, const uint32_t* row_map, const float* row_scale, cudaStream_t st);

// the same with block-scaled e4m3 experts; xs fp16 in natural order; norm: 2 floats of device scratch
bool qmatmul_tc_moe_fp8_supported(int n, int k, int by, int bx);
void qmatmul_tc_moe_fp8(const void* xs_f16, int xs_rows, const void* w, const float* scale, int num_experts, float* y, int64_t ldy, int n, int k, int by, int bx,
                        const MoeItem* items, const int* num_items, int max_items, const uint32_t* row_map, const float* row_scale, float* norm,
                        cudaStream_t st);

}  // namespace b200
