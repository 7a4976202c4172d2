// qmatmul_tc.cu -- placeholder until the tcgen05 dequant-GEMM lands.
#include "qmatmul.cuh"
namespace b200 {
bool qmatmul_tc_supported(int, int, int, int) { return false; }
void qmatmul_tc(const void*, const void*, float*, int64_t, int, int, int, int, int, cudaStream_t) {}
}  // namespace b200
