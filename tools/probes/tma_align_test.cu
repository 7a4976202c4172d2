// probe: does a 2-D TMA load (u16 elements, no swizzle, 224-byte box rows) tolerate a start coordinate
// whose byte address is only 2-byte aligned?  (decides how Q6_K blocks -- 210 B, 2-byte aligned -- are staged)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, uint16_t* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem + 8192), dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(112 * 2 * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(&m), "r"(bar), "r"(c0), "r"(0) : "memory");
        asm volatile("{\n.reg .pred p;\nL: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra L;\nD:\n}\n" ::"r"(bar) : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 112 * 4; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}
int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int W = 1680, H = 4;                       // u16 elements per row (3360 B pitch), rows
    std::vector<uint16_t> h(W * H); for (int i = 0; i < W * H; ++i) h[i] = (uint16_t)i;
    uint16_t *d, *o; cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 112 * 4 * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap m; cuuint64_t dims[2] = {W, H}, str[1] = {W * 2}; cuuint32_t box[2] = {112, 4}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    for (int c0 : {0, 8, 104, 105, 1, 1575}) {
        k<<<1, 128, 16384>>>(m, c0, o);
        cudaError_t e = cudaDeviceSynchronize();
        uint16_t ho[4] = {0}; if (e == cudaSuccess) cudaMemcpy(ho, o, 8, cudaMemcpyDeviceToHost);
        printf("c0=%d: %s first=%u %u (expect %d %d)\n", c0, cudaGetErrorString(e), ho[0], ho[1], c0, c0 + 1);
        if (e != cudaSuccess) break;
    }
    return 0;
}
