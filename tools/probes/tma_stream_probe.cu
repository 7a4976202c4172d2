// probe: achievable HBM->smem streaming rate of TMA for a row-major byte matrix [N][pitch] (Q4_K weights:
// pitch = (K/256)*144) as a function of the box shape {BW bytes, BR rows} and ring depth.  148 persistent
// CTAs walk the matrix tile by tile (128 rows) along K; the consumer only recycles stages.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nL: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra L;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
struct P { int pitch, n, bw, br, stages, boxes_per_stage, mode; const uint8_t* base; int elem; };
// stage = boxes_per_stage boxes of {bw, br}; a "tile row group" = 128 rows; walk k fastest
__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap m, const P p, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const int stage_bytes = p.bw * p.br * p.boxes_per_stage;
    const uint32_t bars = base + p.stages * ((stage_bytes + 1023) / 1024 * 1024);
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * p.stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + s * 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // units: (row group of br*boxes rows?) we keep 128-row tiles: boxes_per_stage * br == 128 rows, bw bytes along k
    const int kcols = p.pitch / p.bw, tiles = p.n / 128;
    const long long total = (long long)tiles * kcols;
    const long long u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x == 0) {
        int it = 0;
        for (long long u = u0; u < u1; ++u, ++it) {
            const int tile = (int)(u / kcols), kc = (int)(u % kcols), s = it % p.stages;
            mbar_wait(bars + (p.stages + s) * 8, ((it / p.stages) & 1) ^ 1);
            const uint32_t dst = base + s * ((stage_bytes + 1023) / 1024 * 1024), bar = bars + s * 8;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
            if (p.mode == 0) {
                for (int b = 0; b < p.boxes_per_stage; ++b)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(dst + b * p.bw * p.br), "l"(&m), "r"(bar), "r"(kc * p.bw / p.elem), "r"(tile * 128 + b * p.br) : "memory");
            } else {   // 1-D bulk: pretend the tile is contiguous (repacked weights)
                const uint8_t* src = p.base + ((long long)u * stage_bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(src), "r"(stage_bytes), "r"(bar) : "memory");
            }
        }
    } else if (threadIdx.x == 32) {
        int it = 0; unsigned long long acc = 0;
        for (long long u = u0; u < u1; ++u, ++it) {
            const int s = it % p.stages;
            mbar_wait(bars + s * 8, (it / p.stages) & 1);
            acc += smem[s * ((stage_bytes + 1023) / 1024 * 1024)];
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + (p.stages + s) * 8) : "memory");
        }
        if (acc == 0x123456789ull) *sink = acc;
    }
}
__global__ void fill(uint32_t* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (uint32_t)(i * 2654435761u) ^ (uint32_t)(i >> 7) * 40503u;
}
int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int N = 28672 * 8, pitch = 16 * 144;          // 8 x gate|up of Llama-3-8B = 528 MB per launch (>> L2, fixed costs amortised)
    uint8_t* d; cudaMalloc(&d, (size_t)N * pitch);
    fill<<<1024, 256>>>((uint32_t*)d, (size_t)N * pitch / 4);   // incompressible data
    unsigned long long* sink; cudaMalloc(&sink, 8);
    struct Cfg { int bw, br, stages, mode; } cfgs[] = {{144, 128, 6, 0}, {144, 128, 10, 0}, {288, 128, 3, 0}, {288, 128, 5, 0}, {288, 64, 6, 0}, {576, 64, 5, 0},
                                                        {576, 128, 2, 0}, {18432, 1, 6, 1}, {18432, 1, 10, 1}, {36864, 1, 5, 1}};
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (auto c : cfgs) {
        P p{pitch, N, c.bw, c.br, c.stages, c.mode ? 1 : 128 / c.br, c.mode, d, 1};
        if (c.mode) { p.boxes_per_stage = 1; p.bw = c.bw; p.br = 1; }
        const int stage_bytes = c.mode ? c.bw : c.bw * 128;
        const size_t smem = (size_t)c.stages * ((stage_bytes + 1023) / 1024 * 1024) + 256;
        if (smem > 220 * 1024) { printf("skip %d x %d x %d (smem)\n", c.bw, c.br, c.stages); continue; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            CUtensorMap m; cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)N}, str[1] = {(cuuint64_t)pitch};
            cuuint32_t box[2] = {(cuuint32_t)(c.mode ? 144 : c.bw > 256 ? c.bw / 4 : c.bw), (cuuint32_t)(c.mode ? 1 : c.br)}, es[2] = {1, 1};
            // boxes wider than 256 bytes need wider elements: use u32 elements (bw / 4 per box)
            const bool wide = !c.mode && c.bw > 256;
            cuuint64_t dimsw[2] = {(cuuint64_t)pitch / 4, (cuuint64_t)N};
            CUresult r = enc(&m, wide ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, wide ? dimsw : dims, str, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); break; }
            P pp = p; pp.base = d;
            pp.elem = wide ? 4 : 1;     // box start coordinate is in tensor-map elements
            cudaEventRecord(e0);
            k<<<148, 64, smem>>>(m, pp, sink);
            cudaEventRecord(e1); cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cfg %d x %d: %s\n", c.bw, c.br, cudaGetErrorString(e)); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep >= 1 && ms < best) best = ms;
        }
        printf("%s box {%5d B x %3d rows} stages %2d: %.1f us  %.0f GB/s\n", c.mode ? "1D " : "2D ", c.bw, c.mode ? 1 : c.br, c.stages, best * 1e3,
               (double)N * pitch / best / 1e6);
    }
    return 0;
}
