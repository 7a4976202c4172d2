// launch_gap_probe.cu -- how long does one dependent kernel node cost inside a CUDA graph on this GPU?
// Chains of N trivially short kernels (tiny grid / full-GPU grid with big smem), with and without programmatic dependent
// launch; reports microseconds per node.  nvcc -arch=sm_100a -o launch_gap_probe launch_gap_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_plain(int* p) { if (p && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, 1); }
__global__ void k_pdl(int* p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(p, 1);
}

static float run(void (*kern)(int*), bool pdl, int grid, int block, size_t smem, int n, int* d) {
    cudaStream_t st; cudaStreamCreate(&st);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaGraph_t g; cudaGraphExec_t ex;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
    for (int i = 0; i < n; ++i) {
        cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
        cfg.attrs = a; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kern, d);
    }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ex, g, 0);
    cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0, st); cudaGraphLaunch(ex, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaGraphExecDestroy(ex); cudaGraphDestroy(g); cudaStreamDestroy(st);
    return best * 1e3f / n;
}

int main() {
    int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    const int n = 512;
    struct { const char* name; int grid, block; size_t smem; } cfgs[] = {
        {"1 CTA x 32 thr", 1, 32, 0}, {"32 CTA x 256 thr", 32, 256, 0}, {"1536 CTA x 64 thr", 1536, 64, 0},
        {"148 CTA x 608 thr, 200 KB smem", 148, 608, 200 * 1024}};
    for (auto& c : cfgs) {
        const float a = run(k_plain, false, c.grid, c.block, c.smem, n, d);
        const float b = run(k_pdl, true, c.grid, c.block, c.smem, n, d);
        printf("%-34s: %.2f us/node plain, %.2f us/node with programmatic dependent launch\n", c.name, a, b);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
