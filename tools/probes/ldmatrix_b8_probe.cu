// Probe (sm_100a): fragment layout of ldmatrix.m16n16.trans.b8 and the issue rate of cvt.rn.f16x2.e4m3x2 against an integer
// e4m3 -> f16 expansion.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o ldm8 ldmatrix_b8_probe.cu
#include <cstdint>
#include <cstdio>
__global__ void layout(uint32_t* out) {
    __shared__ __align__(128) uint8_t sm[2 * 16 * 16];
    for (int i = threadIdx.x; i < 2 * 256; i += 32) sm[i] = (uint8_t)i;      // matrix 0: byte = row * 16 + col; matrix 1: same + 256 (wraps)
    __syncwarp();
    uint32_t r0, r1, q[4];
    const uint32_t a1 = (uint32_t)__cvta_generic_to_shared(sm + (threadIdx.x & 15) * 16);
    asm volatile("ldmatrix.sync.aligned.m16n16.x1.trans.shared.b8 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a1));
    const uint32_t a2 = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x * 16);
    asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(a2));
    out[threadIdx.x * 6] = r0; out[threadIdx.x * 6 + 1] = r1;
    for (int i = 0; i < 4; ++i) out[threadIdx.x * 6 + 2 + i] = q[i];
}
template <int kMode>
__global__ void rate(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kMode == 0) {
                uint32_t a, b;
                asm volatile("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\ncvt.rn.f16x2.e4m3x2 %0, lo;\ncvt.rn.f16x2.e4m3x2 %1, hi;\n}\n" : "=r"(a), "=r"(b) : "r"(v[i]));
                v[i] = a ^ b;
            } else {
                const uint32_t ylo = __byte_perm(v[i], 0u, 0x1404), yhi = __byte_perm(v[i], 0u, 0x3424);
                const uint32_t a = ((ylo >> 1) & 0x3f803f80u) | (ylo & 0x80008000u), b = ((yhi >> 1) & 0x3f803f80u) | (yhi & 0x80008000u);
                v[i] = a ^ b;
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    uint32_t* d; cudaMalloc(&d, 1 << 20); long long* c; cudaMalloc(&c, 8);
    layout<<<1, 32>>>(d);
    uint32_t h[192]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("x1 (bytes low..high of r0 | r1), source byte = row*16+col:\n");
    for (int i = 0; i < 32; ++i) printf("lane %2d: %08x %08x   x2: %08x %08x %08x %08x\n", i, h[6 * i], h[6 * i + 1], h[6 * i + 2], h[6 * i + 3], h[6 * i + 4], h[6 * i + 5]);
    for (int warps = 4; warps <= 16; warps *= 2) {
        long long hc;
        rate<0><<<1, warps * 32>>>(d, c, 12345u); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("%2d warps/SM: cvt route     %lld cycles for 256 x 8 words (2 cvt each) = %.2f cycles per word per warp\n", warps, hc, hc / 2048.0);
        rate<1><<<1, warps * 32>>>(d, c, 12345u); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("%2d warps/SM: integer route %lld cycles = %.2f cycles per word per warp\n", warps, hc, hc / 2048.0);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
