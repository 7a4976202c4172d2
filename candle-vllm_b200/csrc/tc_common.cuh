// tc_common.cuh -- pieces shared by the tcgen05 kernels (qmatmul_tc.cu, layer_mega.cu): PTX wrappers for mbarrier / TMA /
// tcgen05, the UMMA shared-memory descriptor, the CTA role layout constants and the Q4_K dequant-into-TMEM quarter.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace b200 {
namespace tc {

constexpr int kTileN = 128;          // weight rows per tile (UMMA M)
constexpr int kTypeM4 = 1004;        // internal: symmetric int4 (GPTQ) in this library's repacked row-major layout (marlin_4bit_*)
constexpr int kTypeF8 = 1008;        // internal: e4m3 weights [N,K] with one f32 scale per [by, bx] tile (fp8_matmul)
constexpr int kTypeNV4 = 1016;       // internal: e2m1 pairs [N,K/2] + e4m3 scale per 16 weights (nvfp4_matmul)
constexpr int kTypeMX4 = 1032;       // internal: e2m1 pairs [N,K/2] + e8m0 scale per 32 weights (mxfp4_matmul)
constexpr bool is_fp4(int t) { return t == kTypeNV4 || t == kTypeMX4; }
constexpr bool is_4bit_rows(int t) { return t == kTypeM4 || is_fp4(t); }      // 128 raw bytes per row and unit
constexpr int kSB = 256;             // weights per super-block
constexpr int kDequantWarps = 16;     // 4 TMEM lane quadrants x 4 quarters of a super-block (64 weights per thread per unit)
constexpr int kThreads = (kDequantWarps + 3) * 32;     // + W producer, X producer, MMA issuer
constexpr int kXSubBytes = 64 * 2;                      // 128-byte swizzled row
constexpr int kColD = 0, kColA = 128, kABufs = 3;        // TMEM columns (512 allocated): D accumulators [0,128), 3 A buffers of 128


// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
// one elected lane of a converged warp (warp-uniform predicate source for the single-thread tcgen05 / TMA issue)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major) = 16 B
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset        bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B
    return d;
}


// the two 6-bit scales and mins of sub-blocks 2kC and 2kC+1 from the 12 packed bytes held in three 32-bit words
// (ggml get_scale_min_k4), packed as (lo | hi << 16) integers
template <int kC>
__device__ __forceinline__ void scale_min_pair(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t& sc2, uint32_t& mn2) {
    constexpr uint32_t sel = (kC & 1) ? 0x4342u : 0x4140u;       // bytes (2,3) or (0,1) of a word -> low bytes of the two halves
    if constexpr (kC < 2) {          // j < 4: sc = q[j] & 63, m = q[j + 4] & 63
        sc2 = __byte_perm(s0, 0u, sel) & 0x003f003fu;
        mn2 = __byte_perm(s1, 0u, sel) & 0x003f003fu;
    } else {                         // j >= 4: sc = (q[j+4] & 0xF) | ((q[j-4] >> 6) << 4), m = (q[j+4] >> 4) | ((q[j] >> 6) << 4)
        const uint32_t a = __byte_perm(s2, 0u, sel), hs = __byte_perm(s0, 0u, sel), hm = __byte_perm(s1, 0u, sel);
        sc2 = (a & 0x000f000fu) | ((hs >> 2) & 0x00300030u);
        mn2 = ((a >> 4) & 0x000f000fu) | ((hm >> 2) & 0x00300030u);
    }
}

// Dequantise sub-blocks 4*kHf .. 4*kHf+3 (128 weights) of this thread's Q4_K block into fp16 and store
// them to 64 TMEM columns (two weights per 32-bit column; K4 order inside each group of four).
// Fast path: the nibble is dropped into fp16 mantissa bits 6-9 (a subnormal = q * 2^-18) and ONE HFMA2 with
// (d*sc*2^18, -dmin*m) yields d*sc*q - dmin*m with a single rounding.  It needs d*sc*2^18 <= 65504
// (sub-block scale < 0.25, i.e. weight range < 3.75 -- always true for LLM weights); otherwise the
// exact two-step path (1024+q magic, HSUB2, HFMA2) is taken for that row.
template <int kC>        // kC = 32-byte chunk of qs: sub-blocks 2 kC (lo nibbles) and 2 kC + 1 (hi nibbles) -> TMEM columns [32 kC, 32 kC + 32)
struct Q4KQuarter {
    static constexpr int kRaw = 12;
    static __device__ __forceinline__ void load(const uint8_t* blk, int, uint32_t (&raw)[kRaw]) {
        const uint4 hdr = *reinterpret_cast<const uint4*>(blk);              // d | dmin | scales[12]
        const uint4 qa = *reinterpret_cast<const uint4*>(blk + 16 + kC * 32);
        const uint4 qb = *reinterpret_cast<const uint4*>(blk + 32 + kC * 32);
        raw[0] = hdr.x; raw[1] = hdr.y; raw[2] = hdr.z; raw[3] = hdr.w;
        raw[4] = qa.x; raw[5] = qa.y; raw[6] = qa.z; raw[7] = qa.w; raw[8] = qb.x; raw[9] = qb.y; raw[10] = qb.z; raw[11] = qb.w;
    }
    static __device__ __forceinline__ void compute(const uint32_t (&raw)[kRaw], int, uint32_t a_col) {
        const __half2 dd = *reinterpret_cast<const __half2*>(&raw[0]);          // (d, dmin)
        const float d = __low2float(dd);
        // warp-uniform so that the .aligned tcgen05.st below is reached convergently
        const bool fast = __all_sync(0xffffffffu, fabsf(d) * 63.f * 262144.f <= 65504.f);
        const __half2 dk = __float2half2_rn(fast ? d * 262144.f : d);            // exact: power-of-two scaling inside the fp16 range
        const __half2 ndmin = __hneg2(__high2half2(dd));
        // 6-bit scales / mins of sub-blocks 2kC (low half) and 2kC+1 (high half) as exact fp16 integers via the 1024+q
        // magic; one HMUL2 each then rounds d*sc and dmin*m exactly like the fp32 product followed by a cast would
        uint32_t sc2, mn2;
        scale_min_pair<kC>(raw[1], raw[2], raw[3], sc2, mn2);
        const uint32_t magic1024 = 0x64006400u;
        const __half2 h1024 = *reinterpret_cast<const __half2*>(&magic1024);
        sc2 |= magic1024; mn2 |= magic1024;
        const __half2 S = __hmul2(dk, __hsub2(*reinterpret_cast<__half2*>(&sc2), h1024));
        const __half2 N = __hmul2(ndmin, __hsub2(*reinterpret_cast<__half2*>(&mn2), h1024));
        const __half2 s_lo = __low2half2(S), s_hi = __high2half2(S), n_lo = __low2half2(N), n_hi = __high2half2(N);
        uint32_t v[32];
        if (fast) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t x = raw[4 + i];
                uint32_t t0 = (x << 6) & 0x03c003c0u, t1 = (x >> 2) & 0x03c003c0u;      // lo nibbles of bytes (0,2) and (1,3)
                uint32_t t2 = (x << 2) & 0x03c003c0u, t3 = (x >> 6) & 0x03c003c0u;      // hi nibbles of bytes (0,2) and (1,3)
                const __half2 r0 = __hfma2(*reinterpret_cast<__half2*>(&t0), s_lo, n_lo);
                const __half2 r1 = __hfma2(*reinterpret_cast<__half2*>(&t1), s_lo, n_lo);
                const __half2 r2 = __hfma2(*reinterpret_cast<__half2*>(&t2), s_hi, n_hi);
                const __half2 r3 = __hfma2(*reinterpret_cast<__half2*>(&t3), s_hi, n_hi);
                v[2 * i] = *reinterpret_cast<const uint32_t*>(&r0);
                v[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r1);
                v[16 + 2 * i] = *reinterpret_cast<const uint32_t*>(&r2);
                v[16 + 2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r3);
            }
        } else {
            const uint32_t magic = 0x64006400u;                          // half2(1024, 1024): (q | 0x6400) = 1024 + q exactly
            const __half2 k1024 = *reinterpret_cast<const __half2*>(&magic);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t x = raw[4 + i];
                uint32_t t0 = (x & 0x000f000fu) | magic, t1 = ((x >> 8) & 0x000f000fu) | magic;
                uint32_t t2 = ((x >> 4) & 0x000f000fu) | magic, t3 = ((x >> 12) & 0x000f000fu) | magic;
                const __half2 r0 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t0), k1024), s_lo, n_lo);
                const __half2 r1 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t1), k1024), s_lo, n_lo);
                const __half2 r2 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t2), k1024), s_hi, n_hi);
                const __half2 r3 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t3), k1024), s_hi, n_hi);
                v[2 * i] = *reinterpret_cast<const uint32_t*>(&r0);
                v[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r1);
                v[16 + 2 * i] = *reinterpret_cast<const uint32_t*>(&r2);
                v[16 + 2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r3);
            }
        }
        tc_st32(a_col + kC * 32, v);
    }
};

}  // namespace tc
}  // namespace b200
