// qmatmul_tc.cu -- tcgen05 dequant-GEMM for weight-only quantised linears at decode batch sizes (m <= 64):
// GGML Q4_K / Q6_K (QMatMul), symmetric int4 (marlin_4bit_*) and block-scaled FP8 (fp8_matmul).
//
//   y[m, n] (+)= sum_k x[m, k] * dequant(W)[n, k]        W row-major over n, verbatim checkpoint bytes (GGUF blocks, ...)
//
// The op is an HBM-bound weight stream (0.5625 B/weight for Q4_K, 114 flop/B at m = 32); the tensor core is there so that
// the CUDA cores only have to DEQUANTISE, never multiply.  19 warps per CTA, one CTA per SM, persistent:
//   * "swap-AB": the weight tile is the UMMA A operand (M = 128 weight rows), the activations are the B operand
//     (N = 32 or 64 batch rows), D[128 x N] fp32 lives in TMEM (two accumulators, folded in the epilogue).
//   * W producer warp: raw weight bytes of one unit (128 rows x 256 weights: 18 KB for Q4_K) by TMA into a deep ring that
//     the dequant warps hand back as soon as the bytes are in registers.  X producer warp: the matching 256-wide slice of
//     the fp16 activations (4 swizzled {64, N} boxes, L2 resident) into a shallow ring released by the MMA commit.
//   * 16 dequant warps: thread = weight row = TMEM lane; warp w serves lane quadrant w & 3 and quarter w >> 2 (64 weights)
//     of the unit.  Nibbles become fp16 with a subnormal bit trick (nibble in mantissa bits 6-9 = q * 2^-18, one HFMA2 with
//     (d*sc*2^18, -dmin*m) gives d*sc*q - dmin*m with a single rounding) and go straight into TMEM (tcgen05.st) as the
//     A operand (3 buffers) -- dequantised weights never touch shared memory.  Nibble pairs come out as (k0,k2),(k1,k3)
//     per 4 weights; instead of permuting them the activations are stored in the matching "K4" order (B200_F16_K4).
//   * MMA warp: warp-uniform loop, one elected lane issues tcgen05.mma kind::f16 (A from TMEM, B from the swizzled smem
//     tile), 16 k-steps per unit, commits to the X-stage / A-buffer barriers.
//   * stream-K: the (tile, unit) space is cut into one contiguous range per CTA (whole tiles per CTA when there are >= 4
//     tiles per SM); split tiles are reduced with fp32 red.global.add (the residual add of wo / w2 falls out for free) or
//     written to per-CTA slabs (deterministic mode; the 16-bit-output GEMMs add a finishing pass).
// Measured behaviour, fixes and dead ends: DESIGN.md 4.2, profiles/r01_qmatmul_tc_ncu.md.
//
// Reference semantics: QMatMul::forward (/root/reference/src/openai/models/linear.rs:765-806); block formats SURVEY.md
// Appendix A.  The reference's CUDA path quantises activations to Q8_1 and uses dp4a; here activations are fp16 and
// accumulation fp32 (strictly closer to the fp32 target).
#include <cuda.h>

#include <cstdlib>

#include "qmatmul.cuh"
#include "tc_common.cuh"

namespace b200 {

namespace {

using namespace tc;

template <int kMB, int kType>   // batch rows padded to kMB (32 or 64) = UMMA N; GGML type of W
struct Cfg {
    // bytes of one super-block as staged in shared memory: Q4_K 144; Q6_K a 240-byte window that starts at the
    // 210-byte block's address rounded down to 16 (TMA box starts must be 16-byte aligned)
    // Independent accumulators: back-to-back tcgen05.mma into ONE accumulator serialise on the accumulate
    // dependency; k-step ks goes to D[ks % kAcc] and the epilogue adds them up.  Two suffice (the dequant side sets the
    // pace) and keep the epilogue's TMEM reads (64 B/clk) short: kAcc * kMB <= 128 TMEM columns.
    static constexpr int kAcc = 2;
    static constexpr int kBlk = kType == B200_GGML_Q4_K ? 144 : (is_4bit_rows(kType) ? 128 : (kType == kTypeF8 ? 256 : 240));
    static constexpr int kWBytes = kTileN * kBlk;                  // raw weights of one unit (18 KB / 30 KB)
    static constexpr int kXBytes = 4 * kMB * kXSubBytes;           // 4 sub-tiles of [kMB][64] fp16 (16 KB / 32 KB)
    // Two rings.  Raw weight bytes are dead as soon as the dequant warps have pulled them into registers, so the
    // W ring is deep and recycles fast (more HBM reads in flight); the activation slices live until their MMAs
    // retire and come from L2, so the X ring is shallow.
    static constexpr int kXStages = kMB == 32 ? 4 : 2;
    static constexpr int kWStages = (227 * 1024 - 1024 - kXStages * kXBytes) / kWBytes > 8 ? 8 : (227 * 1024 - 1024 - kXStages * kXBytes) / kWBytes;
    static constexpr int kXOff = 0;                                // 1024-aligned (128-byte swizzle)
    static constexpr int kWOff = kXStages * kXBytes;
    static constexpr int kBars = kWOff + kWStages * kWBytes;       // w_full[kW] w_empty[kW] x_full[kX] x_empty[kX] a_ready[3] a_free[3] d_full d_empty
    static constexpr int kNumBars = 2 * kWStages + 2 * kXStages + 2 * kABufs + 2;
    static constexpr int kTmemSlot = kBars + kNumBars * 8;
    static constexpr int kTotal = kTmemSlot + 16;
    static_assert(kWStages >= 3, "W ring too shallow");
    static_assert(kTotal <= 232448, "exceeds the 227 KB shared memory of an SM");
};

}  // namespace
}  // namespace b200
#include "moe.cuh"
namespace b200 {
namespace {

constexpr int kMaxSeg = 3;           // weight matrices sharing one activation (fused QKV, gate|up) in one launch
struct GemmParams {
    float* y[kMaxSeg];             // output base of each segment (row stride ldy)
    int n[kMaxSeg];                // rows of each weight matrix
    int tile_end[kMaxSeg];         // cumulative 128-row tile count
    int64_t ldy;
    int m, nsb;                    // nsb = k / 256
    int n_tiles;                   // tiles over all segments
    int accumulate;
    int whole_tiles;               // 1: CTA ranges are whole tiles (no split-K, no atomics)
    int slabs;                     // > 0: split tiles write their partial sums to DISTINCT slabs (plain stores, no atomics, no
    int64_t slab_stride;           //      pre-zeroed output): slab s of segment sg starts at y[sg] + s * slab_stride; the consumer sums
    const void* scales;            // marlin: [K/g, N] in marlin-permuted order (f16 / bf16: scale_bf16); fp8: f32 [N/by, K/bx]
    const void* scales_seg[kMaxSeg];   // marlin, several weight matrices in one launch (engine: QKV, gate|up): scales / zero points of each
    const uint32_t* zp_seg[kMaxSeg];   //   segment (slot 0 = scales / zp)
    int group_size, k;             // marlin group size / fp8 bx
    int scale_bf16, scale_by, scale_sk;
    const float* norm;             // fp8: {2^p, 2^-p} range shift of the tile scales (device)
    const uint32_t* zp;            // AWQ: packed zero points [K/g, N/8] in the marlin layout (null: symmetric, zero point 8)
    // grouped (mixture-of-experts) mode: the tile list lives in device memory -- built from the routing of this step (moe.cu) --
    // and every "tile" is one (expert, 128-row weight tile, chunk of <= kMB routed rows): weights at row w_row0 of the stacked
    // [E * N, K] tensor, activations at row x_row0 of the expert-sorted fp16 copy, results scattered to row_map[x_row0 + i]
    const MoeItem* items;          // null: dense mode
    const int* num_items;          // device: number of items of this step
    const uint32_t* row_map;       // sorted position -> output row
    const float* row_scale;        // per OUTPUT row multiplier (top-k routing weight) or null
    // L2 prefetch of the NEXT GEMM's weights (decode engine): once this CTA's own loads are all issued, its W producer asks the L2 to
    // fetch its 1 / grid share of up to three tensors -- HBM is idle during the drain / epilogue / next prologue (~3 us per boundary)
    const char* pf_ptr[3];
    unsigned int pf_bytes[3];      // 0 = unused; multiples of 16
    long long* trace;              // profiling aid: per-unit clock64 stamps of CTA 0 (B200_GEMM_TRACE)
    int debug;                     // profiling aid (B200_GEMM_DEBUG): 1 = skip MMA issue, 2 = skip dequant, 4 = skip epilogue stores
};
__device__ __forceinline__ int seg_of_tile(const GemmParams& p, int tile) { return (tile >= p.tile_end[0]) + (tile >= p.tile_end[1]); }
// first tile of segment sg (constant indices only: dynamic indexing would spill the parameter struct to local memory)
__device__ __forceinline__ int seg_first_tile(const GemmParams& p, int sg) { return sg == 0 ? 0 : (sg == 1 ? p.tile_end[0] : p.tile_end[1]); }

// Symmetric int4 (GPTQ) in the repacked layout written by gptq_repack(): per row, per 64-k chunk, 32 bytes whose low
// nibbles are k = 0..31 and high nibbles k = 32..63 of the chunk -- the same nibble geometry as a Q4_K chunk, so the
// subnormal-placement + HFMA2 trick applies unchanged: w = s*q - 8*s with one per-group scale s
// (/root/reference/src/backend/gptq.rs:115-178 call site; scales arrive marlin-permuted, linear.rs:341-379).
struct M4Ctx { const void* scales; int n_total, group_size, k0, n_idx, bf16, dbg, by, sk; const float* norm; const uint32_t* zp;
               float pre[4]; int has_pre; };   // int4: {s0, s1, z0, z1} of this thread's quarter, loaded one unit ahead by the caller   // by / sk: fp8 scale-tile rows, scale columns; norm: fp8 range shift; zp: AWQ zero points (marlin layout) or null
__device__ __forceinline__ int marlin_scale_pos(int n, bool grouped) {
    // inverse of marlin_permute_scales: position of original column n inside its permuted block
    if (grouped) { const int b = n & 63; return (n & ~63) | (8 * (b & 7) + (b >> 3)); }
    const int b = n & 31, i = (b & 7) >> 1, r = b - 2 * i;          // r = 8a + c, c in {0,1}
    return (n & ~31) | (8 * i + 2 * (r >> 3) + (r & 1));
}
template <int kC>
struct M4Quarter {
    static constexpr int kRaw = 8;
    // the unit's 128 rows x 128 bytes land with the 128-byte TMA swizzle (16-byte chunk c of row r sits at chunk c ^ (r & 7)): the threads of
    // a quarter-warp read 8 different bank groups.  [Unswizzled, every lane of a warp hit the same 4 banks -- 8-way conflicts on both
    // LDS.128 of every unit, which alone cost as much as a whole Q4_K unit.]
    static __device__ __forceinline__ void load(const uint8_t* blk, int r7, uint32_t (&raw)[kRaw]) {
        const uint4 qa = *reinterpret_cast<const uint4*>(blk + (((2 * kC) ^ r7) << 4));
        const uint4 qb = *reinterpret_cast<const uint4*>(blk + (((2 * kC + 1) ^ r7) << 4));
        raw[0] = qa.x; raw[1] = qa.y; raw[2] = qa.z; raw[3] = qa.w; raw[4] = qb.x; raw[5] = qb.y; raw[6] = qb.z; raw[7] = qb.w;
    }
    // raw fetches (no arithmetic on the loaded value: issued one unit ahead, the consumer converts them when it needs them)
    static __device__ __forceinline__ uint32_t scale_bits_at(const M4Ctx& c, int k) {
        const bool grouped = c.group_size > 0;
        const int g = grouped ? k / c.group_size : 0;
        const int64_t idx = (int64_t)g * c.n_total + marlin_scale_pos(c.n_idx, grouped);
        return static_cast<const unsigned short*>(c.scales)[idx];
    }
    static __device__ __forceinline__ float scale_from_bits(const M4Ctx& c, uint32_t bits) {
        return c.bf16 ? __uint_as_float(bits << 16) : __half2float(__ushort_as_half((unsigned short)bits));
    }
    static __device__ __forceinline__ float scale_at(const M4Ctx& c, int k) { return scale_from_bits(c, scale_bits_at(c, k)); }
    // zero point of column n_idx in group g: symmetric GPTQ = 8; AWQ = nibble of the packed [G, N/8] tensor in the layout the
    // reference's converter writes (examples/convert_awq_marlin.py:75-113: scale_perm inside 64-column blocks, then the
    // [0,2,4,6,1,3,5,7] interleave inside 8): column b of a block sits in word (b & 7), nibble inv_interleave[b >> 3]
    static __device__ __forceinline__ uint32_t zero_word_at(const M4Ctx& c, int k) {
        if (!c.zp) return 0u;
        const int g = c.group_size > 0 ? k / c.group_size : 0;
        const int b = c.n_idx & 63;
        return c.zp[(int64_t)g * (c.n_total >> 3) + ((c.n_idx >> 6) << 3) + (b & 7)];
    }
    static __device__ __forceinline__ float zero_from_word(const M4Ctx& c, uint32_t word) {
        if (!c.zp) return 8.f;
        const int b = c.n_idx & 63;
        const int nib = (0x73625140u >> (4 * (b >> 3))) & 7;          // inverse of [0,2,4,6,1,3,5,7]: 0,4,1,5,2,6,3,7
        return (float)((word >> (4 * nib)) & 0xFu);
    }
    static __device__ __forceinline__ float zero_at(const M4Ctx& c, int k) { return zero_from_word(c, zero_word_at(c, k)); }
    static __device__ __forceinline__ void compute(const uint32_t (&raw)[kRaw], const M4Ctx& c, uint32_t a_col) {
        // c.pre: the raw scale bits / zero-point words of this quarter, fetched one unit ahead (bit-cast floats)
        const float s0 = c.has_pre ? scale_from_bits(c, __float_as_uint(c.pre[0])) : scale_at(c, c.k0 + kC * 64);
        const float s1 = c.has_pre ? scale_from_bits(c, __float_as_uint(c.pre[1])) : scale_at(c, c.k0 + kC * 64 + 32);
        const float z0 = c.has_pre ? zero_from_word(c, __float_as_uint(c.pre[2])) : zero_at(c, c.k0 + kC * 64);
        const float z1 = c.has_pre ? zero_from_word(c, __float_as_uint(c.pre[3])) : zero_at(c, c.k0 + kC * 64 + 32);
        const bool fast = __all_sync(0xffffffffu, fmaxf(fabsf(s0), fabsf(s1)) * 262144.f <= 65504.f);
        const __half2 s_lo = __float2half2_rn(fast ? s0 * 262144.f : s0), s_hi = __float2half2_rn(fast ? s1 * 262144.f : s1);
        const __half2 n_lo = __float2half2_rn(-z0 * s0), n_hi = __float2half2_rn(-z1 * s1);
        uint32_t v[32];
        if (fast) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t x = raw[i];
                uint32_t t0 = (x << 6) & 0x03c003c0u, t1 = (x >> 2) & 0x03c003c0u, t2 = (x << 2) & 0x03c003c0u, t3 = (x >> 6) & 0x03c003c0u;
                const __half2 r0 = __hfma2(*reinterpret_cast<__half2*>(&t0), s_lo, n_lo), r1 = __hfma2(*reinterpret_cast<__half2*>(&t1), s_lo, n_lo);
                const __half2 r2 = __hfma2(*reinterpret_cast<__half2*>(&t2), s_hi, n_hi), r3 = __hfma2(*reinterpret_cast<__half2*>(&t3), s_hi, n_hi);
                v[2 * i] = *reinterpret_cast<const uint32_t*>(&r0); v[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r1);
                v[16 + 2 * i] = *reinterpret_cast<const uint32_t*>(&r2); v[16 + 2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r3);
            }
        } else {
            const uint32_t magic = 0x64006400u;
            const __half2 k1024 = *reinterpret_cast<const __half2*>(&magic);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t x = raw[i];
                uint32_t t0 = (x & 0x000f000fu) | magic, t1 = ((x >> 8) & 0x000f000fu) | magic;
                uint32_t t2 = ((x >> 4) & 0x000f000fu) | magic, t3 = ((x >> 12) & 0x000f000fu) | magic;
                const __half2 r0 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t0), k1024), s_lo, n_lo);
                const __half2 r1 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t1), k1024), s_lo, n_lo);
                const __half2 r2 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t2), k1024), s_hi, n_hi);
                const __half2 r3 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t3), k1024), s_hi, n_hi);
                v[2 * i] = *reinterpret_cast<const uint32_t*>(&r0); v[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r1);
                v[16 + 2 * i] = *reinterpret_cast<const uint32_t*>(&r2); v[16 + 2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r3);
            }
        }
        tc_st32(a_col + kC * 32, v);
    }
};

// Block-scaled FP8 (fp8_matmul): a unit is 128 rows x 256 e4m3 bytes, staged as two {128 B x 128 rows} TMA boxes with the
// 128-byte swizzle (a plain 256-byte row pitch would put a warp's 16-byte reads 8 deep on the same banks).  Quarter kC =
// 64 bytes of this thread's row; cvt.rn.f16x2.e4m3x2 expands two weights per instruction, one HMUL2 applies the tile scale
// (group_size = bx must be a multiple of 64 so a quarter sees one scale).  Natural k order: the activation copy is plain fp16.
template <int kC>
struct F8Quarter {
    static constexpr int kRaw = 16;
    static __device__ __forceinline__ void load(const uint8_t* stage, int row, uint32_t (&raw)[kRaw]) {
        const uint8_t* base = stage + (kC >> 1) * 16384 + row * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 v = *reinterpret_cast<const uint4*>(base + ((((kC & 1) * 4 + i) ^ (row & 7)) << 4));
            raw[4 * i] = v.x; raw[4 * i + 1] = v.y; raw[4 * i + 2] = v.z; raw[4 * i + 3] = v.w;
        }
    }
    static __device__ __forceinline__ void compute(const uint32_t (&raw)[kRaw], const M4Ctx& c, uint32_t a_col) {
        // norm[0] = power of two that moves the largest tile scale to [32, 64): scaled weights then sit in fp16's normal
        // range whatever the checkpoint's scale magnitude (DeepSeek-style scales are ~1e-4); the finishing pass undoes it
        const float sc = static_cast<const float*>(c.scales)[(int64_t)(c.n_idx / c.by) * c.sk + (c.k0 + kC * 64) / c.group_size] * __ldg(c.norm);
        const __half2 S = __float2half2_rn(sc);
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(raw[i] & 0xffffu), __NV_E4M3);
            const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(raw[i] >> 16), __NV_E4M3);
            const __half2 r0 = __hmul2(__half2(lo), S), r1 = __hmul2(__half2(hi), S);
            v[2 * i] = *reinterpret_cast<const uint32_t*>(&r0);
            v[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r1);
        }
        tc_st32(a_col + kC * 32, v);
    }
};

// FP4 (e2m1) weights, two per byte in natural order (low nibble = even k), one scale per 16 (NVFP4: e4m3 byte) or 32 (MXFP4: e8m0
// byte) weights along k; a unit is 128 rows x 128 bytes like int4.  A 32-bit word holds k = 8i .. 8i+7; the nibble's three magnitude
// bits dropped at f16 bits 9-11 ARE the f16 with the same value * 2^-14 (exponent field e, mantissa bit m; e = 0 lands on an f16
// subnormal m * 2^-15 = 0.5 m * 2^-14), the sign bit goes to bit 15: two shifts and two logic ops per pair, then one HMUL2 with the
// block scale.  Pair q of word i is (k = 8i + q, k = 8i + 4 + q) -- the activations come in the matching "K8" order (fp_linear.cu).
// The block scale is kept as f16 * 2^6 (NVFP4: e4m3 bits << 7 = scale * 2^-8, times 2^14, exact incl. subnormal scales; MXFP4:
// exponent field e - 106, scales below 2^-20 flush to zero, above 2^9 saturate -- far outside any checkpoint), so the product is the
// weight * 2^-8, exact (<= 6 significant bits); the finishing pass multiplies by 2^8 (and NVFP4's global scale).
template <int kC, bool kMx>
struct F4Quarter {
    static constexpr int kRaw = 8;
    // the unit's 128 rows x 128 bytes land with the 128-byte TMA swizzle (16-byte chunk c of row r sits at chunk c ^ (r & 7)): the threads of
    // a quarter-warp read 8 different bank groups.  [Unswizzled, every lane of a warp hit the same 4 banks -- 8-way conflicts on both
    // LDS.128 of every unit, which alone cost as much as a whole Q4_K unit.]
    static __device__ __forceinline__ void load(const uint8_t* blk, int r7, uint32_t (&raw)[kRaw]) {
        const uint4 qa = *reinterpret_cast<const uint4*>(blk + (((2 * kC) ^ r7) << 4));
        const uint4 qb = *reinterpret_cast<const uint4*>(blk + (((2 * kC + 1) ^ r7) << 4));
        raw[0] = qa.x; raw[1] = qa.y; raw[2] = qa.z; raw[3] = qa.w; raw[4] = qb.x; raw[5] = qb.y; raw[6] = qb.z; raw[7] = qb.w;
    }
    static __device__ __forceinline__ void compute(const uint32_t (&raw)[kRaw], const M4Ctx& c, uint32_t a_col) {
        const uint8_t* srow = static_cast<const uint8_t*>(c.scales) + (int64_t)c.n_idx * c.sk;      // sk = scale bytes per row
        uint32_t S[4];                                   // block scale * 2^6 as f16x2, per 16 (NVFP4) / 32 (MXFP4) weights of the quarter
        if constexpr (kMx) {
            const uint32_t sw = c.has_pre ? __float_as_uint(c.pre[0]) : (uint32_t)*reinterpret_cast<const uint16_t*>(srow + ((c.k0 + kC * 64) >> 5));
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int e = (int)((sw >> (8 * j)) & 0xffu) - 106;
                e = e < 1 ? 0 : (e > 30 ? 30 : e);
                S[2 * j] = S[2 * j + 1] = ((uint32_t)e << 10) * 0x00010001u;
            }
        } else {
            const uint32_t sw = c.has_pre ? __float_as_uint(c.pre[0]) : *reinterpret_cast<const uint32_t*>(srow + ((c.k0 + kC * 64) >> 4));
            const uint32_t k16384 = 0x74007400u;         // 2^14 as f16x2
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t b = (((sw >> (8 * j)) & 0x7fu) << 7) * 0x00010001u;       // e4m3 scale * 2^-8 in both halves
                const __half2 r = __hmul2(*reinterpret_cast<__half2*>(&b), *reinterpret_cast<const __half2*>(&k16384));
                S[j] = *reinterpret_cast<const uint32_t*>(&r);
            }
        }
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t x = raw[i];
            const __half2 s = *reinterpret_cast<const __half2*>(&S[i >> 1]);
            uint32_t t0 = ((x << 9) & 0x0e000e00u) | ((x << 12) & 0x80008000u), t1 = ((x << 5) & 0x0e000e00u) | ((x << 8) & 0x80008000u);
            uint32_t t2 = ((x << 1) & 0x0e000e00u) | ((x << 4) & 0x80008000u), t3 = ((x >> 3) & 0x0e000e00u) | (x & 0x80008000u);
            const __half2 r0 = __hmul2(*reinterpret_cast<__half2*>(&t0), s), r1 = __hmul2(*reinterpret_cast<__half2*>(&t1), s);
            const __half2 r2 = __hmul2(*reinterpret_cast<__half2*>(&t2), s), r3 = __hmul2(*reinterpret_cast<__half2*>(&t3), s);
            v[4 * i] = *reinterpret_cast<const uint32_t*>(&r0); v[4 * i + 1] = *reinterpret_cast<const uint32_t*>(&r1);
            v[4 * i + 2] = *reinterpret_cast<const uint32_t*>(&r2); v[4 * i + 3] = *reinterpret_cast<const uint32_t*>(&r3);
        }
        tc_st32(a_col + kC * 32, v);
    }
};

// Q6_K staging.  A block = ql[128] | qh[64] | scales i8[16] | d f16 = 210 bytes and is only 2-byte aligned
// in memory, while a TMA box must START 16-byte aligned (probed on B200: an unaligned start raises "illegal
// instruction", tools/probes/tma_align_test.cu).  So the box starts at the block address rounded down to 16
// and is 240 bytes wide (15 x 16: an odd number of 16-byte units per row keeps LDS.128 conflict-free); the
// block then sits off = (210 * sb) % 16 bytes into the row slot, off in {0, 2, .., 14}, uniform per stage.
// Each thread loads its slot with aligned LDS.128 and realigns the words it needs in registers:
// word select by off/4 (two SEL levels) and a 16-bit funnel shift when off % 4 == 2 -- one code path.
template <int kJ0, int kN>
__device__ __forceinline__ void realign(const uint32_t (&t)[57], int wo, int sh16, uint32_t (&out)[kN]) {
    uint32_t s1[kN + 3], s2[kN + 1];
#pragma unroll
    for (int i = 0; i < kN + 3; ++i) s1[i] = (wo & 1) ? t[kJ0 + i + 1] : t[kJ0 + i];
#pragma unroll
    for (int i = 0; i < kN + 1; ++i) s2[i] = (wo & 2) ? s1[i + 2] : s1[i];
#pragma unroll
    for (int i = 0; i < kN; ++i) out[i] = __funnelshift_r(s2[i], s2[i + 1], sh16);
}

// This thread's row, half kHf of the super-block (128 weights = 4 groups of 32): group g weight l is
// ((ql[64 kHf + 32 (g&1) + l] nibble (g>>1)) | ((qh[32 kHf + l] >> 2g) & 3) << 4) - 32, times d*scales[8 kHf + 2g + l/16].
// The 6-bit value goes to fp16 mantissa bits 4-9 (= q * 2^-20); one HFMA2 with (s*2^20, -32 s) finishes it
// (exact path with the 1024+q magic when s*2^20 would overflow fp16).
template <int kGp>      // kGp = 0: groups 0 (lo nibble) and 2 (hi nibble) from ql[0..31]; kGp = 1: groups 1, 3
__device__ __forceinline__ void dequant_q6k_pair(const uint32_t (&ql)[8], uint32_t a_col, const uint32_t (&qh)[8],
                                                 const uint32_t (&scw)[2], float dk, bool fast) {
    // `fast` is warp-uniform; the two paths are separate straight-line blocks (a per-word branch interleaves them in the
    // instruction stream and the kernel then stalls on instruction fetch: ncu showed stall_no_inst on every line)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
        const int g = kGp + 2 * hi;             // group index 0..3 within the half
        __half2 sh[2], nh[2];                   // scale / offset of the two 16-weight sub-groups
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int sidx = 2 * g + j;
            const int sc = (int)(int8_t)((scw[sidx >> 2] >> (8 * (sidx & 3))) & 0xff);
            sh[j] = __float2half2_rn(dk * (float)sc);
            // q = 32 must give exactly 0: the offset is derived from the ROUNDED scale
            nh[j] = __float2half2_rn(-32.f * (fast ? __low2float(sh[j]) * (1.f / 1048576.f) : __low2float(sh[j])));
        }
        uint32_t v[16];
        if (fast) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // 6-bit value -> fp16 mantissa bits 4-9 (= q * 2^-20): low nibble at bits 4-7, the two qh bits at bits 8-9
                const uint32_t l0 = hi ? ql[i] : (ql[i] << 4);                    // bytes (0,2)
                const uint32_t l1 = hi ? (ql[i] >> 8) : (ql[i] >> 4);             // bytes (1,3)
                const uint32_t h0 = qh[i] << (8 - 2 * g), h1 = qh[i] >> (2 * g);
                uint32_t t0 = (l0 & 0x00f000f0u) | (h0 & 0x03000300u);
                uint32_t t1 = (l1 & 0x00f000f0u) | (h1 & 0x03000300u);
                const __half2 r0 = __hfma2(*reinterpret_cast<__half2*>(&t0), sh[i >> 2], nh[i >> 2]);
                const __half2 r1 = __hfma2(*reinterpret_cast<__half2*>(&t1), sh[i >> 2], nh[i >> 2]);
                v[2 * i] = *reinterpret_cast<const uint32_t*>(&r0);
                v[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r1);
            }
        } else {
            const uint32_t magic = 0x64006400u;
            const __half2 k1024 = *reinterpret_cast<const __half2*>(&magic);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t lw = hi ? (ql[i] >> 4) : ql[i];                      // nibbles now at bits 0-3 of every byte
                const uint32_t hw = qh[i] >> (2 * g);                               // 2 high bits now at bits 0-1 of every byte
                uint32_t t0 = (lw & 0x000f000fu) | ((hw << 4) & 0x00300030u) | magic;
                uint32_t t1 = ((lw >> 8) & 0x000f000fu) | ((hw >> 4) & 0x00300030u) | magic;
                const __half2 r0 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t0), k1024), sh[i >> 2], nh[i >> 2]);
                const __half2 r1 = __hfma2(__hsub2(*reinterpret_cast<__half2*>(&t1), k1024), sh[i >> 2], nh[i >> 2]);
                v[2 * i] = *reinterpret_cast<const uint32_t*>(&r0);
                v[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&r1);
            }
        }
        tc_st16(a_col + g * 16, v);
    }
}

template <int kQ>       // quarter kQ: half kHf = kQ / 2, groups (kGp, kGp + 2) with kGp = kQ % 2: 64 weights -> two 16-column TMEM stores
struct Q6KQuarter {
    static constexpr int kHf = kQ >> 1, kGp = kQ & 1;
    static constexpr int kRaw = 57;
    static __device__ __forceinline__ void load(const uint8_t* slot, int, uint32_t (&t)[kRaw]) {
#pragma unroll
        for (int c = 0; c < 14; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(slot + 16 * c);
            t[4 * c] = v.x; t[4 * c + 1] = v.y; t[4 * c + 2] = v.z; t[4 * c + 3] = v.w;
        }
        t[56] = 0u;
    }
    static __device__ __forceinline__ void compute(const uint32_t (&t)[kRaw], int off, uint32_t a_col) {
        const int wo = off >> 2, sh16 = (off & 2) * 8;
        uint32_t dw[1], scw[2], qh[8], ql[8];
        realign<52, 1>(t, wo, sh16, dw);                                              // d (f16) in the low half of dw[0]
        const float d = __half2float(__ushort_as_half((unsigned short)(dw[0] & 0xffffu)));
        realign<48 + 2 * kHf, 2>(t, wo, sh16, scw);                                   // 8 int8 scales of this half
        realign<32 + 8 * kHf, 8>(t, wo, sh16, qh);
        const bool fast = __all_sync(0xffffffffu, fabsf(d) * 128.f * 1048576.f <= 65504.f);   // warp-uniform (see Q4_K)
        const float dk = fast ? d * 1048576.f : d;
        realign<16 * kHf + 8 * kGp, 8>(t, wo, sh16, ql);
        dequant_q6k_pair<kGp>(ql, a_col + kHf * 64, qh, scw, dk, fast);
    }
};

template <int kType, int kQ> struct QuarterOf;
template <int kQ> struct QuarterOf<B200_GGML_Q4_K, kQ> { using type = Q4KQuarter<kQ>; };
template <int kQ> struct QuarterOf<B200_GGML_Q6_K, kQ> { using type = Q6KQuarter<kQ>; };
template <int kQ> struct QuarterOf<kTypeM4, kQ> { using type = M4Quarter<kQ>; };
template <int kQ> struct QuarterOf<kTypeF8, kQ> { using type = F8Quarter<kQ>; };
template <int kQ> struct QuarterOf<kTypeNV4, kQ> { using type = F4Quarter<kQ, false>; };
template <int kQ> struct QuarterOf<kTypeMX4, kQ> { using type = F4Quarter<kQ, true>; };

// =================================================================================================
// One dequant unit for quarter kQ: pull the raw bytes into registers, hand the W stage back to the producer at
// once (the bytes are dead), then wait for a free A buffer, dequantise into TMEM and signal the MMA warp.
template <int kType, int kQ>
__device__ __forceinline__ void dequant_unit(const uint8_t* blk, int off, uint32_t a_col, uint32_t w_empty_bar,
                                             uint32_t a_free_bar, uint32_t a_free_parity, uint32_t a_ready_bar, int lane, bool skip,
                                             const M4Ctx& mc) {
    using Q = typename QuarterOf<kType, kQ>::type;
    uint32_t raw[Q::kRaw];
    Q::load(blk, off, raw);
    // Cross-proxy WAR: our generic-proxy reads (LDS) must be ordered before the TMA (async proxy) refill that the
    // arrive below unblocks.  Without this fence the int4 path read refilled stages (observed on B200: units whose
    // a_free wait is long got the NEXT round's bytes); mbarrier release/acquire alone does not order the two proxies.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(w_empty_bar);
    mbar_wait(a_free_bar, a_free_parity);
    tc_fence_after();
    if (!skip) {
        if constexpr (kType == kTypeM4 || kType == kTypeF8 || is_fp4(kType)) Q::compute(raw, mc, a_col); else Q::compute(raw, off, a_col);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_ready_bar);
}

template <int kMB, int kType>
__global__ void __launch_bounds__(kThreads, 1)
qmatmul_tc_kernel(const __grid_constant__ CUtensorMap wmap0, const __grid_constant__ CUtensorMap wmap1,
                  const __grid_constant__ CUtensorMap wmap2, const __grid_constant__ CUtensorMap xmap, const GemmParams p) {
    using C = Cfg<kMB, kType>;
    constexpr int kW = C::kWStages, kX = C::kXStages;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bars = smem_base + C::kBars;
    auto w_full = [&](int s) { return bars + s * 8; };
    auto w_empty = [&](int s) { return bars + (kW + s) * 8; };
    auto x_full = [&](int s) { return bars + (2 * kW + s) * 8; };
    auto x_empty = [&](int s) { return bars + (2 * kW + kX + s) * 8; };
    auto a_ready = [&](int b) { return bars + (2 * kW + 2 * kX + b) * 8; };
    auto a_free = [&](int b) { return bars + (2 * kW + 2 * kX + kABufs + b) * 8; };
    const uint32_t d_full = bars + (2 * kW + 2 * kX + 2 * kABufs) * 8, d_empty = d_full + 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::kTmemSlot);

    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[32 * 8 + 0] = clock64();       // kernel entry
    if (threadIdx.x == 0) {
        for (int s = 0; s < kW; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), kDequantWarps); }
        for (int s = 0; s < kX; ++s) { mbar_init(x_full(s), 1); mbar_init(x_empty(s), 1); }
        for (int b = 0; b < kABufs; ++b) { mbar_init(a_ready(b), kDequantWarps); mbar_init(a_free(b), 1); }
        mbar_init(d_full, 1);
        mbar_init(d_empty, kDequantWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == kDequantWarps + 2) {      // MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[32 * 8 + 1] = clock64();       // barriers + TMEM ready
    pdl_trigger();         // the next kernel may start its own prologue; it waits for our completion before reading

    // ---- stream-K range of this CTA over the flattened (tile, super-block) space -----------------
    // (32-bit on purpose: 64-bit integer division costs ~500 cycles each on the SM and sat on the critical path of every
    // launch; the host guarantees total * gridDim.x < 2^31)
    const bool moe = p.items != nullptr;
    if (moe) pdl_wait();               // the tile list is produced by the previous kernel (routing of this step)
    const uint32_t n_tiles_rt = moe ? (uint32_t)*reinterpret_cast<const volatile int*>(p.num_items) : (uint32_t)p.n_tiles;
    const uint32_t total = n_tiles_rt * (uint32_t)p.nsb, nsb = (uint32_t)p.nsb;
    const uint32_t u0 = p.whole_tiles ? (n_tiles_rt * blockIdx.x / gridDim.x) * nsb : total * blockIdx.x / gridDim.x;
    const uint32_t u1 = p.whole_tiles ? (n_tiles_rt * (blockIdx.x + 1) / gridDim.x) * nsb : total * (blockIdx.x + 1) / gridDim.x;
    const uint32_t tile0 = u0 / nsb, sb0 = u0 - tile0 * nsb;          // first unit of this CTA

    if (warp == kDequantWarps) {
        // ================================== W PRODUCER (HBM stream) ==============================
        // warp-uniform control flow (addresses stay in uniform registers); one elected lane issues
        const bool leader = elect_one();
        uint64_t pol_w;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_w));
        int it = 0;
        int tile = (int)tile0, sb = (int)sb0;
        for (uint32_t u = u0; u < u1; ++u, ++it) {
            const int s = it % kW;
            mbar_wait(w_empty(s), ((it / kW) & 1) ^ 1);
            if (p.trace && blockIdx.x == 0 && leader && it < 32) p.trace[it * 8 + 0] = clock64();
            const int sg = moe ? 0 : seg_of_tile(p, tile);
            const CUtensorMap* wm = sg == 0 ? &wmap0 : (sg == 1 ? &wmap1 : &wmap2);
            const int wrow = moe ? p.items[tile].w_row0 : (tile - seg_first_tile(p, sg)) * kTileN;      // first weight row of this tile
            if (leader) {
                // W box start (bytes): Q4_K blocks are 144 B (16-aligned); Q6_K blocks (210 B) start at the block address
                // rounded down to 16 -- TMA box starts must be 16-byte aligned
                mbar_expect_tx(w_full(s), C::kWBytes);
                if constexpr (kType == kTypeF8) {       // two 128-byte-swizzled boxes per unit
                    tma_load_2d(smem_base + C::kWOff + s * C::kWBytes, wm, w_full(s), sb * 256, wrow, pol_w);
                    tma_load_2d(smem_base + C::kWOff + s * C::kWBytes + 16384, wm, w_full(s), sb * 256 + 128, wrow, pol_w);
                } else
                tma_load_2d(smem_base + C::kWOff + s * C::kWBytes, wm, w_full(s), kType == B200_GGML_Q4_K ? sb * 144 : (is_4bit_rows(kType) ? sb * 128 : ((sb * 210) & ~15)),
                            wrow, pol_w);
            }
            __syncwarp();
            if (++sb == (int)nsb) { sb = 0; ++tile; }
        }
        if (leader) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                if (p.pf_bytes[i] == 0) continue;
                // this CTA's slice, in 4 KB requests
                const unsigned int per = ((p.pf_bytes[i] / gridDim.x + 4095u) & ~4095u);
                const unsigned int b0 = per * blockIdx.x;
                for (unsigned int o = b0; o < b0 + per && o < p.pf_bytes[i]; o += 4096u) {
                    const unsigned int sz = p.pf_bytes[i] - o < 4096u ? p.pf_bytes[i] - o : 4096u;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.pf_ptr[i] + o), "r"(sz) : "memory");
                }
            }
        }
    } else if (warp == kDequantWarps + 1) {
        // ================================== X PRODUCER (L2 resident) =============================
        pdl_wait();        // activations come from the previous kernel (the weight stream above does not wait: weights are static)
        const bool leader = elect_one();
        uint64_t pol_x;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_x));
        int it = 0;
        int sb = (int)sb0, xtile = (int)tile0;
        for (uint32_t u = u0; u < u1; ++u, ++it) {
            const int s = it % kX;
            mbar_wait(x_empty(s), ((it / kX) & 1) ^ 1);
            const int xrow = moe ? p.items[xtile].x_row0 : 0;          // grouped mode: this tile's rows of the expert-sorted activations
            if (leader) {
                mbar_expect_tx(x_full(s), C::kXBytes);
                const uint32_t dst = smem_base + C::kXOff + s * C::kXBytes;
#pragma unroll
                for (int q = 0; q < 4; ++q) tma_load_2d(dst + q * kMB * kXSubBytes, &xmap, x_full(s), sb * kSB + q * 64, xrow, pol_x);
            }
            __syncwarp();
            if (++sb == (int)nsb) { sb = 0; ++xtile; }
        }
    } else if (warp == kDequantWarps + 2) {
        // ======================================= MMA ISSUER ======================================
        // The whole warp runs the (warp-uniform) loop so descriptors / TMEM addresses live in uniform registers;
        // one elected lane issues.  (With a single-lane loop every tcgen05.mma cost ~110 cycles of R2UR traffic.)
        const bool leader = elect_one();
        // instruction descriptor: D = f32, A = B = f16, both K-major, N = kMB, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(kMB >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);
        int it = 0, seg = 0;
        uint32_t tile = tile0;
        for (uint32_t u = u0; u < u1; ++tile) {
            const uint32_t tile_end = (tile + 1) * nsb;
            const uint32_t seg_end = tile_end < u1 ? tile_end : u1;
            mbar_wait(d_empty, (seg & 1) ^ 1);                 // epilogue of the previous segment has drained D
            tc_fence_after();
            bool first = true;
            for (; u < seg_end; ++u, ++it) {
                const int xs = it % kX, ab = it % kABufs;
                mbar_wait(x_full(xs), (it / kX) & 1);             // activations landed (TMA)
                if (p.trace && blockIdx.x == 0 && leader && it < 32) p.trace[it * 8 + 1] = clock64();
                mbar_wait(a_ready(ab), (it / kABufs) & 1);        // dequantised A tile is in TMEM
                if (p.trace && blockIdx.x == 0 && leader && it < 32) p.trace[it * 8 + 2] = clock64();
                tc_fence_after();
                const uint32_t a_t = tmem + kColA + ab * 128;
                const uint64_t bd0 = make_b_desc(smem_base + C::kXOff + xs * C::kXBytes);
                if (leader) {
                    if (!(p.debug & 1)) {
#pragma unroll
                        for (int ks = 0; ks < 16; ++ks) {
                            // k-step ks: 64-wide sub-tile (ks / 4), 32 bytes per k-step inside its 128-byte swizzled rows
                            const uint64_t bd = bd0 + (uint64_t)((((ks >> 2) * kMB * kXSubBytes) + (ks & 3) * 32) >> 4);
                            tc_mma_ts(tmem + kColD + (ks % C::kAcc) * kMB, a_t + ks * 8, bd, idesc, (first && ks < C::kAcc) ? 0u : 1u);
                        }
                    }
                    tc_commit(x_empty(xs));           // activation slice reusable once these MMAs retire
                    tc_commit(a_free(ab));            // and so is the A buffer
                    if (p.trace && blockIdx.x == 0 && it < 32) p.trace[it * 8 + 3] = clock64();
                }
                __syncwarp();
                first = false;
            }
            if (leader) tc_commit(d_full);            // accumulator of this segment complete
            __syncwarp();
            ++seg;
        }
    } else {
        // ================================ DEQUANT + EPILOGUE WARPS ================================
        const int qd = warp & 3, qt = warp >> 2;           // TMEM lane quadrant; which quarter (64 weights) of the super-block
        bool waited = false;
        const int row = qd * 32 + lane;                    // weight row within the tile = TMEM lane
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        int it = 0, seg = 0;
        int ws = 0, ab = 0;                                // ring positions and phases kept as running state (no div / mod per unit)
        uint32_t wph = 0, aph = 0;
        int tile = (int)tile0;
        // int4: the group scales / zero points live in global memory (the checkpoint's tensors), two dependent L2 round trips per unit
        // if loaded on demand -- they are fetched ONE UNIT AHEAD instead (weights: no dependency on the previous kernel)
        float m4_next[4] = {0.f, 0.f, 0.f, 0.f};
        auto m4_fetch = [&](int tl, int sbx) {
            const int sgx = seg_of_tile(p, tl);
            const int rw = (tl - seg_first_tile(p, sgx)) * kTileN + row, nn = sgx == 0 ? p.n[0] : (sgx == 1 ? p.n[1] : p.n[2]);
            const M4Ctx c{sgx == 0 ? p.scales_seg[0] : (sgx == 1 ? p.scales_seg[1] : p.scales_seg[2]), nn, p.group_size, sbx * kSB, rw < nn ? rw : 0,
                          p.scale_bf16, 0, 1, 0, nullptr, sgx == 0 ? p.zp_seg[0] : (sgx == 1 ? p.zp_seg[1] : p.zp_seg[2]), {0.f, 0.f, 0.f, 0.f}, 0};
            const int kq = sbx * kSB + qt * 64;
            if constexpr (is_fp4(kType)) {              // the quarter's block-scale bytes (4 e4m3 / 2 e8m0), bit-cast
                const uint8_t* srow = static_cast<const uint8_t*>(p.scales) + (int64_t)c.n_idx * p.scale_sk;
                m4_next[0] = __uint_as_float(kType == kTypeMX4 ? (uint32_t)*reinterpret_cast<const uint16_t*>(srow + (kq >> 5))
                                                               : *reinterpret_cast<const uint32_t*>(srow + (kq >> 4)));
            } else {
                const int gs = p.group_size;
                if (gs <= 0 || (gs >= 64 && (gs & (gs - 1)) == 0)) {
                    // the usual groups (64 / 128 / one per column): a quarter (64 weights, 64-aligned) lies in ONE group -- one fetch, a shift
                    // instead of the integer division (ncu: MUFU.RCP + ~20 IMAD per division, IMAD the most executed opcode of the kernel)
                    const uint32_t g = gs > 0 ? (uint32_t)kq >> (31 - __clz(gs)) : 0u;
                    const uint32_t bits = static_cast<const unsigned short*>(c.scales)[g * (uint32_t)nn + (uint32_t)marlin_scale_pos(c.n_idx, gs > 0)];
                    m4_next[0] = m4_next[1] = __uint_as_float(bits);
                    uint32_t zw = 0u;
                    if (c.zp) zw = c.zp[g * ((uint32_t)nn >> 3) + (((uint32_t)c.n_idx >> 6) << 3) + ((uint32_t)c.n_idx & 7u)];
                    m4_next[2] = m4_next[3] = __uint_as_float(zw);
                } else {
                    m4_next[0] = __uint_as_float(M4Quarter<0>::scale_bits_at(c, kq)); m4_next[1] = __uint_as_float(M4Quarter<0>::scale_bits_at(c, kq + 32));
                    m4_next[2] = __uint_as_float(M4Quarter<0>::zero_word_at(c, kq)); m4_next[3] = __uint_as_float(M4Quarter<0>::zero_word_at(c, kq + 32));
                }
            }
        };
        constexpr bool kPrefetchScales = kType == kTypeM4 || is_fp4(kType);
        if constexpr (kPrefetchScales) { if (u0 < u1) m4_fetch(tile, (int)(u0 - (uint32_t)tile * nsb)); }
        for (uint32_t u = u0; u < u1; ++tile) {
            const uint32_t tile_begin = (uint32_t)tile * nsb, tile_end = tile_begin + nsb;
            const uint32_t seg_begin = u, seg_end = tile_end < u1 ? tile_end : u1;
            for (; u < seg_end; ++u, ++it) {
                float m4_cur[4] = {m4_next[0], m4_next[1], m4_next[2], m4_next[3]};
                if constexpr (kPrefetchScales) {
                    if (u + 1 < u1) { if (u + 1 < tile_end) m4_fetch(tile, (int)(u + 1 - tile_begin)); else m4_fetch(tile + 1, 0); }
                }
                mbar_wait(w_full(ws), wph);
                if (p.trace && blockIdx.x == 0 && threadIdx.x == 0 && it < 32) p.trace[it * 8 + 4] = clock64();
                const uint8_t* blk = smem + C::kWOff + ws * C::kWBytes + (kType == kTypeF8 ? 0 : row * C::kBlk);    // fp8: stage base (swizzled rows)
                const uint32_t a_col = tmem + kColA + ab * 128 + lane_addr;
                // fp8: row (stage-relative addressing); 4-bit rows: swizzle key; Q6_K: block offset in its window
                const int off = kType == kTypeF8 ? row : (is_4bit_rows(kType) ? (row & 7) : (kType == B200_GGML_Q4_K ? 0 : (((int)(u - tile_begin) * 210) & 15)));
                const uint32_t afp = aph ^ 1;
                const bool skip = (p.debug & 2) != 0;
                // marlin: the scales / zero points of the weight matrix this tile belongs to (constant indices only, see seg_first_tile)
                const int msg = kType == kTypeM4 ? seg_of_tile(p, tile) : 0;
                int mrow = (tile - seg_first_tile(p, msg)) * kTileN + row;
                const int mn = msg == 0 ? p.n[0] : (msg == 1 ? p.n[1] : p.n[2]);
                if (moe) {                                 // grouped FP8: the scale row is the row of the STACKED [E * n, k] tensor
                    const MoeItem mi = p.items[tile];
                    mrow = mi.n0 + row < mn ? mi.w_row0 + row : 0;
                } else if (mrow >= mn) mrow = 0;
                const M4Ctx mc{msg == 0 ? p.scales_seg[0] : (msg == 1 ? p.scales_seg[1] : p.scales_seg[2]), mn, p.group_size, (int)(u - tile_begin) * kSB,
                               mrow, p.scale_bf16, p.debug, p.scale_by, p.scale_sk, p.norm,
                               msg == 0 ? p.zp_seg[0] : (msg == 1 ? p.zp_seg[1] : p.zp_seg[2]), {m4_cur[0], m4_cur[1], m4_cur[2], m4_cur[3]},
                               kPrefetchScales ? 1 : 0};
                switch (qt) {
                    case 0: dequant_unit<kType, 0>(blk, off, a_col, w_empty(ws), a_free(ab), afp, a_ready(ab), lane, skip, mc); break;
                    case 1: dequant_unit<kType, 1>(blk, off, a_col, w_empty(ws), a_free(ab), afp, a_ready(ab), lane, skip, mc); break;
                    case 2: dequant_unit<kType, 2>(blk, off, a_col, w_empty(ws), a_free(ab), afp, a_ready(ab), lane, skip, mc); break;
                    default: dequant_unit<kType, 3>(blk, off, a_col, w_empty(ws), a_free(ab), afp, a_ready(ab), lane, skip, mc); break;
                }
                if (p.trace && blockIdx.x == 0 && threadIdx.x == 0 && it < 32) p.trace[it * 8 + 6] = clock64();
                if (++ws == kW) { ws = 0; wph ^= 1; }
                if (++ab == kABufs) { ab = 0; aph ^= 1; }
            }
            // ---- epilogue of the segment: D (TMEM) -> y ------------------------------------------------
            if (!waited) { pdl_wait(); waited = true; }       // y may still be in use by earlier kernels
            if (p.trace && blockIdx.x == 0 && threadIdx.x == 0 && seg < 3) p.trace[32 * 8 + 2 + 2 * seg] = clock64();   // dequant of the segment done
            mbar_wait(d_full, seg & 1);
            tc_fence_after();
            if (p.trace && blockIdx.x == 0 && threadIdx.x == 0 && seg < 3) p.trace[32 * 8 + 3 + 2 * seg] = clock64();   // accumulator complete
            const bool whole = (seg_begin == tile_begin) && (seg_end == tile_end);
            if (moe) {
                // grouped mode: whole item per CTA; column i of the accumulator = routed row x_row0 + i of this expert's chunk, scattered
                // to its output row (pair index) with the routing weight folded in
                const MoeItem item = p.items[tile];
                const int n_loc = item.n0 + row;
                constexpr int kColsPerWarpM = kMB / 4;
                const float post = (kType == kTypeF8 && p.norm) ? __ldg(p.norm + 1) : 1.f;      // FP8: undo the range shift of the tile scales
#pragma unroll
                for (int c0 = 0; c0 < kColsPerWarpM; c0 += 8) {
                    uint32_t acc[8], more[8];
                    tc_ld8(tmem + kColD + lane_addr + qt * kColsPerWarpM + c0, acc);
                    tc_ld8(tmem + kColD + kMB + lane_addr + qt * kColsPerWarpM + c0, more);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (n_loc < p.n[0]) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int mi = qt * kColsPerWarpM + c0 + i;
                            if (mi < item.count) {
                                const uint32_t orow = __ldg(p.row_map + item.x_row0 + mi);
                                const float sc = p.row_scale ? __ldg(p.row_scale + orow) : 1.f;
                                p.y[0][(int64_t)orow * p.ldy + n_loc] = (__uint_as_float(acc[i]) + __uint_as_float(more[i])) * (sc * post);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty);
                ++seg;
                continue;
            }
            const int sg = seg_of_tile(p, tile);
            const int n_idx = (tile - seg_first_tile(p, sg)) * kTileN + row;
            float* ybase = sg == 0 ? p.y[0] : (sg == 1 ? p.y[1] : p.y[2]);
            // slab mode: the k-th CTA that works on a tile writes slab k.  CTA b owns units [total*b/grid, total*(b+1)/grid),
            // so the owner of unit u is ((u+1)*grid - 1) / total and only a CTA's first segment can start inside a tile.
            int ordinal = 0;
            if (p.slabs > 0 && seg_begin != tile_begin) ordinal = (int)blockIdx.x - (int)(((tile_begin + 1) * gridDim.x - 1) / total);   // 32-bit, see above
            if (p.slabs > 0) ybase += (int64_t)ordinal * p.slab_stride;
            const int zero_slabs = (p.slabs > 0 && seg_end == tile_end) ? p.slabs - 1 - ordinal : 0;   // slabs nobody else writes
            const int n_rows = sg == 0 ? p.n[0] : (sg == 1 ? p.n[1] : p.n[2]);
            constexpr int kColsPerWarp = kMB / 4;                 // this warp's share of the batch columns (8 or 16)
#pragma unroll
            for (int c0 = 0; c0 < kColsPerWarp; c0 += 8) {
                uint32_t acc[8];
                tc_ld8(tmem + kColD + lane_addr + qt * kColsPerWarp + c0, acc);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int a = 1; a < C::kAcc; ++a) {            // fold the independent accumulators
                    uint32_t more[8];
                    tc_ld8(tmem + kColD + a * kMB + lane_addr + qt * kColsPerWarp + c0, more);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(more[i]));
                }
                if (n_idx < n_rows && !(p.debug & 4)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int mi = qt * kColsPerWarp + c0 + i;
                        if (mi < p.m) {
                            float* o = ybase + (int64_t)mi * p.ldy + n_idx;
                            const float val = __uint_as_float(acc[i]);
                            if (p.slabs > 0) {
                                *o = val;
                                for (int z = 1; z <= zero_slabs; ++z) o[(int64_t)z * p.slab_stride] = 0.f;
                            } else if (whole && !p.accumulate) *o = val;
                            else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(o), "f"(val) : "memory");
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d_empty);
            ++seg;
        }
    }

    // ---- teardown ---------------------------------------------------------------------------------
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[33 * 8 + 0] = clock64();           // this warp's last epilogue done
    tc_fence_before();
    __syncthreads();
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[33 * 8 + 1] = clock64();           // all roles done
    if (warp == kDequantWarps + 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// ---- host ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// CTAs of a launch.  Slab mode keeps >= nsb/8 units per CTA so that at most 9 CTAs (= slabs) ever share one tile.
int gemm_grid(int64_t n_tiles, int nsb, bool whole_tiles, bool slab_mode) {
    const int64_t total = whole_tiles ? n_tiles : n_tiles * nsb;
    int64_t grid = sm_count();
    if (total < grid) grid = total;
    if (slab_mode && !whole_tiles && n_tiles * 8 < grid) grid = n_tiles * 8;
    return (int)grid;
}

template <int kMB, int kType>
void launch(const CUtensorMap* wm, const CUtensorMap& xm, const GemmParams& p, cudaStream_t st) {
    auto kern = qmatmul_tc_kernel<kMB, kType>;
    ensure_dynamic_smem(reinterpret_cast<const void*>(kern), Cfg<kMB, kType>::kTotal);
    const int grid = gemm_grid(p.n_tiles, p.nsb, p.whole_tiles != 0, p.slabs > 0);
    launch_pdl(kern, dim3(grid), dim3(kThreads), Cfg<kMB, kType>::kTotal, st, wm[0], wm[1], wm[2], xm, p);
    count_launch();
}

bool make_w_map(CUtensorMap* wm, const void* w, int n, int nsb, int ggml_type) {
    EncodeTiledFn enc = encode_fn();
    const bool q4 = ggml_type == B200_GGML_Q4_K, m4 = is_4bit_rows(ggml_type), f8 = ggml_type == kTypeF8;
    // byte tensor [n][nsb * block]; box {144, 128} (Q4_K), {128, 128} (int4; fp8 with the 128-byte swizzle, two per unit) or
    // {240, 128} (Q6_K: 210-byte block + alignment slack)
    const cuuint64_t pitch = (cuuint64_t)nsb * (q4 ? 144 : (m4 ? 128 : (f8 ? 256 : 210)));
    const cuuint64_t dims[2] = {pitch, (cuuint64_t)n};
    const cuuint64_t strides[1] = {pitch};
    const cuuint32_t box[2] = {(cuuint32_t)(q4 ? 144 : ((m4 || f8) ? 128 : 240)), (cuuint32_t)kTileN};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(wm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           (f8 || m4) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error(kErrCuda, "qmatmul: weight tensor map failed (%d)", (int)r); return false; }
    return true;
}

// fp8: norm = {2^p, 2^-p} with max|scale| * 2^p in [32, 64)
__global__ void fp8_scale_norm_kernel(const float* __restrict__ scale, int64_t count, float* __restrict__ norm) {
    pdl_wait();
    pdl_trigger();
    float mx = 0.f;
    for (int64_t i = threadIdx.x; i < count; i += blockDim.x) mx = fmaxf(mx, fabsf(scale[i]));
    __shared__ float red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
        int e = 0;
        if (mx > 0.f && mx < 3.0e38f) { frexpf(mx, &e); e = 6 - e; }           // mx = f * 2^e', f in [0.5, 1): mx * 2^(6 - e') in [32, 64)
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
        norm[0] = ldexpf(1.f, e); norm[1] = ldexpf(1.f, -e);
    }
}

// out[m,n] (f16 / bf16) = sum of the fp32 slabs (+ bias): the finishing pass of the 16-bit-output GEMMs below
template <typename T>
__global__ void finish_slabs_kernel(const float* __restrict__ slabs, int n_slabs, int64_t slab_stride, const T* __restrict__ bias,
                                    T* __restrict__ out, int64_t total, int n, const float* __restrict__ norm) {
    pdl_wait();
    pdl_trigger();
    const float post = norm ? norm[1] : 1.f;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += (int64_t)gridDim.x * blockDim.x * 4) {
        float4 a = *reinterpret_cast<const float4*>(slabs + i);
        for (int sl = 1; sl < n_slabs; ++sl) {
            const float4 b = *reinterpret_cast<const float4*>(slabs + sl * slab_stride + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        a.x *= post; a.y *= post; a.z *= post; a.w *= post;
        if (bias) { const int c = (int)(i % n); a.x += to_f32(bias[c]); a.y += to_f32(bias[c + 1]); a.z += to_f32(bias[c + 2]); a.w += to_f32(bias[c + 3]); }
        T o[4] = {from_f32<T>(a.x), from_f32<T>(a.y), from_f32<T>(a.z), from_f32<T>(a.w)};
        *reinterpret_cast<uint2*>(out + i) = *reinterpret_cast<const uint2*>(o);
    }
}

// y[m][ldy] (f32) (+)= sum of the fp32 slabs: the finishing pass of the int4 GEMM inside the decode engine (f32 residual stream)
__global__ void finish_slabs_f32_kernel(const float* __restrict__ slabs, int n_slabs, int64_t slab_stride, float* __restrict__ y, int64_t ldy,
                                        int64_t total, int n, int accumulate) {
    pdl_wait();
    pdl_trigger();
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += (int64_t)gridDim.x * blockDim.x * 4) {
        float4 a = *reinterpret_cast<const float4*>(slabs + i);
        for (int sl = 1; sl < n_slabs; ++sl) {
            const float4 b = *reinterpret_cast<const float4*>(slabs + sl * slab_stride + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        float4* o = reinterpret_cast<float4*>(y + (i / n) * ldy + (i % n));
        if (accumulate) { const float4 c = *o; a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w; }
        *o = a;
    }
}

}  // namespace

bool qmatmul_tc_supported(int m, int n, int k, int ggml_type) {
    if (m < 1 || m > 64 || n < 1 || k < 256 || k % 256) return false;
    if ((int64_t)((n + kTileN - 1) / kTileN + 2) * (k / 256) * sm_count() >= (int64_t)1 << 30) return false;   // kernel uses 32-bit unit arithmetic
    if (ggml_type == B200_GGML_Q4_K) return true;                         // row pitch (k/256)*144 is always a multiple of 16
    if (ggml_type == B200_GGML_Q6_K) return ((int64_t)(k / 256) * 210) % 16 == 0;   // TMA row pitch: k % 2048 == 0
    return false;
}

// Decomposition: with many tiles per SM every CTA gets whole tiles (no split); otherwise stream-K over (tile, super-block).
static bool use_whole_tiles(int64_t n_tiles) { return n_tiles >= 4 * (int64_t)sm_count(); }

// true when some output tile is produced by more than one CTA (or accumulate): y must then hold the
// addend (zeros for a plain product) before the launch
bool qmatmul_tc_needs_zeroed_output(int n, int k) {
    const int64_t n_tiles = (n + kTileN - 1) / kTileN, nsb = k / 256, total = n_tiles * nsb;
    if (use_whole_tiles(n_tiles)) return false;
    const int64_t grid = total < sm_count() ? total : sm_count();
    for (int64_t c = 1; c < grid; ++c)
        if ((total * c / grid) % nsb != 0) return true;
    return false;
}

// slab mode: how many slabs the decomposition writes for n_tiles tiles of nsb super-blocks (max CTAs sharing one tile)
int qmatmul_tc_slab_count(int64_t n_tiles, int nsb) {
    if (n_tiles <= 0 || nsb <= 0 || use_whole_tiles(n_tiles)) return 1;
    const int64_t total = n_tiles * nsb, grid = gemm_grid(n_tiles, nsb, false, true);
    int mx = 1;
    for (int64_t t = 0; t < n_tiles; ++t) {
        const int64_t first = ((t * nsb + 1) * grid - 1) / total, last = ((t + 1) * nsb * grid - 1) / total;
        if ((int)(last - first + 1) > mx) mx = (int)(last - first + 1);
    }
    return mx;
}

// up to kMaxSeg weight matrices (same type, same k) applied to one activation in one launch.
// slabs_avail = 0: y receives the product (split tiles red.add into it: see qmatmul_tc_needs_zeroed_output).
// slabs_avail > 0: split tiles store their partial sums to distinct slabs y[sg] + s * slab_stride (plain stores, nothing to
// pre-zero, bitwise deterministic); returns the number of slabs the consumer has to add up.
namespace { thread_local const void* t_pf_ptr[3]; thread_local size_t t_pf_bytes[3]; thread_local int t_pf_n = 0; }
// the weights the NEXT GEMM of the caller's chain will stream: the next qmatmul_tc_multi / marlin_tc_f32_multi launch of this thread asks
// the L2 to prefetch them from its tail (consumed by that launch)
void qmatmul_tc_prefetch_next(int n, const void* const* ptrs, const size_t* bytes) {
    static const bool off = [] { const char* e = getenv("B200_GEMM_PREFETCH"); return e && atoi(e) == 0; }();
    t_pf_n = 0;
    if (off) return;
    for (int i = 0; i < n && i < 3; ++i) {
        if (!ptrs[i] || ((uintptr_t)ptrs[i] & 15) || bytes[i] < 16 || bytes[i] >= ((size_t)1 << 31)) continue;
        t_pf_ptr[t_pf_n] = ptrs[i]; t_pf_bytes[t_pf_n] = bytes[i] & ~(size_t)15; ++t_pf_n;
    }
}
static void take_prefetch(GemmParams& p) {
    for (int i = 0; i < 3; ++i) { p.pf_ptr[i] = i < t_pf_n ? static_cast<const char*>(t_pf_ptr[i]) : nullptr; p.pf_bytes[i] = i < t_pf_n ? (unsigned int)t_pf_bytes[i] : 0u; }
    t_pf_n = 0;
}
int qmatmul_tc_multi(const void* x_f16, int nseg, const void* const* w, float* const* y, const int* n, int64_t ldy, int m, int k,
                     int ggml_type, int accumulate, int slabs_avail, int64_t slab_stride, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error(kErrCuda, "qmatmul: cuTensorMapEncodeTiled unavailable"); return 0; }
    if (nseg < 1 || nseg > kMaxSeg) { set_error(kErrBadArg, "qmatmul: %d segments (max %d)", nseg, kMaxSeg); return 0; }
    if ((uintptr_t)x_f16 & 15) { set_error(kErrBadArg, "qmatmul: x must be 16-byte aligned"); return 0; }
    const int nsb = k / 256;
    const int mb = m <= 32 ? 32 : 64;
    CUtensorMap wm[kMaxSeg], xm;
    GemmParams p{};
    int tiles = 0;
    for (int i = 0; i < kMaxSeg; ++i) {
        const int j = i < nseg ? i : nseg - 1;               // unused slots alias the last segment
        if ((uintptr_t)w[j] & 15) { set_error(kErrBadArg, "qmatmul: w must be 16-byte aligned"); return 0; }
        if (!make_w_map(&wm[i], w[j], n[j], nsb, ggml_type)) return 0;
        if (i < nseg) tiles += (n[i] + kTileN - 1) / kTileN;
        p.y[i] = y[j]; p.n[i] = n[j]; p.tile_end[i] = i < nseg ? tiles : 0x7fffffff;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)m};
        const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)mb};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x_f16), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error(kErrCuda, "qmatmul: activation tensor map failed (%d)", (int)r); return 0; }
    }
    if ((int64_t)tiles * nsb * sm_count() >= (int64_t)1 << 31) { set_error(kErrUnsupported, "qmatmul: %d tiles x %d super-blocks exceed the 32-bit unit range", tiles, nsb); return 0; }
    p.ldy = ldy; p.m = m; p.nsb = nsb; p.n_tiles = tiles; p.accumulate = accumulate;
    p.whole_tiles = use_whole_tiles(tiles) ? 1 : 0;
    take_prefetch(p);
    if (slabs_avail > 0) {
        p.slabs = qmatmul_tc_slab_count(tiles, nsb);
        p.slab_stride = slab_stride;
        p.accumulate = 0;
        if (p.slabs > slabs_avail) { set_error(kErrBadArg, "qmatmul: %d slabs needed, %d provided", p.slabs, slabs_avail); return 0; }
    }
    { static const char* dbg = getenv("B200_GEMM_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
    { static const char* tr = getenv("B200_GEMM_TRACE"); p.trace = tr ? reinterpret_cast<long long*>(strtoull(tr, nullptr, 0)) : nullptr; }
    if (ggml_type == B200_GGML_Q4_K) { if (mb == 32) launch<32, B200_GGML_Q4_K>(wm, xm, p, st); else launch<64, B200_GGML_Q4_K>(wm, xm, p, st); }
    else { if (mb == 32) launch<32, B200_GGML_Q6_K>(wm, xm, p, st); else launch<64, B200_GGML_Q6_K>(wm, xm, p, st); }
    check_launch("qmatmul_tc");
    return p.slabs;
}

// ---- weight-only GEMMs with 16-bit output (marlin_4bit_*, fp8_matmul) ----------------------------------------------
// stream-K over all SMs with fp32 partial-sum slabs in the caller's scratch, then one small finishing pass that adds the
// slabs (and the bias) and rounds once to f16 / bf16.  scratch = [S][m][n] f32, S = wq16_slabs(n, k).
int wq16_slabs(int n, int k) { return qmatmul_tc_slab_count((n + kTileN - 1) / kTileN, k / 256); }

static void wq16_launch(int kind, const void* x_f16, const void* w, const void* scales, int scale_bf16, int group_or_bx, int by, int sk,
                        const float* norm, const uint32_t* zp, const void* bias, void* out, int out_dtype, int m, int n, int k, float* slabs, cudaStream_t st,
                        const char* who, int64_t ldy_f32 = 0, int accumulate_f32 = 0) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error(kErrCuda, "%s: cuTensorMapEncodeTiled unavailable", who); return; }
    if (((uintptr_t)x_f16 | (uintptr_t)w) & 15) { set_error(kErrBadArg, "%s: x and weights must be 16-byte aligned", who); return; }
    const int nsb = k / 256;
    const int mb = m <= 32 ? 32 : 64;
    CUtensorMap wm[kMaxSeg], xm;
    for (int i = 0; i < kMaxSeg; ++i) if (!make_w_map(&wm[i], w, n, nsb, kind)) return;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)m};
        const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)mb};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x_f16), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error(kErrCuda, "%s: activation tensor map failed (%d)", who, (int)r); return; }
    }
    GemmParams p{};
    const int tiles = (n + kTileN - 1) / kTileN;
    for (int i = 0; i < kMaxSeg; ++i) { p.y[i] = slabs; p.n[i] = n; p.tile_end[i] = i == 0 ? tiles : 0x7fffffff; }
    p.ldy = n; p.m = m; p.nsb = nsb; p.n_tiles = tiles; p.accumulate = 0;
    p.whole_tiles = use_whole_tiles(tiles) ? 1 : 0;
    p.slabs = qmatmul_tc_slab_count(tiles, nsb);
    p.slab_stride = (int64_t)m * n;
    p.scales = scales; p.group_size = group_or_bx; p.k = k; p.scale_bf16 = scale_bf16; p.scale_by = by; p.scale_sk = sk; p.norm = norm; p.zp = zp;
    for (int i = 0; i < kMaxSeg; ++i) { p.scales_seg[i] = scales; p.zp_seg[i] = zp; }
    { static const char* dbg = getenv("B200_GEMM_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
    if (kind == kTypeM4) { if (mb == 32) launch<32, kTypeM4>(wm, xm, p, st); else launch<64, kTypeM4>(wm, xm, p, st); }
    else if (kind == kTypeNV4) { if (mb == 32) launch<32, kTypeNV4>(wm, xm, p, st); else launch<64, kTypeNV4>(wm, xm, p, st); }
    else if (kind == kTypeMX4) { if (mb == 32) launch<32, kTypeMX4>(wm, xm, p, st); else launch<64, kTypeMX4>(wm, xm, p, st); }
    else { if (mb == 32) launch<32, kTypeF8>(wm, xm, p, st); else launch<64, kTypeF8>(wm, xm, p, st); }
    if (!check_launch(who)) return;
    const int64_t total = (int64_t)m * n;
    int64_t g = (total / 4 + 255) / 256;
    if (g > (int64_t)sm_count() * 4) g = (int64_t)sm_count() * 4;
    if (out_dtype == B200_F32)
        launch_pdl(finish_slabs_f32_kernel, dim3((int)g), dim3(256), 0, st, (const float*)slabs, p.slabs, p.slab_stride, (float*)out, ldy_f32, total, n, accumulate_f32);
    else if (out_dtype == B200_BF16)
        launch_pdl(finish_slabs_kernel<__nv_bfloat16>, dim3((int)g), dim3(256), 0, st, (const float*)slabs, p.slabs, p.slab_stride, (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, total, n, norm);
    else
        launch_pdl(finish_slabs_kernel<__half>, dim3((int)g), dim3(256), 0, st, (const float*)slabs, p.slabs, p.slab_stride, (const __half*)bias, (__half*)out, total, n, norm);
    count_launch();
    check_launch(who);
}

// int4 x fp16 (marlin_4bit_* / marlin_awq_4bit_*): w in the gptq_repack() / awq_repack() layout, scales marlin-permuted,
// qzeros = null (symmetric, zero point 8) or the converter's packed AWQ zero points; activations fp16 in K4 order
void marlin_tc(const void* x_f16_k4, const void* w, const void* scales, const void* qzeros, void* out, int out_dtype, int m, int n, int k,
               int group_size, float* slabs, cudaStream_t st) {
    wq16_launch(kTypeM4, x_f16_k4, w, scales, out_dtype == B200_BF16, group_size, 1, 0, nullptr, static_cast<const uint32_t*>(qzeros), nullptr, out,
                out_dtype, m, n, k, slabs, st, qzeros ? "marlin_awq_4bit" : "marlin_4bit");
}

// the same GEMM for the decode engine: up to three int4 weight matrices that read the same activations (QKV, gate|up) in ONE launch,
// f32 output y[i][m][ldy] (+)= x . ((q - z) * s)^T.  Like the GGML engine path, tiles split over K meet in y through red.global.add:
// with accumulate = 0 the caller provides zeroed outputs (the engine's consumers leave their inputs zeroed).  n[i] % 4 == 0.
void marlin_tc_f32_multi(const void* x_f16_k4, int nseg, const void* const* w, const void* const* scales, int scale_bf16, const void* const* qzeros,
                         float* const* y, const int* n, int64_t ldy, int m, int k, int group_size, int accumulate, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error(kErrCuda, "marlin_4bit(f32): cuTensorMapEncodeTiled unavailable"); return; }
    if (nseg < 1 || nseg > kMaxSeg) { set_error(kErrBadArg, "marlin_4bit(f32): %d segments (max %d)", nseg, kMaxSeg); return; }
    if ((uintptr_t)x_f16_k4 & 15) { set_error(kErrBadArg, "marlin_4bit(f32): x must be 16-byte aligned"); return; }
    const int nsb = k / 256;
    const int mb = m <= 32 ? 32 : 64;
    CUtensorMap wm[kMaxSeg], xm;
    GemmParams p{};
    int tiles = 0;
    for (int i = 0; i < kMaxSeg; ++i) {
        const int j = i < nseg ? i : nseg - 1;               // unused slots alias the last segment
        if ((uintptr_t)w[j] & 15) { set_error(kErrBadArg, "marlin_4bit(f32): w must be 16-byte aligned"); return; }
        if (!make_w_map(&wm[i], w[j], n[j], nsb, kTypeM4)) return;
        if (i < nseg) tiles += (n[i] + kTileN - 1) / kTileN;
        p.y[i] = y[j]; p.n[i] = n[j]; p.tile_end[i] = i < nseg ? tiles : 0x7fffffff;
        p.scales_seg[i] = scales[j]; p.zp_seg[i] = static_cast<const uint32_t*>(qzeros ? qzeros[j] : nullptr);
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)m};
        const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)mb};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x_f16_k4), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error(kErrCuda, "marlin_4bit(f32): activation tensor map failed (%d)", (int)r); return; }
    }
    p.ldy = ldy; p.m = m; p.nsb = nsb; p.n_tiles = tiles; p.accumulate = accumulate;
    p.whole_tiles = use_whole_tiles(tiles) ? 1 : 0;
    p.scales = p.scales_seg[0]; p.zp = p.zp_seg[0]; p.group_size = group_size; p.k = k; p.scale_bf16 = scale_bf16; p.scale_by = 1;
    take_prefetch(p);
    { static const char* dbg = getenv("B200_GEMM_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
    if (mb == 32) launch<32, kTypeM4>(wm, xm, p, st); else launch<64, kTypeM4>(wm, xm, p, st);
    check_launch("marlin_4bit(f32)");
}

// block-scaled e4m3 weights x fp16 (fp8_matmul); activations fp16 in natural order
bool fp8_tc_supported(int m, int n, int k, int by, int bx) {
    if (m < 1 || m > 64 || n < 4 || n % 4 || k < 256 || k % 256 || by < 1 || bx < 64 || bx % 64) return false;
    return (int64_t)((n + kTileN - 1) / kTileN + 2) * (k / 256) * sm_count() < ((int64_t)1 << 30);
}
void fp8_tc_run(const void* x_f16, const void* w, const float* scale, const void* bias, void* out, int out_dtype, int m, int n, int k,
                int by, int bx, float* slabs, float* norm, cudaStream_t st) {
    const int sk = (k + bx - 1) / bx;
    launch_pdl(fp8_scale_norm_kernel, dim3(1), dim3(256), 0, st, scale, (int64_t)((n + by - 1) / by) * sk, norm);
    count_launch();
    wq16_launch(kTypeF8, x_f16, w, scale, 0, bx, by, sk, norm, nullptr, bias, out, out_dtype, m, n, k, slabs, st, "fp8_matmul");
}

// e2m1 weights x fp16 (nvfp4_matmul / mxfp4_matmul); activations fp16 in K8 order (within every aligned group of 8 along k:
// positions 0..7 hold k = 0,4,1,5,2,6,3,7 -- the order the nibble pairs of a 32-bit word fall out in)
__global__ void fp4_post_scale_kernel(float* __restrict__ norm, float post) {
    pdl_wait();
    pdl_trigger();
    norm[0] = 1.f; norm[1] = post;
}
bool fp4_tc_supported(int m, int n, int k) {
    if (m < 1 || m > 64 || n < 4 || n % 4 || k < 256 || k % 256) return false;
    return (int64_t)((n + kTileN - 1) / kTileN + 2) * (k / 256) * sm_count() < ((int64_t)1 << 30);
}
void fp4_tc_run(bool mx, const void* x_f16_k8, const void* blocks, const void* scales, float global_scale, const void* bias, void* out, int out_dtype,
                int m, int n, int k, float* slabs, float* norm, cudaStream_t st) {
    launch_pdl(fp4_post_scale_kernel, dim3(1), dim3(1), 0, st, norm, 256.f * global_scale);
    count_launch();
    wq16_launch(mx ? kTypeMX4 : kTypeNV4, x_f16_k8, blocks, scales, 0, 0, 1, mx ? k / 32 : k / 16, norm, nullptr, bias, out, out_dtype, m, n, k, slabs, st,
                mx ? "mxfp4_matmul" : "nvfp4_matmul");
}

// ---- grouped (mixture-of-experts) form ------------------------------------------------------------------------------------------------
bool qmatmul_tc_moe_supported(int n, int k, int ggml_type) {
    if (n < 1 || k < 256 || k % 256) return false;
    if (ggml_type == B200_GGML_Q4_K) return true;
    if (ggml_type == B200_GGML_Q6_K) return ((int64_t)(k / 256) * 210) % 16 == 0;
    return false;
}

// block-scaled FP8 experts (moe_gemm_fp8): w = e4m3 [E * n, k], scale f32 [E * n / by, ceil(k / bx)] (n % by == 0), activations fp16 in
// NATURAL order; norm = 2 device floats (scratch): the range shift of the scales, computed here
bool qmatmul_tc_moe_fp8_supported(int n, int k, int by, int bx) { return n >= 1 && k >= 256 && k % 256 == 0 && by >= 1 && n % by == 0 && bx >= 64 && bx % 64 == 0; }
void qmatmul_tc_moe_fp8(const void* xs_f16, int xs_rows, const void* w, const float* scale, int num_experts, float* y, int64_t ldy, int n, int k, int by, int bx,
                        const MoeItem* items, const int* num_items, int max_items, const uint32_t* row_map, const float* row_scale, float* norm,
                        cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error(kErrCuda, "moe_gemm_fp8: cuTensorMapEncodeTiled unavailable"); return; }
    if (((uintptr_t)xs_f16 | (uintptr_t)w) & 15) { set_error(kErrBadArg, "moe_gemm_fp8: x and w must be 16-byte aligned"); return; }
    const int nsb = k / 256, sk = (k + bx - 1) / bx;
    launch_pdl(fp8_scale_norm_kernel, dim3(1), dim3(256), 0, st, scale, (int64_t)num_experts * (n / by) * sk, norm);
    count_launch();
    CUtensorMap wm[kMaxSeg], xm;
    for (int i = 0; i < kMaxSeg; ++i) if (!make_w_map(&wm[i], w, num_experts * n, nsb, kTypeF8)) return;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)xs_rows};
        const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
        const cuuint32_t box[2] = {64, 32};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(xs_f16), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error(kErrCuda, "moe_gemm_fp8: activation tensor map failed (%d)", (int)r); return; }
    }
    if ((int64_t)max_items * nsb * sm_count() >= (int64_t)1 << 31) { set_error(kErrUnsupported, "moe_gemm_fp8: %d items x %d super-blocks exceed the 32-bit unit range", max_items, nsb); return; }
    GemmParams p{};
    for (int i = 0; i < kMaxSeg; ++i) { p.y[i] = y; p.n[i] = n; p.tile_end[i] = 0x7fffffff; p.scales_seg[i] = scale; p.zp_seg[i] = nullptr; }
    p.ldy = ldy; p.m = 32; p.nsb = nsb; p.n_tiles = max_items; p.accumulate = 0; p.whole_tiles = 1;
    p.items = items; p.num_items = num_items; p.row_map = row_map; p.row_scale = row_scale;
    p.scales = scale; p.group_size = bx; p.k = k; p.scale_by = by; p.scale_sk = sk; p.norm = norm;
    launch<32, kTypeF8>(wm, xm, p, st);
    check_launch("moe_gemm_fp8");
}

void qmatmul_tc_moe(const void* xs_f16_k4, int xs_rows, const void* w, int num_experts, float* y, int64_t ldy, int n, int k, int ggml_type,
                    const MoeItem* items, const int* num_items, int max_items, const uint32_t* row_map, const float* row_scale, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error(kErrCuda, "moe_gemm: cuTensorMapEncodeTiled unavailable"); return; }
    if (((uintptr_t)xs_f16_k4 | (uintptr_t)w) & 15) { set_error(kErrBadArg, "moe_gemm: x and w must be 16-byte aligned"); return; }
    const int nsb = k / 256;
    CUtensorMap wm[kMaxSeg], xm;
    for (int i = 0; i < kMaxSeg; ++i) if (!make_w_map(&wm[i], w, num_experts * n, nsb, ggml_type)) return;      // experts stacked along the rows
    {
        const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)xs_rows};
        const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
        const cuuint32_t box[2] = {64, 32};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&xm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(xs_f16_k4), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error(kErrCuda, "moe_gemm: activation tensor map failed (%d)", (int)r); return; }
    }
    if ((int64_t)max_items * nsb * sm_count() >= (int64_t)1 << 31) { set_error(kErrUnsupported, "moe_gemm: %d items x %d super-blocks exceed the 32-bit unit range", max_items, nsb); return; }
    GemmParams p{};
    for (int i = 0; i < kMaxSeg; ++i) { p.y[i] = y; p.n[i] = n; p.tile_end[i] = 0x7fffffff; }
    p.ldy = ldy; p.m = 32; p.nsb = nsb; p.n_tiles = max_items; p.accumulate = 0; p.whole_tiles = 1;
    p.items = items; p.num_items = num_items; p.row_map = row_map; p.row_scale = row_scale;
    if (ggml_type == B200_GGML_Q4_K) launch<32, B200_GGML_Q4_K>(wm, xm, p, st); else launch<32, B200_GGML_Q6_K>(wm, xm, p, st);
    check_launch("moe_gemm");
}

void qmatmul_tc(const void* x_f16, const void* w, float* y, int64_t ldy, int m, int n, int k, int ggml_type,
                int accumulate, cudaStream_t st) {
    qmatmul_tc_multi(x_f16, 1, &w, &y, &n, ldy, m, k, ggml_type, accumulate, 0, 0, st);
}

}  // namespace b200
