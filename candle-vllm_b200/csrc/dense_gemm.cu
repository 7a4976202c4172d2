// dense_gemm.cu -- dense 16-bit GEMM on tcgen05:  y[m, n] = sum_k x[m, k] * w[n, k]  (+ bias[n])      x, w f16 or bf16, fp32 accumulate
//
// Two users:
//   * `Linear::forward` on unquantised weights (/root/reference/src/openai/models/linear.rs:124-172; the reference calls cuBLAS through
//     candle's matmul) -- BASELINE config 2 (Llama-3-8B BF16);
//   * prefill chunks of the weight-only quantised linears (m > 64): the weights are dequantised ONCE into a 16-bit scratch
//     (dequant_to_16bit below, 2 bytes / weight of extra traffic = ~7 % of the GEMM time at m = 8192) and every m-tile reuses
//     them through this kernel -- instead of re-streaming and re-dequantising the matrix once per 64 rows.  This is also what the
//     reference does for large m (QTensor::dequantize + matmul, linear.rs:808-842 forward_via_dequant).
//
// Kernel: persistent, warp-specialised, one CTA per SM.
//   warp 0  TMA producer: A tile [128 x 64] and B tile [BN x 64] (128-byte swizzled rows) into a 4 - 6 stage ring;
//   warp 1  MMA issuer: one elected lane, tcgen05.mma.kind::f16 with both operands from shared memory (SS), UMMA 128 x BN x 16,
//           accumulators in TMEM, DOUBLE-BUFFERED (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1;
//   warps 2-5  epilogue: tcgen05.ld (each warp its TMEM lane quadrant), convert, 16-byte stores.
//   BN = 256 (128 / 64 while 256-wide tiles would leave SMs idle) for m > 64 (x is the 128-row operand); "swap-AB" with BN = 32 / 64 for m <= 64 (the weight tile is the 128-row operand,
//   the few activation rows are the UMMA N dimension), so decode-size calls stream the weights once with full-width MMAs.
// Tiles are walked n-major within a band of m-tiles so that CTAs running together share the W tile (L2).
#include <cuda.h>

#include "qmatmul.cuh"
#include "tc_common.cuh"

namespace b200 {

namespace {

using namespace tc;

constexpr int kBM = 128, kBK = 64;
constexpr int kGemmThreads = 6 * 32;

template <int kBN>
struct DCfg {
    static constexpr int kABytes = kBM * kBK * 2, kBBytes = kBN * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr int kBars = kStages * kStageBytes;              // full[kStages] empty[kStages] tmem_full[2] tmem_empty[2]
    static constexpr int kTmemSlot = kBars + (2 * kStages + 4) * 8;
    static constexpr int kTotal = kTmemSlot + 16;
    static constexpr int kTmemCols = 2 * kBN < 32 ? 32 : 2 * kBN;    // power of two >= 32: 64, 128, 512
    static_assert(kTotal <= 232448, "shared memory");
};

__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
        "%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

struct DenseParams {
    void* y; const void* bias;
    int m, n, k;
    int64_t ldy;
    int tiles_a, tiles_b;          // tiles of the 128-row operand / of the BN-row operand
    int swap;                      // 1: A = weights (rows = n), B = activations (rows = m); output transposed on the way out
    int out_dtype, bf16;
    int ksplit;                    // swap mode, f32 output: the k range is cut into ksplit parts per tile (more CTAs than weight tiles); partial
    int accumulate;                // sums meet in y through fp32 atomics.  accumulate: y += result (f32 output)
};

template <typename TOut>
__device__ __forceinline__ void store_row16(TOut* dst, const uint32_t* acc, const TOut* bias, bool has_bias) {
    // 16 consecutive outputs of one row: 32 bytes (16-bit) or 64 bytes (f32)
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(acc[i]) + (has_bias ? to_f32(bias[i]) : 0.f);
    if constexpr (sizeof(TOut) == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    } else {
        TOut o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = from_f32<TOut>(f[i]);
        reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(o)[0];
        reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(o)[1];
    }
}

template <int kBN, typename TOut>
__global__ void __launch_bounds__(kGemmThreads, 1)
dense_gemm_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, const DenseParams p) {
    using C = DCfg<kBN>;
    constexpr int kS = C::kStages;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bars = smem_base + C::kBars;
    auto full = [&](int s) { return bars + s * 8; };
    auto empty = [&](int s) { return bars + (kS + s) * 8; };
    auto tmem_full = [&](int b) { return bars + (2 * kS + b) * 8; };
    auto tmem_empty = [&](int b) { return bars + (2 * kS + 2 + b) * 8; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::kTmemSlot);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kS; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(C::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_trigger();

    const int n_tiles = p.tiles_a * p.tiles_b * p.ksplit;
    const int nkb_all = (p.k + kBK - 1) / kBK;
    // tile order: bands of 8 A-tiles; inside a band the B-tile index is the slow one, so the CTAs running together (consecutive tile
    // ids) work on the same B tile (weights, or activations when swapped) and neighbouring A tiles -> both stay in L2
    // split-K (swap mode only): tile id = (weight tile, k part); part z owns k blocks [nkb_all * z / ksplit, nkb_all * (z + 1) / ksplit)
    auto k_range = [&](int t, int& kb0, int& kb1) {
        const int z = p.ksplit > 1 ? t % p.ksplit : 0;
        kb0 = nkb_all * z / p.ksplit; kb1 = nkb_all * (z + 1) / p.ksplit;
    };
    auto tile_coords = [&](int t, int& ta, int& tb) {
        if (p.ksplit > 1) { ta = t / p.ksplit; tb = 0; return; }
        constexpr int kBand = 8;
        const int per_band = kBand * p.tiles_b;
        const int band = t / per_band, r = t - band * per_band;
        const int rows = min(kBand, p.tiles_a - band * kBand);
        tb = r / rows;
        ta = band * kBand + (r - tb * rows);
    };

    if (warp == 0) {
        // ============================================ TMA PRODUCER ==========================================================
        pdl_wait();
        const bool leader = elect_one();
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
        int it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int ta, tb, kb0, kb1;
            tile_coords(t, ta, tb);
            k_range(t, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % kS;
                mbar_wait(empty(s), ((it / kS) & 1) ^ 1);
                if (leader) {
                    mbar_expect_tx(full(s), C::kStageBytes);
                    const uint32_t dst = smem_base + s * C::kStageBytes;
                    tma_load_2d(dst, &amap, full(s), kb * kBK, ta * kBM, pol);
                    tma_load_2d(dst + C::kABytes, &bmap, full(s), kb * kBK, tb * kBN, pol);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ============================================ MMA ISSUER ============================================================
        const bool leader = elect_one();
        const uint32_t fmt = p.bf16 ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        int it = 0, tcount = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
            const int ab = tcount & 1;
            mbar_wait(tmem_empty(ab), ((tcount >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_t = tmem + ab * kBN;
            int kb0, kb1;
            k_range(t, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % kS;
                mbar_wait(full(s), (it / kS) & 1);
                tc_fence_after();
                const uint64_t ad0 = make_b_desc(smem_base + s * C::kStageBytes);
                const uint64_t bd0 = make_b_desc(smem_base + s * C::kStageBytes + C::kABytes);
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < kBK / 16; ++ks)
                        tc_mma_ss(d_t, ad0 + (uint64_t)(ks * 2), bd0 + (uint64_t)(ks * 2), idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
                    tc_commit(empty(s));
                }
                __syncwarp();
            }
            if (leader) tc_commit(tmem_full(ab));
            __syncwarp();
        }
    } else {
        // ============================================ EPILOGUE ==============================================================
        const int qd = warp & 3;                               // TMEM lane quadrant this warp may read
        const int r = qd * 32 + lane;                          // row of the 128-row operand
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        const TOut* bias = static_cast<const TOut*>(p.bias);
        int tcount = 0;
        bool waited = false;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
            int ta, tb;
            tile_coords(t, ta, tb);
            const int ab = tcount & 1;
            mbar_wait(tmem_full(ab), (tcount >> 1) & 1);
            tc_fence_after();
            if (!waited) { pdl_wait(); waited = true; }
            const int ia = ta * kBM + r;                        // index along the 128-row operand
#pragma unroll 1
            for (int c0 = 0; c0 < kBN; c0 += 32) {
                uint32_t acc[32];
                tc_ld32(tmem + ab * kBN + lane_addr + c0, acc);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int jb = tb * kBN + c0;                   // first index along the BN-row operand
                if (!p.swap) {
                    // y[ia][jb .. jb+32): this lane owns a row, 32 consecutive columns
                    if (ia < p.m) {
                        TOut* dst = static_cast<TOut*>(p.y) + (int64_t)ia * p.ldy + jb;
                        if constexpr (sizeof(TOut) == 4) {
                            if (p.accumulate) {         // y += : this CTA owns the tile exclusively (no split-K here), plain read-modify-write
                                for (int i = 0; i < 32; ++i)
                                    if (jb + i < p.n) dst[i] = dst[i] + __uint_as_float(acc[i]) + (bias ? to_f32(bias[jb + i]) : 0.f);
                                continue;
                            }
                        }
                        if (jb + 32 <= p.n && (((uintptr_t)dst) & 15) == 0) {
                            store_row16<TOut>(dst, acc, bias ? bias + jb : nullptr, bias != nullptr);
                            store_row16<TOut>(dst + 16, acc + 16, bias ? bias + jb + 16 : nullptr, bias != nullptr);
                        } else {
                            for (int i = 0; i < 32; ++i)
                                if (jb + i < p.n) dst[i] = from_f32<TOut>(__uint_as_float(acc[i]) + (bias ? to_f32(bias[jb + i]) : 0.f));
                        }
                    }
                } else {
                    // swapped: the lane owns output COLUMN ia (a weight row); y[jb + i][ia]: lanes -> consecutive addresses
                    if (ia < p.n) {
                        const bool first_part = p.ksplit <= 1 || (t % p.ksplit) == 0;
                        const float bv = (bias && first_part) ? to_f32(bias[ia]) : 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (jb + i < p.m) {
                                TOut* o = static_cast<TOut*>(p.y) + (int64_t)(jb + i) * p.ldy + ia;
                                if constexpr (sizeof(TOut) == 4) {
                                    if (p.ksplit > 1 || p.accumulate) { atomicAdd(o, __uint_as_float(acc[i]) + bv); continue; }
                                }
                                *o = from_f32<TOut>(__uint_as_float(acc[i]) + bv);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(ab));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::kTmemCols));
    }
}

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

// [rows, k] 16-bit row-major (row pitch ld elements) -> box {64, box_rows}, 128-byte swizzle, zero fill out of bounds
bool make_map(CUtensorMap* map, const void* base, int rows, int k, int64_t ld, int box_rows, bool bf16) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error(kErrCuda, "dense_gemm: cuTensorMapEncodeTiled unavailable"); return false; }
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {kBK, (cuuint32_t)box_rows};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides,
                           box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error(kErrCuda, "dense_gemm: tensor map failed (%d)", (int)r); return false; }
    return true;
}

template <int kBN, typename TOut>
void launch_dense(const CUtensorMap& am, const CUtensorMap& bm, const DenseParams& p, cudaStream_t st) {
    auto kern = dense_gemm_kernel<kBN, TOut>;
    ensure_dynamic_smem(reinterpret_cast<const void*>(kern), DCfg<kBN>::kTotal);
    const int tiles = p.tiles_a * p.tiles_b * p.ksplit;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    launch_pdl(kern, dim3(grid), dim3(kGemmThreads), DCfg<kBN>::kTotal, st, am, bm, p);
    count_launch();
}

template <int kBN>
void launch_out(const CUtensorMap& am, const CUtensorMap& bm, const DenseParams& p, cudaStream_t st) {
    if (p.out_dtype == B200_F32) launch_dense<kBN, float>(am, bm, p, st);
    else if (p.out_dtype == B200_BF16) launch_dense<kBN, __nv_bfloat16>(am, bm, p, st);
    else launch_dense<kBN, __half>(am, bm, p, st);
}

}  // namespace

// y[m, n] (out_dtype f16 / bf16 / f32, row pitch ldy) = x[m, k] . w[n, k]^T (+ bias[n], of out_dtype); x and w of `dtype` (f16 / bf16),
// row pitches ldx / ldw elements.  k % 8 == 0, 16-byte aligned bases and pitches (TMA).
bool dense_gemm_16(const void* x, const void* w, const void* bias, void* y, int m, int n, int k, int64_t ldx, int64_t ldw, int64_t ldy,
                   int dtype, int out_dtype, cudaStream_t st, int accumulate, int allow_split_k) {
    if (m <= 0 || n <= 0 || k <= 0) return true;
    if (dtype != B200_F16 && dtype != B200_BF16) { set_error(kErrUnsupported, "dense_gemm: operand dtype %d (f16 / bf16)", dtype); return false; }
    if (out_dtype != B200_F16 && out_dtype != B200_BF16 && out_dtype != B200_F32) { set_error(kErrUnsupported, "dense_gemm: out dtype %d", out_dtype); return false; }
    if (k % 8 || ldx % 8 || ldw % 8 || (((uintptr_t)x | (uintptr_t)w) & 15)) {
        set_error(kErrBadArg, "dense_gemm: k, ldx, ldw must be multiples of 8 elements and x, w 16-byte aligned (TMA)");
        return false;
    }
    const bool bf16 = dtype == B200_BF16;
    DenseParams p{};
    p.y = y; p.bias = bias; p.m = m; p.n = n; p.k = k; p.ldy = ldy; p.out_dtype = out_dtype; p.bf16 = bf16 ? 1 : 0;
    p.ksplit = 1; p.accumulate = accumulate ? 1 : 0;
    if (accumulate && out_dtype != B200_F32) { set_error(kErrUnsupported, "dense_gemm: accumulate needs f32 output"); return false; }
    CUtensorMap am, bm;
    if (m <= 64) {                       // swap-AB: weights are the 128-row operand
        const int bn = m <= 32 ? 32 : 64;
        p.swap = 1; p.tiles_a = (n + kBM - 1) / kBM; p.tiles_b = 1;
        // decode sizes are a pure weight stream: with fewer weight tiles than SMs, cut k so that every SM streams a part (f32 output
        // pre-zeroed by the caller, or accumulate: the parts meet in fp32 atomics)
        if (allow_split_k && out_dtype == B200_F32 && p.tiles_a < sm_count()) {
            const int nkb = (k + kBK - 1) / kBK;
            int ks = sm_count() / p.tiles_a;
            if (ks > nkb / 4) ks = nkb / 4;
            if (ks > 1) p.ksplit = ks;
        }
        if (!make_map(&am, w, n, k, ldw, kBM, bf16) || !make_map(&bm, x, m, k, ldx, bn, bf16)) return false;
        if (bn == 32) launch_out<32>(am, bm, p, st); else launch_out<64>(am, bm, p, st);
    } else {
        // 128 x 256 tiles when they fill the machine; mid-size m (a few hundred rows) gets narrower tiles so that every SM has one
        int bn = 256;
        auto tiles = [&](int b) { return (int64_t)((m + kBM - 1) / kBM) * ((n + b - 1) / b); };
        if (tiles(256) < sm_count()) bn = 128;
        if (bn == 128 && tiles(128) < sm_count()) bn = 64;
        p.swap = 0; p.tiles_a = (m + kBM - 1) / kBM; p.tiles_b = (n + bn - 1) / bn;
        if (!make_map(&am, x, m, k, ldx, kBM, bf16) || !make_map(&bm, w, n, k, ldw, bn, bf16)) return false;
        if (bn == 256) launch_out<256>(am, bm, p, st); else if (bn == 128) launch_out<128>(am, bm, p, st); else launch_out<64>(am, bm, p, st);
    }
    return check_launch("dense_gemm");
}

}  // namespace b200

using namespace b200;

extern "C" {

// Linear::forward on dense 16-bit weights (linear.rs:124-172): out[m, n] = x[m, k] . weight[n, k]^T (+ bias[n]); x / weight / bias / out of
// `dtype` (B200_F16 / B200_BF16), contiguous.
void linear_16bit(const void* x, const void* weight, const void* bias, void* out, int32_t m, int32_t n, int32_t k, int32_t dtype, int64_t stream) {
    if (m == 0 || n == 0) return;
    B200_REQUIRE(x && weight && out && m > 0 && n > 0 && k > 0, kErrBadArg, "linear_16bit: bad arguments");
    dense_gemm_16(x, weight, bias, out, m, n, k, k, k, n, dtype, dtype, as_stream(stream), 0, 0);
}

}  // extern "C"
