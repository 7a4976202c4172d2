// layer_mega.cu -- one persistent launch for the whole quantised-projection chain of a decoder layer:
//
//     wo  ->  [+residual, RMSNorm]  ->  gate|up  ->  [SiLU * up]  ->  w2  ->  [+residual, RMSNorm]  ->  QKV of the next layer
//
// (GGUFLLaMa::forward_inner /root/reference/src/openai/models/quantized_llama.rs:424-506, Mlp::forward :32-44,
// QuantizedAttention::forward layers/attention.rs:910-1011; tensor parallel: AllReduce after the row-parallel linears,
// distributed.rs:572-653, :696-710).
//
// Why: at decode sizes every one of those GEMMs streams 9 - 66 MB of Q4_K weights (1.6 - 11 us of HBM time), and as separate
// launches each paid ~6 us of fixed cost -- launch gap, barrier / TMEM setup, first-TMA latency, first dequant, MMA drain,
// a red.global.add epilogue -- plus ~3 us per small kernel in between (profiles/r01_qmatmul_tc_ncu.md).  Here the 19-warp CTA
// of qmatmul_tc.cu (W producer / X producer / MMA issuer / 16 dequant warps, one CTA per SM) stays resident for all phases:
//   * the WEIGHT stream never stops at a phase boundary: the W producer walks straight into the next matrix (weights do not
//     depend on anything), the dequant warps pre-dequantise its first three units into the TMEM A buffers; only the MMAs wait
//     for the activations;
//   * the elementwise op between two GEMMs (residual add + RMSNorm, SiLU * up, or -- tensor parallel -- the two-hop peer-memory
//     all-reduce + residual + RMSNorm of tp.cu) runs inside the kernel on the dequant warps, between two grid-wide counters
//     (phase p-1 complete on every CTA -> op -> activations of phase p ready);
//   * split tiles (stream-K) are reduced DETERMINISTICALLY: the k-th CTA of a tile stores its partial sum to slab k with plain
//     coalesced stores (no red.global.add: the SM retires atomics lane by lane), and the consumer -- the elementwise op, or the
//     RoPE kernel after the QKV phase -- adds the slabs of a tile in slab order.  Same inputs, same bits, every run.
// All CTAs of the launch are co-resident (grid <= SM count, 1 CTA / SM by shared-memory size), which the grid-wide counters
// need; every spin is bounded (-> trap with a host-visible flag) so a scheduling surprise cannot wedge the GPU.
#include <cuda.h>

#include <cstdlib>

#include "mega.cuh"
#include "tc_common.cuh"

namespace b200 {

namespace {

using namespace tc;

constexpr int kDeqThreads = kDequantWarps * 32;      // 512: the warps that dequantise, run the epilogues and the elementwise ops

template <int kMB>
struct MCfg {
    static constexpr int kAcc = 2;
    static constexpr int kBlk = 144;
    static constexpr int kWBytes = kTileN * kBlk;                  // 18 KB: 128 rows x one Q4_K super-block
    static constexpr int kXBytes = 4 * kMB * kXSubBytes;           // 4 sub-tiles of [kMB][64] fp16
    static constexpr int kXStages = kMB == 32 ? 4 : 2;
    static constexpr int kWStages = (227 * 1024 - 2048 - kXStages * kXBytes) / kWBytes > 8 ? 8 : (227 * 1024 - 2048 - kXStages * kXBytes) / kWBytes;
    static constexpr int kXOff = 0;
    static constexpr int kWOff = kXStages * kXBytes;
    static constexpr int kBars = kWOff + kWStages * kWBytes;
    static constexpr int kNumBars = 2 * kWStages + 2 * kXStages + 2 * kABufs + 2;
    static constexpr int kTmemSlot = kBars + kNumBars * 8;
    static constexpr int kRed = kTmemSlot + 16;                    // 32 floats of reduction scratch for the elementwise ops
    static constexpr int kTotal = kRed + 128;
    static_assert(kWStages >= 3, "W ring too shallow");
    static_assert(kTotal <= 232448, "exceeds the 227 KB shared memory of an SM");
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// grid-wide counter wait: ONE thread spins (bounded: a CTA that never arrives means the launch was not co-resident or a peer
// died -- raise the host-visible flag and trap instead of hanging the GPU)
__device__ __forceinline__ void wait_counter(const uint32_t* ctr, uint32_t target, volatile uint32_t* err_word) {
    if (ld_acquire_gpu(ctr) >= target) return;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        if (ld_acquire_gpu(ctr) >= target) return;
        if ((spins & 0xfffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) {           // 10 s
                if (err_word) { *err_word = 2u; __threadfence_system(); }
                __trap();
            }
        }
    }
}

// sum of the slabs of 4 consecutive columns (one tile) of row `mi`, added in slab order.  ALL loads are issued before the first add
// (a data-dependent loop serialises the L2 round trips: 6 slabs cost 4.9 us instead of 0.8, measured with B200_MEGA_TRACE).
template <int kMaxS>
__device__ __forceinline__ float4 slab_sum4(const float* base, int64_t ld, int64_t slab_stride, int mi, int col, int slabs) {
    const float* p = base + (int64_t)mi * ld + col;
    float4 b[kMaxS];
#pragma unroll
    for (int s = 0; s < kMaxS; ++s) b[s] = s < slabs ? ldcg4(p + (int64_t)s * slab_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a = b[0];
#pragma unroll
    for (int s = 1; s < kMaxS; ++s) { a.x += b[s].x; a.y += b[s].y; a.z += b[s].z; a.w += b[s].w; }
    return a;
}

__device__ __forceinline__ void store_f16_k4(__half* o, float a, float b, float c, float d) {
    // K4 order: the middle two of every aligned group of four swapped (include/b200_backend.h)
    const __half2 p0 = __halves2half2(from_f32<__half>(a), from_f32<__half>(c)), p1 = __halves2half2(from_f32<__half>(b), from_f32<__half>(d));
    *reinterpret_cast<uint2*>(o) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
}

// sum of squares over the 512 threads of the dequant warps (named barrier 1)
__device__ __forceinline__ float block_sum_512(float v, float* red, int tid) {
    v = warp_sum(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    named_bar_sync(1, kDeqThreads);
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < kDequantWarps; ++i) tot += red[i];
    named_bar_sync(1, kDeqThreads);                    // red[] may be reused
    return tot;
}

// ---- elementwise ops between the GEMM phases (run by the 512 dequant-warp threads of every CTA) -------------------------------
// x[row] += sum of the previous phase's slabs; act_out[row] = RMSNorm(x[row]) * w  (f16, K4 order).  One row per CTA.
__device__ __forceinline__ void eop_norm(const MegaParams& P, const MegaPhase& prev, const MegaPhase& cur, float* red, int tid, uint32_t gemm_grid) {
    const int row = blockIdx.x;
    if (row >= P.m) return;
    const int n = P.hidden, nv = n >> 2;
    const uint32_t nsb = (uint32_t)prev.nsb, total = (uint32_t)prev.n_tiles * nsb, G = total < gemm_grid ? total : gemm_grid;
    float* xr = P.x + (int64_t)row * n;
    float4 v[4];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int i = tid + it * kDeqThreads;
        if (i < nv) {
            float4 a = ldcg4(xr + 4 * i);
            if (total) {
                const float4 b = slab_sum4<kMegaMaxSlabs>(prev.y[0], prev.ldy, prev.slab_stride, row, 4 * i, tile_slabs((uint32_t)(4 * i) / kTileN, nsb, total, G));
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            *reinterpret_cast<float4*>(xr + 4 * i) = a;
            v[it] = a;
            ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
    }
    const float tot = block_sum_512(ss, red, tid);
    const float sc = rsqrtf(tot / (float)n + P.eps);
    __half* o = static_cast<__half*>(cur.act_out) + (int64_t)row * n;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int i = tid + it * kDeqThreads;
        if (i < nv) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(cur.norm_w) + i);
            store_f16_k4(o + 4 * i, v[it].x * sc * g.x, v[it].y * sc * g.y, v[it].z * sc * g.z, v[it].w * sc * g.w);
        }
    }
}

// act_out = silu(gate) * up (f16, K4 order); gate / up = slab sums of segments 0 / 1 of the previous phase.  Grid-stride.
__device__ __forceinline__ void eop_silu(const MegaParams& P, const MegaPhase& prev, const MegaPhase& cur, int tid, uint32_t gemm_grid) {
    const int F = prev.n[0], fv = F >> 2;
    const uint32_t nsb = (uint32_t)prev.nsb, total = (uint32_t)prev.n_tiles * nsb, G = total < gemm_grid ? total : gemm_grid;
    const int t_up = prev.tile_end[0];
    const int items = P.m * fv;
    __half* out = static_cast<__half*>(cur.act_out);
    for (int i = blockIdx.x * kDeqThreads + tid; i < items; i += gridDim.x * kDeqThreads) {
        const int mi = i / fv, c4 = (i - mi * fv) * 4;
        const uint32_t tile = (uint32_t)c4 / kTileN;
        const float4 g = slab_sum4<4>(prev.y[0], prev.ldy, prev.slab_stride, mi, c4, tile_slabs(tile, nsb, total, G));
        const float4 u = slab_sum4<4>(prev.y[1], prev.ldy, prev.slab_stride, mi, c4, tile_slabs(tile + (uint32_t)t_up, nsb, total, G));
        store_f16_k4(out + (int64_t)mi * F + c4, g.x / (1.f + __expf(-g.x)) * u.x, g.y / (1.f + __expf(-g.y)) * u.y,
                     g.z / (1.f + __expf(-g.z)) * u.z, g.w / (1.f + __expf(-g.w)) * u.w);
    }
}

// tensor parallel: partial = slab sum of the row-parallel GEMM; two-hop LL exchange over NVLink peer memory, residual add, next
// RMSNorm (the protocol of tp.cu: row r owned by rank r % world; {value, epoch} words; double-buffered by epoch parity)
__device__ __forceinline__ void st_ll(void* addr, uint32_t a, uint32_t b, uint32_t e) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(a), "r"(e), "r"(b), "r"(e) : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const void* addr) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void ll_wait(const void* addr, uint32_t e, uint4& w, const MegaParams& P) {
    w = ld_ll(addr);
    if (w.y == e && w.w == e) return;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        w = ld_ll(addr);
        if (w.y == e && w.w == e) return;
        if ((spins & 0x3ffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > P.tp_timeout_ns) { *P.timeout_word = 1u; __threadfence_system(); w = make_uint4(0x7fc00000u, e, 0x7fc00000u, e); return; }
        }
    }
}
__device__ __forceinline__ void eop_tp_norm(const MegaParams& P, const MegaPhase& prev, const MegaPhase& cur, float* red, int tid, uint32_t gemm_grid) {
    const int row = blockIdx.x;
    if (row >= P.m) return;
    const int n = P.hidden, nv = n >> 2, world = P.tp_world, rank = P.tp_rank, rows_max = P.rows_max;
    const uint32_t nsb = (uint32_t)prev.nsb, total = (uint32_t)prev.n_tiles * nsb, G = total < gemm_grid ? total : gemm_grid;
    const TpInboxLayout lay(world, rows_max, n);
    char* const mine = P.peers.p[rank];
    uint32_t* epoch = reinterpret_cast<uint32_t*>(mine + lay.epoch_off);
    const uint32_t e = epoch[row] + 1u;
    named_bar_sync(1, kDeqThreads);                       // everyone has read the epoch before thread 0 bumps it at the end
    const int par = (int)(e & 1u);
    const int owner = row % world, lrow = row / world;
    float4 v[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int i = tid + it * kDeqThreads;
        if (i < nv) v[it] = slab_sum4<kMegaMaxSlabs>(prev.y[0], prev.ldy, prev.slab_stride, row, 4 * i, tile_slabs((uint32_t)(4 * i) / kTileN, nsb, total, G));
    }
    __half* o = static_cast<__half*>(cur.act_out) + (int64_t)row * n;
    if (rank != owner) {
        char* dst = P.peers.p[owner] + (((size_t)par * world + rank) * lay.rows_owned + lrow) * n * 8;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int i = tid + it * kDeqThreads;
            if (i < nv) {
                st_ll(dst + (size_t)i * 32, __float_as_uint(v[it].x), __float_as_uint(v[it].y), e);
                st_ll(dst + (size_t)i * 32 + 16, __float_as_uint(v[it].z), __float_as_uint(v[it].w), e);
            }
        }
        const char* src = mine + lay.bcast_off + ((size_t)par * rows_max + row) * (n / 2) * 8;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int i = tid + it * kDeqThreads;
            if (i < nv) {
                uint4 w;
                ll_wait(src + (size_t)i * 16, e, w, P);
                *reinterpret_cast<uint2*>(o + 4 * i) = make_uint2(w.x, w.z);
            }
        }
    } else {
        float* xr = P.x + (int64_t)row * n;
        float ss = 0.f;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int i = tid + it * kDeqThreads;
            if (i < nv) {
                float4 a = ldcg4(xr + 4 * i);
                for (int p = 0; p < world; ++p) {              // rank order: one well-defined fp32 sum
                    if (p == rank) { a.x += v[it].x; a.y += v[it].y; a.z += v[it].z; a.w += v[it].w; continue; }
                    const char* src = mine + (((size_t)par * world + p) * lay.rows_owned + lrow) * n * 8 + (size_t)i * 32;
                    uint4 w0, w1;
                    ll_wait(src, e, w0, P);
                    ll_wait(src + 16, e, w1, P);
                    a.x += __uint_as_float(w0.x); a.y += __uint_as_float(w0.z); a.z += __uint_as_float(w1.x); a.w += __uint_as_float(w1.z);
                }
                *reinterpret_cast<float4*>(xr + 4 * i) = a;
                v[it] = a;
                ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
            }
        }
        const float tot = block_sum_512(ss, red, tid);
        const float sc = rsqrtf(tot / (float)n + P.eps);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int i = tid + it * kDeqThreads;
            if (i < nv) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(cur.norm_w) + i);
                const __half2 p0 = __halves2half2(from_f32<__half>(v[it].x * sc * g.x), from_f32<__half>(v[it].z * sc * g.z));
                const __half2 p1 = __halves2half2(from_f32<__half>(v[it].y * sc * g.y), from_f32<__half>(v[it].w * sc * g.w));
                const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&p0), u1 = *reinterpret_cast<const uint32_t*>(&p1);
                *reinterpret_cast<uint2*>(o + 4 * i) = make_uint2(u0, u1);
                for (int p = 0; p < world; ++p)
                    if (p != rank) st_ll(P.peers.p[p] + lay.bcast_off + ((size_t)par * rows_max + row) * (n / 2) * 8 + (size_t)i * 16, u0, u1, e);
            }
        }
    }
    named_bar_sync(1, kDeqThreads);
    if (tid == 0) epoch[row] = e;
}

// profiling aid: stamp k of phase ph of this CTA (0 phase entered, 1 units dequantised ahead / op reached, 2 previous phase complete
// everywhere, 3 op done, 4 last unit of the phase dequantised, 5 last epilogue stored, 6 first activations landed (MMA warp), 7 first
// accumulator complete)
__device__ __forceinline__ void trace_stamp(const MegaParams& P, int ph, int k) {
    if (P.trace) P.trace[((int64_t)blockIdx.x * kMegaMaxPhases + ph) * 8 + k] = clock64();
}

// The elementwise op that produces phase ph's activations from phase ph-1's slabs: wait until phase ph-1 is complete on every CTA,
// run this CTA's share, publish.  Called by all 512 dequant-warp threads at the same point of their loop.
__device__ __forceinline__ void run_eop(const MegaParams& P, int ph, float* red) {
    const int tid = threadIdx.x;
    const MegaPhase& g = P.phase[ph];
    const MegaPhase& prev = P.phase[ph - 1];
    pdl_wait();                                                                // x / the slabs may still be in use by the previous kernel
    if (tid == 0) { trace_stamp(P, ph, 1); wait_counter(P.counters + 2 * ph, gridDim.x, P.error_word); trace_stamp(P, ph, 2); }   // phase ph-1 complete everywhere
    named_bar_sync(1, kDeqThreads);
    if (g.eop == kEopNorm) eop_norm(P, prev, g, red, tid, gridDim.x);
    else if (g.eop == kEopSilu) eop_silu(P, prev, g, tid, gridDim.x);
    else eop_tp_norm(P, prev, g, red, tid, gridDim.x);
    named_bar_sync(1, kDeqThreads);                                            // every thread's stores are issued
    // bar.sync orders the other threads' stores before thread 0 at CTA scope; its gpu-scope release is cumulative over them
    if (tid == 0) { red_release_gpu_add(P.counters + 2 * ph + 1, 1u); trace_stamp(P, ph, 3); }
}

__device__ __forceinline__ int seg_of_tile(const MegaPhase& g, int tile) { return (tile >= g.tile_end[0]) + (tile >= g.tile_end[1]); }
__device__ __forceinline__ int seg_first_tile(const MegaPhase& g, int sg) { return sg == 0 ? 0 : (sg == 1 ? g.tile_end[0] : g.tile_end[1]); }

// One dequant unit for quarter kQ (see qmatmul_tc.cu): raw bytes -> registers, hand the W stage back, wait for a free A buffer,
// dequantise into TMEM, signal the MMA warp.
template <int kQ>
__device__ __forceinline__ void dequant_unit(const uint8_t* blk, uint32_t a_col, uint32_t w_empty_bar, uint32_t a_free_bar, uint32_t a_free_parity,
                                             uint32_t a_ready_bar, int lane) {
    using Q = Q4KQuarter<kQ>;
    uint32_t raw[Q::kRaw];
    Q::load(blk, 0, raw);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // our LDS reads before the TMA refill this arrive unblocks
    __syncwarp();
    if (lane == 0) mbar_arrive(w_empty_bar);
    mbar_wait(a_free_bar, a_free_parity);
    tc_fence_after();
    Q::compute(raw, 0, a_col);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_ready_bar);
}

template <int kMB>
__global__ void __launch_bounds__(kThreads, 1)
layer_mega_kernel(const __grid_constant__ MegaParams P) {
    using C = MCfg<kMB>;
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
    float* red = reinterpret_cast<float*>(smem + C::kRed);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kW; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), kDequantWarps); }
        for (int s = 0; s < kX; ++s) { mbar_init(x_full(s), 1); mbar_init(x_empty(s), 1); }
        for (int b = 0; b < kABufs; ++b) { mbar_init(a_ready(b), kDequantWarps); mbar_init(a_free(b), 1); }
        mbar_init(d_full, 1);
        mbar_init(d_empty, kDequantWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == kDequantWarps + 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (P.early_trigger) pdl_trigger();

    const uint32_t G = gridDim.x;
    // counters of this launch: [2 * ph] = GEMM phase ph-1 complete on every CTA, [2 * ph + 1] = activations of phase ph ready
    uint32_t* const ctr = P.counters;

    // unit range of this CTA in phase ph (32-bit arithmetic: the host guarantees total * G < 2^31)
    // A phase with fewer units than CTAs uses only the first `total` CTAs (one unit each): every CTA that takes part owns at
    // least one unit, so the CTAs sharing a tile are consecutive -- which the slab arithmetic (mega.cuh) relies on.
    auto range = [&](const MegaPhase& g, uint32_t& u0, uint32_t& u1) {
        const uint32_t total = (uint32_t)g.n_tiles * (uint32_t)g.nsb;
        const uint32_t ge = total < G ? total : G;
        if (blockIdx.x >= ge) { u0 = u1 = 0; return; }
        u0 = total * blockIdx.x / ge;
        u1 = total * (blockIdx.x + 1) / ge;
    };

    if (warp == kDequantWarps) {
        // ================================== W PRODUCER (HBM stream; never waits for a phase) ==============================
        const bool leader = elect_one();
        uint64_t pol_w;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_w));
        int it = 0;
        for (int ph = 0; ph < P.n_phases; ++ph) {
            const MegaPhase& g = P.phase[ph];
            uint32_t u0, u1;
            range(g, u0, u1);
            const uint32_t nsb = (uint32_t)g.nsb;
            int tile = nsb ? (int)(u0 / nsb) : 0, sb = nsb ? (int)(u0 - (uint32_t)tile * nsb) : 0;
            for (uint32_t u = u0; u < u1; ++u, ++it) {
                const int s = it % kW;
                mbar_wait(w_empty(s), ((it / kW) & 1) ^ 1);
                const int sg = seg_of_tile(g, tile);
                const CUtensorMap* wm = &P.maps[sg == 0 ? g.w_map[0] : (sg == 1 ? g.w_map[1] : g.w_map[2])];
                const int ltile = tile - seg_first_tile(g, sg);
                if (leader) {
                    mbar_expect_tx(w_full(s), C::kWBytes);
                    tma_load_2d(smem_base + C::kWOff + s * C::kWBytes, wm, w_full(s), sb * 144, ltile * kTileN, pol_w);
                }
                __syncwarp();
                if (++sb == (int)nsb) { sb = 0; ++tile; }
            }
        }
    } else if (warp == kDequantWarps + 1) {
        // ================================== X PRODUCER (L2 resident activations) =========================================
        const bool leader = elect_one();
        uint64_t pol_x;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_x));
        int it = 0;
        for (int ph = 0; ph < P.n_phases; ++ph) {
            const MegaPhase& g = P.phase[ph];
            uint32_t u0, u1;
            range(g, u0, u1);
            if (u0 == u1) continue;
            if (g.eop == kEopNone) pdl_wait();             // activations come from the previous kernel
            else {
                if (lane == 0) wait_counter(ctr + 2 * ph + 1, G, P.error_word);
                __syncwarp();
                asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy writes of other SMs before our async-proxy (TMA) reads
            }
            const CUtensorMap* xm = &P.maps[g.x_map];
            const uint32_t nsb = (uint32_t)g.nsb;
            int sb = (int)(u0 % nsb);
            for (uint32_t u = u0; u < u1; ++u, ++it) {
                const int s = it % kX;
                mbar_wait(x_empty(s), ((it / kX) & 1) ^ 1);
                if (leader) {
                    mbar_expect_tx(x_full(s), C::kXBytes);
                    const uint32_t dst = smem_base + C::kXOff + s * C::kXBytes;
#pragma unroll
                    for (int q = 0; q < 4; ++q) tma_load_2d(dst + q * kMB * kXSubBytes, xm, x_full(s), sb * kSB + q * 64, 0, pol_x);
                }
                __syncwarp();
                if (++sb == (int)nsb) sb = 0;
            }
        }
    } else if (warp == kDequantWarps + 2) {
        // ======================================= MMA ISSUER ==================================================================
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | ((uint32_t)(kMB >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);
        int it = 0, seg = 0;
        for (int ph = 0; ph < P.n_phases; ++ph) {
            const MegaPhase& g = P.phase[ph];
            uint32_t u0, u1;
            range(g, u0, u1);
            const uint32_t nsb = (uint32_t)g.nsb;
            uint32_t tile = nsb ? u0 / nsb : 0;
            for (uint32_t u = u0; u < u1; ++tile) {
                const uint32_t tile_end = (tile + 1) * nsb;
                const uint32_t seg_end = tile_end < u1 ? tile_end : u1;
                mbar_wait(d_empty, (seg & 1) ^ 1);
                tc_fence_after();
                bool first = true;
                for (; u < seg_end; ++u, ++it) {
                    const int xs = it % kX, ab = it % kABufs;
                    mbar_wait(x_full(xs), (it / kX) & 1);
                    if (P.trace && leader && u == u0) trace_stamp(P, ph, 6);
                    mbar_wait(a_ready(ab), (it / kABufs) & 1);
                    tc_fence_after();
                    const uint32_t a_t = tmem + kColA + ab * 128;
                    const uint64_t bd0 = make_b_desc(smem_base + C::kXOff + xs * C::kXBytes);
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < 16; ++ks) {
                            const uint64_t bd = bd0 + (uint64_t)((((ks >> 2) * kMB * kXSubBytes) + (ks & 3) * 32) >> 4);
                            tc_mma_ts(tmem + kColD + (ks % C::kAcc) * kMB, a_t + ks * 8, bd, idesc, (first && ks < C::kAcc) ? 0u : 1u);
                        }
                        tc_commit(x_empty(xs));
                        tc_commit(a_free(ab));
                    }
                    __syncwarp();
                    first = false;
                }
                if (leader) tc_commit(d_full);
                __syncwarp();
                ++seg;
            }
        }
    } else {
        // ================================ DEQUANT + EPILOGUE + ELEMENTWISE WARPS ==============================================
        const int tid = threadIdx.x;                       // 0 .. 511
        const int qd = warp & 3, qt = warp >> 2;
        const int row = qd * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        int seg = 0, ws = 0, ab = 0;
        uint32_t wph = 0, aph = 0;
        for (int ph = 0; ph < P.n_phases; ++ph) {
            const MegaPhase& g = P.phase[ph];
            uint32_t u0, u1;
            range(g, u0, u1);
            const uint32_t nsb = (uint32_t)g.nsb;
            bool eop_pending = g.eop != kEopNone;
            if (tid == 0) trace_stamp(P, ph, 0);
            // Units of this phase run segment by segment (a segment = this CTA's share of one tile).  The elementwise op that
            // produces the phase's activations is slotted in ONCE: after up to kABufs units have been dequantised ahead (their
            // A buffers only depend on MMAs of the previous phase), at the latest before the first epilogue wait.
            uint32_t u = u0;
            int tile = nsb ? (int)(u0 / nsb) : 0;
            uint32_t seg_begin = u0, seg_end = nsb ? ((uint32_t)(tile + 1) * nsb < u1 ? (uint32_t)(tile + 1) * nsb : u1) : u1;
            for (;;) {
                uint32_t stop = seg_end;
                if (eop_pending) { const uint32_t h = u0 + kABufs; if (h < stop) stop = h > u ? h : u; }
                for (; u < stop; ++u) {
                    mbar_wait(w_full(ws), wph);
                    const uint8_t* blk = smem + C::kWOff + ws * C::kWBytes + row * C::kBlk;
                    const uint32_t a_col = tmem + kColA + ab * 128 + lane_addr;
                    const uint32_t afp = aph ^ 1;
                    switch (qt) {
                        case 0: dequant_unit<0>(blk, a_col, w_empty(ws), a_free(ab), afp, a_ready(ab), lane); break;
                        case 1: dequant_unit<1>(blk, a_col, w_empty(ws), a_free(ab), afp, a_ready(ab), lane); break;
                        case 2: dequant_unit<2>(blk, a_col, w_empty(ws), a_free(ab), afp, a_ready(ab), lane); break;
                        default: dequant_unit<3>(blk, a_col, w_empty(ws), a_free(ab), afp, a_ready(ab), lane); break;
                    }
                    if (++ws == kW) { ws = 0; wph ^= 1; }
                    if (++ab == kABufs) { ab = 0; aph ^= 1; }
                }
                if (eop_pending) {
                    run_eop(P, ph, red);
                    eop_pending = false;
                    if (u < seg_end) continue;
                }
                if (u0 == u1) break;                       // no units in this phase: this CTA only took part in the op
                // ---- epilogue of the segment: D (TMEM) -> this CTA's slab of the tile, plain coalesced stores --------------
                pdl_wait();                                // the slabs may still be read by the previous kernel (no-op after the first time)
                if (tid == 0 && u == u1) trace_stamp(P, ph, 4);
                mbar_wait(d_full, seg & 1);
                tc_fence_after();
                {
                    const uint32_t total = (uint32_t)g.n_tiles * nsb, tile_begin = (uint32_t)tile * nsb;
                    const int sg = seg_of_tile(g, tile);
                    const int n_idx = (tile - seg_first_tile(g, sg)) * kTileN + row;
                    float* ybase = sg == 0 ? g.y[0] : (sg == 1 ? g.y[1] : g.y[2]);
                    const int ordinal = seg_begin != tile_begin ? (int)blockIdx.x - (int)unit_owner(tile_begin, total, total < G ? total : G) : 0;
                    ybase += (int64_t)ordinal * g.slab_stride;
                    const int n_rows = sg == 0 ? g.n[0] : (sg == 1 ? g.n[1] : g.n[2]);
                    constexpr int kColsPerWarp = kMB / 4;
#pragma unroll
                    for (int c0 = 0; c0 < kColsPerWarp; c0 += 8) {
                        uint32_t acc[8], more[8];
                        tc_ld8(tmem + kColD + lane_addr + qt * kColsPerWarp + c0, acc);
                        tc_ld8(tmem + kColD + kMB + lane_addr + qt * kColsPerWarp + c0, more);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (n_idx < n_rows) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int mi = qt * kColsPerWarp + c0 + i;
                                if (mi < P.m) ybase[(int64_t)mi * g.ldy + n_idx] = __uint_as_float(acc[i]) + __uint_as_float(more[i]);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty);
                ++seg;
                if (u == u1) break;
                ++tile;
                seg_begin = u;
                seg_end = (uint32_t)(tile + 1) * nsb < u1 ? (uint32_t)(tile + 1) * nsb : u1;
            }
            // ---- this CTA is done with GEMM phase ph: its slab stores must be visible before the arrival ----------------------
            if (ph + 1 < P.n_phases) {
                named_bar_sync(1, kDeqThreads);
                if (tid == 0) red_release_gpu_add(ctr + 2 * (ph + 1), 1u);
            }
            if (tid == 0) trace_stamp(P, ph, 5);
        }
    }

    // ---- teardown ------------------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == kDequantWarps + 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// The same elementwise ops as stand-alone kernels ("split" mode: one launch per GEMM phase with programmatic dependent launch between
// them -- kernel boundaries turned out cheaper than in-kernel grid syncs -- but still the deterministic slab reduction).  `src` / `dst`
// are the phase that produced the slabs and the phase that will read the activations; gemm_grid = the grid of the producing launch.
__global__ void __launch_bounds__(kDeqThreads)
mega_eop_kernel(const __grid_constant__ MegaParams P, int eop, uint32_t gemm_grid) {
    __shared__ float red[32];
    pdl_wait();
    pdl_trigger();
    const MegaPhase& prev = P.phase[0];
    const MegaPhase& cur = P.phase[1];
    if (eop == kEopNorm) eop_norm(P, prev, cur, red, threadIdx.x, gemm_grid);
    else if (eop == kEopSilu) eop_silu(P, prev, cur, threadIdx.x, gemm_grid);
    else eop_tp_norm(P, prev, cur, red, threadIdx.x, gemm_grid);
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

}  // namespace

int mega_grid() { return sm_count(); }

// slabs a phase of n_tiles x nsb units needs on a grid of G CTAs (max CTAs that share one tile)
int mega_phase_slabs(int n_tiles, int nsb, int G) {
    if (n_tiles <= 0 || nsb <= 0) return 1;
    const int64_t total = (int64_t)n_tiles * nsb;
    if (total < G) G = (int)total;                      // effective grid of the phase (see range() in the kernel)
    int mx = 1;
    for (int64_t t = 0; t < n_tiles; ++t) {
        const int64_t first = ((t * nsb + 1) * G - 1) / total, last = ((t + 1) * nsb * G - 1) / total;
        if ((int)(last - first + 1) > mx) mx = (int)(last - first + 1);
    }
    return mx;
}

bool mega_supported(int m, int hidden, int k_max) {
    return m >= 1 && m <= 64 && hidden % 256 == 0 && hidden <= 8192 && k_max % 256 == 0 && encode_fn() != nullptr;
}

bool mega_make_w_map(CUtensorMap* map, const void* w, int n, int k) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error(kErrCuda, "layer_mega: cuTensorMapEncodeTiled unavailable"); return false; }
    if ((uintptr_t)w & 15) { set_error(kErrBadArg, "layer_mega: weights must be 16-byte aligned"); return false; }
    const cuuint64_t pitch = (cuuint64_t)(k / 256) * 144;
    const cuuint64_t dims[2] = {pitch, (cuuint64_t)n};
    const cuuint64_t strides[1] = {pitch};
    const cuuint32_t box[2] = {144, (cuuint32_t)kTileN};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(w), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error(kErrCuda, "layer_mega: weight tensor map failed (%d)", (int)r); return false; }
    return true;
}

bool mega_make_x_map(CUtensorMap* map, const void* x_f16, int m, int k) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error(kErrCuda, "layer_mega: cuTensorMapEncodeTiled unavailable"); return false; }
    const int mb = m <= 32 ? 32 : 64;
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)m};
    const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)mb};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x_f16), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error(kErrCuda, "layer_mega: activation tensor map failed (%d)", (int)r); return false; }
    return true;
}

// stand-alone elementwise op between two single-phase launches: P.phase[0] = the phase whose slabs are folded, P.phase[1] = the consumer
// (eop, norm_w, act_out); no tensor maps needed
void mega_eop_launch(const MegaParams& P, int eop, cudaStream_t st) {
    const int G = mega_grid();
    const int grid = eop == kEopSilu ? G : P.m;
    launch_pdl(mega_eop_kernel, dim3(grid), dim3(kDeqThreads), 0, st, P, eop, (uint32_t)G);
    count_launch();
    check_launch("layer_mega(eop)");
}

void mega_launch(const MegaParams& P, cudaStream_t st) {
    const int G = mega_grid();
    for (int ph = 0; ph < P.n_phases; ++ph) {
        if ((int64_t)P.phase[ph].n_tiles * P.phase[ph].nsb * (int64_t)(G + 1) >= ((int64_t)1 << 31)) {
            set_error(kErrUnsupported, "layer_mega: phase %d exceeds the 32-bit unit range", ph);
            return;
        }
    }
    // tuning knobs (measured in profiles/): B200_MEGA_PDL=0 launches the kernel without programmatic serialization (it then starts
    // only after its predecessor has drained); MegaParams.early_trigger lets the NEXT kernel's grid be scheduled early
    static const int pdl_in = [] { const char* e = getenv("B200_MEGA_PDL"); return e ? atoi(e) : 1; }();
    auto go = [&](auto kern, int smem) {
        ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
        if (pdl_in) launch_pdl(kern, dim3(G), dim3(kThreads), smem, st, P);
        else kern<<<dim3(G), dim3(kThreads), smem, st>>>(P);
    };
    if (P.m <= 32) go(layer_mega_kernel<32>, MCfg<32>::kTotal);
    else go(layer_mega_kernel<64>, MCfg<64>::kTotal);
    count_launch();
    check_launch("layer_mega");
}

}  // namespace b200
