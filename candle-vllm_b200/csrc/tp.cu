// tp.cu -- tensor-parallel collectives on the residual path (AllReduce after the row-parallel
// o_proj / down_proj: /root/reference/src/openai/distributed.rs:547-654, :696-710).
// NCCL is resolved at run time from the already-loaded torch-bundled libnccl.so.2 (the host
// process creates the communicator); no link-time dependency.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"
#include "mega.cuh"

namespace b200 {

typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
static nccl_allreduce_fn g_allreduce = nullptr;
static nccl_allgather_fn g_allgather = nullptr;

static bool resolve() {
    if (g_allreduce && g_allgather) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    g_allreduce = reinterpret_cast<nccl_allreduce_fn>(h ? dlsym(h, "ncclAllReduce") : dlsym(RTLD_DEFAULT, "ncclAllReduce"));
    g_allgather = reinterpret_cast<nccl_allgather_fn>(h ? dlsym(h, "ncclAllGather") : dlsym(RTLD_DEFAULT, "ncclAllGather"));
    return g_allreduce != nullptr && g_allgather != nullptr;
}

// VocabParallelLinear + AllGather (/root/reference/src/openai/distributed.rs:1632-1667) reduced to what greedy
// decoding needs: every rank contributes (max logit, global index) per sequence; gather 8 bytes per sequence.
void tp_allgather_bytes(void* comm, const void* src, void* dst, size_t bytes_per_rank, cudaStream_t st) {
    if (!comm) { set_error(kErrBadArg, "tp_allgather: no communicator"); return; }
    if (!resolve()) { set_error(kErrUnsupported, "tp_allgather: ncclAllGather not found (libnccl.so.2 not loaded)"); return; }
    const int rc = g_allgather(src, dst, bytes_per_rank, /*ncclInt8*/ 0, comm, st);
    if (rc != 0) set_error(kErrCuda, "tp_allgather: ncclAllGather rc=%d", rc);
}

void tp_allreduce_f32(void* comm, float* buf, int64_t n, cudaStream_t st) {
    if (!comm) { set_error(kErrBadArg, "tp_allreduce: no communicator"); return; }
    if (!resolve()) { set_error(kErrUnsupported, "tp_allreduce: ncclAllReduce not found (libnccl.so.2 not loaded)"); return; }
    // ncclFloat32 = 7, ncclSum = 0
    const int rc = g_allreduce(buf, buf, (size_t)n, 7, 0, comm, st);
    if (rc != 0) set_error(kErrCuda, "tp_allreduce: ncclAllReduce rc=%d", rc);
}

// ---- fused all-reduce over NVLink peer memory ------------------------------------------------------------------------
// Replaces {zero partial, GEMM, ncclAllReduce, add, rms_norm} of the row-parallel linears by {GEMM, this kernel}
// (reference semantics: AllReduce::cuda_fwd + residual add + RmsNorm, /root/reference/src/openai/distributed.rs:572-653,
// quantized_llama.rs:470-488).  Two hops over NVLink / NVSwitch, no fences, no separate flags:
//   * token row r is OWNED by rank r % world: only the owner keeps the fp32 residual stream of that row;
//   * hop 1 (reduce): every other rank pushes its partial row into the owner's gather slot as {value, epoch} 8-byte words
//     (NCCL's "LL" idea: an aligned 8-byte store is atomic, so the epoch travelling WITH the datum is the ready flag --
//     a measured fence + flag protocol cost two extra NVLink round trips per all-reduce);
//   * the owner polls the words, adds the partials in rank order (one fp32 result per row, bitwise identical for every
//     consumer), adds the residual, applies the NEXT RMSNorm and
//   * hop 2 (broadcast): pushes the normalised fp16 row (what the next GEMM consumes, K4 order) to every rank as
//     {half2, epoch} words.  Non-owners poll their copy.  Per all-reduce a rank sends (w-1)/w x 32 KB/row out and 16 KB/row x
//     (w-1) for its own rows: 1.35 MB at world 8, B = 32, against 3.6 MB for a one-shot exchange.
// Buffers are double-buffered by epoch parity: a rank cannot be two all-reduces ahead of a peer because each one needs that
// peer's words.  One CTA per row; every CTA of a launch is resident (rows <= 148).

__device__ __forceinline__ void st_ll(void* addr, uint32_t a, uint32_t b, uint32_t e) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(a), "r"(e), "r"(b), "r"(e) : "memory");
}
__device__ __forceinline__ uint4 ld_ll(const void* addr) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(addr) : "memory");
    return v;
}

// bounded spin: a peer that never shows up (crashed rank) must not wedge the GPU -- after `timeout_ns` (B200_TP_TIMEOUT_MS,
// default 120 s: long enough for a peer that is lazily capturing a graph, paused in a debugger or collecting garbage) the
// row gives up, poisons its output with NaN and raises the timeout word.  The word lives in host-mapped pinned memory, so
// b200_llama_decode / read_* see it right after their stream sync and report an error instead of returning token 0.
__device__ unsigned long long g_tp_timeout_ns = 120000000000ull;
__device__ __forceinline__ bool ll_wait(const void* addr, uint32_t e, uint4& w, volatile uint32_t* timeout_word) {
    w = ld_ll(addr);
    if (w.y == e && w.w == e) return true;
    unsigned long long t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        w = ld_ll(addr);
        if (w.y == e && w.w == e) return true;
        if ((spins & 0x3ffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > g_tp_timeout_ns) { *timeout_word = 1u; __threadfence_system(); w = make_uint4(0x7fc00000u, e, 0x7fc00000u, e); return false; }
        }
    }
}

using InboxLayout = TpInboxLayout;

__global__ void __launch_bounds__(256)
tp_allreduce_add_norm_kernel(float* __restrict__ partial, float* __restrict__ x, const float* __restrict__ norm_w, __half* __restrict__ xn,
                             const __grid_constant__ PeerSet peers, int rank, int world, int n, int rows_max, float eps,
                             uint32_t* __restrict__ timeout_word) {
    pdl_wait();
    pdl_trigger();
    constexpr int kMaxIt = 8;                                 // rows up to 8192 columns stay in registers
    const int row = blockIdx.x, nv = n >> 2;
    const InboxLayout lay(world, rows_max, n);
    char* const mine = peers.p[rank];
    uint32_t* epoch = reinterpret_cast<uint32_t*>(mine + lay.epoch_off);
    const uint32_t e = epoch[row] + 1u;
    __syncthreads();                                          // everyone has read the epoch before thread 0 may bump it at the end
    const int par = (int)(e & 1u);
    const int owner = row % world, lrow = row / world;
    // my partial -> registers (and leave the accumulator zeroed for the next split-K GEMM)
    float4* pr = reinterpret_cast<float4*>(partial + (int64_t)row * n);
    float4 v[kMaxIt];
#pragma unroll
    for (int it = 0; it < kMaxIt; ++it) {
        const int i = threadIdx.x + it * 256;
        if (i < nv) { v[it] = pr[i]; pr[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
    }
    __half* o = xn + (int64_t)row * n;
    if (rank != owner) {
        // hop 1: my partial -> the owner's gather slot [par][rank][lrow]
        char* dst = peers.p[owner] + (((size_t)par * world + rank) * lay.rows_owned + lrow) * n * 8;
#pragma unroll
        for (int it = 0; it < kMaxIt; ++it) {
            const int i = threadIdx.x + it * 256;
            if (i < nv) {
                st_ll(dst + (size_t)i * 32, __float_as_uint(v[it].x), __float_as_uint(v[it].y), e);
                st_ll(dst + (size_t)i * 32 + 16, __float_as_uint(v[it].z), __float_as_uint(v[it].w), e);
            }
        }
        // hop 2: wait for the owner's normalised row in my broadcast slot [par][row]
        const char* src = mine + lay.bcast_off + ((size_t)par * rows_max + row) * (n / 2) * 8;
#pragma unroll
        for (int it = 0; it < kMaxIt; ++it) {
            const int i = threadIdx.x + it * 256;            // 4 halves = 2 words = one 16-byte LL pair
            if (i < nv) {
                uint4 w;
                ll_wait(src + (size_t)i * 16, e, w, timeout_word);
                *reinterpret_cast<uint2*>(o + 4 * i) = make_uint2(w.x, w.z);
            }
        }
    } else {
        // owner: gather the peers' partials (rank order -> one well-defined fp32 sum), residual add, RMSNorm
        float4* xr = reinterpret_cast<float4*>(x + (int64_t)row * n);
        float ss = 0.f;
#pragma unroll
        for (int it = 0; it < kMaxIt; ++it) {
            const int i = threadIdx.x + it * 256;
            if (i < nv) {
                // all peers' words for these four columns in flight together; only stale ones are re-read (polling them one peer after
                // the other made the exchange grow with the world size: 11 us at 2 ranks, 23 us at 8)
                uint4 w0[8], w1[8];
                uint32_t have = 1u << rank;
                const uint32_t all = (1u << world) - 1u;
                unsigned long long t0 = 0;
                for (uint32_t spins = 0; have != all; ++spins) {
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        if (p < world && !(have >> p & 1u)) {
                            const char* src = mine + (((size_t)par * world + p) * lay.rows_owned + lrow) * n * 8 + (size_t)i * 32;
                            w0[p] = ld_ll(src);
                            w1[p] = ld_ll(src + 16);
                        }
                    }
#pragma unroll
                    for (int p = 0; p < 8; ++p)
                        if (p < world && !(have >> p & 1u) && w0[p].y == e && w0[p].w == e && w1[p].y == e && w1[p].w == e) have |= 1u << p;
                    if (have != all && (spins & 0x3ffu) == 0x3ffu) {
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > g_tp_timeout_ns) {           // a peer never showed up: poison the row, raise the host-visible flag
                            *timeout_word = 1u; __threadfence_system();
#pragma unroll
                            for (int p = 0; p < 8; ++p)
                                if (p < world && !(have >> p & 1u)) { w0[p] = make_uint4(0x7fc00000u, e, 0x7fc00000u, e); w1[p] = w0[p]; }
                            have = all;
                        }
                    }
                }
                float4 a = xr[i];
#pragma unroll
                for (int p = 0; p < 8; ++p) {                 // rank order: one well-defined fp32 sum
                    if (p >= world) continue;
                    if (p == rank) { a.x += v[it].x; a.y += v[it].y; a.z += v[it].z; a.w += v[it].w; continue; }
                    a.x += __uint_as_float(w0[p].x); a.y += __uint_as_float(w0[p].z); a.z += __uint_as_float(w1[p].x); a.w += __uint_as_float(w1[p].z);
                }
                xr[i] = a;
                v[it] = a;
                ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
            }
        }
        __shared__ float red[8];
        ss = warp_sum(ss);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[i];
        const float sc = rsqrtf(tot / (float)n + eps);
        const float4* wr = reinterpret_cast<const float4*>(norm_w);
#pragma unroll
        for (int it = 0; it < kMaxIt; ++it) {
            const int i = threadIdx.x + it * 256;
            if (i < nv) {
                const float4 g = __ldg(wr + i);
                // K4 order: the middle two of every four swapped (include/b200_backend.h)
                const __half2 p0 = __floats2half2_rn(v[it].x * sc * g.x, v[it].z * sc * g.z), p1 = __floats2half2_rn(v[it].y * sc * g.y, v[it].w * sc * g.w);
                const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&p0), u1 = *reinterpret_cast<const uint32_t*>(&p1);
                *reinterpret_cast<uint2*>(o + 4 * i) = make_uint2(u0, u1);
                // hop 2: broadcast to every other rank's slot [par][row]
                for (int p = 0; p < world; ++p)
                    if (p != rank) st_ll(peers.p[p] + lay.bcast_off + ((size_t)par * rows_max + row) * (n / 2) * 8 + (size_t)i * 16, u0, u1, e);
            }
        }
    }
    if (threadIdx.x == 0) epoch[row] = e;
}

size_t tp_peer_inbox_bytes(int world, int rows_max, int n) { return InboxLayout(world, rows_max, n).total; }
void tp_set_timeout_ms(long long ms) {
    const unsigned long long ns = (unsigned long long)(ms > 0 ? ms : 1) * 1000000ull;
    cudaMemcpyToSymbol(g_tp_timeout_ns, &ns, sizeof(ns));
}

void tp_allreduce_add_norm(float* partial, float* x, const float* norm_w, void* xn_f16_k4, void* const* peers, int rank, int world,
                           int rows, int n, int rows_max, float eps, uint32_t* timeout_word, cudaStream_t st) {
    if (world < 2 || world > 8 || n % 4 || n > 8192 || rows > 128) { set_error(kErrUnsupported, "tp_allreduce_add_norm: world %d, n %d, rows %d", world, n, rows); return; }
    PeerSet ps{};
    for (int i = 0; i < world; ++i) ps.p[i] = static_cast<char*>(peers[i]);
    launch_pdl(tp_allreduce_add_norm_kernel, dim3(rows), dim3(256), 0, st, partial, x, norm_w, static_cast<__half*>(xn_f16_k4), ps, rank, world, n,
               rows_max, eps, timeout_word);
    count_launch();
    check_launch("tp_allreduce_add_norm");
}

}  // namespace b200

// ---- CUDA IPC plumbing for the peer inboxes (one process per GPU) --------------------------------------------------------------
extern "C" void* b200_ipc_alloc(size_t bytes, void* handle_out) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { b200::set_error(b200::kErrCuda, "b200_ipc_alloc: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError())); return nullptr; }
    cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { b200::set_error(b200::kErrCuda, "b200_ipc_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); cudaFree(p); cudaGetLastError(); return nullptr; }
    memcpy(handle_out, &h, sizeof(h));
    cudaDeviceSynchronize();
    return p;
}
extern "C" void* b200_ipc_open(const void* handle) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { b200::set_error(b200::kErrCuda, "b200_ipc_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void b200_ipc_close(void* p) { if (p) cudaIpcCloseMemHandle(p); }
extern "C" void b200_ipc_free(void* p) { if (p) cudaFree(p); }

extern "C" void b200_allreduce_f32(void* nccl_comm, float* buf, int64_t n, int64_t stream) {
    b200::tp_allreduce_f32(nccl_comm, buf, n, b200::as_stream(stream));
}
