// marlin_api.cu -- C ABI of the GPTQ / Marlin int4 weight-only GEMM: replaces attention_rs::kernels::ffi::
// {gptq_repack, marlin_4bit_f16, marlin_4bit_bf16} (call sites /root/reference/src/backend/gptq.rs:115-178, :313-332;
// host-side preparation /root/reference/src/openai/models/linear.rs:300-413).
//
// gptq_repack turns the GPTQ checkpoint layout (u32 [K/8, N], 8 nibbles along K per word) into this library's
// row-major int4 layout (same number of words; only marlin_4bit_* reads it, so the layout is private to the
// library exactly as Marlin's tile layout is private to attention-rs).  Pre-repacked "marlin" checkpoints
// (checkpoint_format == "marlin", linear.rs:219-220) carry Marlin's own tile order and are NOT accepted.
// The GEMM runs on the tcgen05 dequant pipeline of qmatmul_tc.cu (stream-K into fp32 slabs + a finishing pass to 16 bit).
// Marlin path: 4-bit, group size 64 / 128 / -1, no act-order, f16 / bf16, m <= 64, k % 256 == 0 -- the set the reference itself
// repacks to Marlin (linear.rs:319-325) -- symmetric GPTQ (marlin_4bit_*) and AWQ with zero points (awq_repack +
// marlin_awq_4bit_*, zero points in the layout of examples/convert_awq_marlin.py).  Everything else (act-order, asymmetric GPTQ,
// 8 bit) takes the shape-generic gemm_half_q_half_alt kernel, as in the reference (gptq.rs:182-197).
#include <map>
#include <mutex>
#include <set>
#include <utility>

#include "qmatmul.cuh"

namespace b200 {

// out bytes: row n, 64-k chunk c, byte b: low nibble q[64c + b][n], high nibble q[64c + 32 + b][n]
__global__ void gptq_repack_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int k_packed, int n) {
    const int64_t total = (int64_t)k_packed * n;            // output words: [n][k_packed]
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % n);                       // consecutive threads -> consecutive n (coalesced reads)
        const int wk = (int)(i / n);                        // output word index along k: chunk c = wk / 8, word w = wk % 8
        const int c = wk >> 3, w = wk & 7;
        const int kp_lo = 8 * c + (w >> 1), kp_hi = kp_lo + 4, sh = 16 * (w & 1);
        const uint32_t lo = in[(int64_t)kp_lo * n + col] >> sh, hi = in[(int64_t)kp_hi * n + col] >> sh;
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) o |= (((lo >> (4 * j)) & 0xFu) | (((hi >> (4 * j)) & 0xFu) << 4)) << (8 * j);
        out[(int64_t)col * k_packed + wk] = o;
    }
}

// AWQ checkpoint layout: u32 [K, N/8], nibble i of word (k, j) = q[k][8 j + order[i]], order = [0,2,4,6,1,3,5,7] (AutoAWQ).
// Same private output layout as gptq_repack: [n][k/8 words], byte b of 64-k chunk c: low nibble q[64c + b][n], high q[64c + 32 + b][n].
__global__ void awq_repack_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int k, int n_packed) {
    const int n = n_packed * 8, kw = k / 8;
    const int64_t total = (int64_t)kw * n;                  // output words: [n][k / 8]
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % n);                       // consecutive threads -> consecutive n
        const int wk = (int)(i / n);
        const int c = wk >> 3, w = wk & 7;
        const int sh = 4 * ((0x73625140u >> (4 * (col & 7))) & 7);     // nibble that holds column (col & 7): inverse of the order
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = 4 * w + j;
            const uint32_t lo = (in[(int64_t)(64 * c + b) * n_packed + (col >> 3)] >> sh) & 0xFu;
            const uint32_t hi = (in[(int64_t)(64 * c + 32 + b) * n_packed + (col >> 3)] >> sh) & 0xFu;
            o |= (lo | (hi << 4)) << (8 * j);
        }
        out[(int64_t)col * kw + wk] = o;
    }
}

// Marlin checkpoint format (`B` u32 [K/16, 2N], checkpoint_format == "marlin", /root/reference/src/openai/models/linear.rs:219-251) -> GPTQ packing
// u32 [K/8, N], from which gptq_repack makes this library's layout.  The tile order is the Marlin project's published one (IST-DASLab/marlin,
// `_get_perms` / `Layer.pack`): w[k][n] sits in row k / 16, tile-flat position (n / 16) * 256 + (k % 16) * 16 + n % 16,
// permuted inside every 1024 values by `perm` (inv_perm below, built once on the host) and packed 8 nibbles per word with stride 8.
__constant__ uint16_t c_marlin_inv_perm[1024];
__global__ void marlin_to_gptq_kernel(const uint32_t* __restrict__ B, uint32_t* __restrict__ out, int k, int n) {
    const int64_t total = (int64_t)(k / 8) * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % n), kp = (int)(i / n);
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kk = kp * 8 + j;
            const int64_t t = (int64_t)(col >> 4) * 256 + (kk & 15) * 16 + (col & 15);         // position in the row of 16 x 16 tiles
            const int64_t c = (t & ~(int64_t)1023) + c_marlin_inv_perm[t & 1023];               // after the permutation
            const uint32_t word = B[(int64_t)(kk >> 4) * (2 * n) + (c >> 3)];
            o |= ((word >> (4 * (c & 7))) & 0xFu) << (4 * j);
        }
        out[i] = o;
    }
}
static bool upload_marlin_inv_perm() {
    static std::mutex mu;
    static std::set<int> done;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lk(mu);
    if (done.count(dev)) return true;
    int perm[1024], n = 0;
    for (int i = 0; i < 32; ++i) {
        int perm1[8], c = 0;
        const int col = i / 4;
        for (int block = 0; block < 2; ++block)
            for (int row : {2 * (i % 4), 2 * (i % 4) + 1, 2 * (i % 4 + 4), 2 * (i % 4 + 4) + 1}) perm1[c++] = 16 * row + col + 8 * block;
        for (int j = 0; j < 4; ++j) for (int q = 0; q < 8; ++q) perm[n++] = perm1[q] + 256 * j;
    }
    static const int interleave[8] = {0, 2, 4, 6, 1, 3, 5, 7};
    uint16_t inv[1024];
    for (int g = 0; g < 128; ++g)
        for (int q = 0; q < 8; ++q) inv[perm[g * 8 + interleave[q]]] = (uint16_t)(g * 8 + q);      // res[x] = w[perm'[x]], perm' = interleaved perm
    if (cudaMemcpyToSymbol(c_marlin_inv_perm, inv, sizeof(inv)) != cudaSuccess) return false;
    done.insert(dev);
    return true;
}

// Conventional GPTQ (act-order and / or asymmetric, 4 or 8 bit): the shape-generic path behind gemm_half_q_half_alt
// (call site /root/reference/src/backend/gptq.rs:182-197).  qweight [K / pack, N] packed along K, qzeros [G, N / pack] packed
// along N and stored minus one (GPTQ v1), scales f16 [G, N], group of row k = g_idx[k].  One thread per output column
// (coalesced weight words), 8 activation rows per pass, fp32 accumulation, f16 output.
template <int kBits>
__global__ void __launch_bounds__(128)
gptq_alt_kernel(const __half* __restrict__ x, const uint32_t* __restrict__ qw, const uint32_t* __restrict__ qz, const __half* __restrict__ sc,
                const int32_t* __restrict__ g_idx, __half* __restrict__ out, int m, int n, int k) {
    constexpr int kPack = 32 / kBits;
    constexpr uint32_t kMask = (1u << kBits) - 1u;
    const int col = blockIdx.x * 128 + threadIdx.x;
    const int m0 = blockIdx.y * 8;
    if (col >= n) return;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    int g_prev = -1;
    float s = 0.f, z = 0.f;
    for (int kp = 0; kp < k / kPack; ++kp) {
        const uint32_t word = qw[(int64_t)kp * n + col];
#pragma unroll
        for (int j = 0; j < kPack; ++j) {
            const int kk = kp * kPack + j;
            const int g = g_idx[kk];
            if (g != g_prev) {                               // act-order: the group can change at any row
                g_prev = g;
                s = __half2float(sc[(int64_t)g * n + col]);
                z = (float)(((qz[(int64_t)g * (n / kPack) + col / kPack] >> (kBits * (col % kPack))) & kMask) + 1u);
            }
            const float wv = ((float)((word >> (kBits * j)) & kMask) - z) * s;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (m0 + i < m) acc[i] += __half2float(x[(int64_t)(m0 + i) * k + kk]) * wv;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (m0 + i < m) out[(int64_t)(m0 + i) * n + col] = __float2half_rn(acc[i]);
}

// activation scratch (fp16, K4 order) + fp32 partial-sum slabs: the reference ABI has no slot for it (its `workspace` is N
// words of locks), so the library owns one buffer per (device, stream) -- two streams or two devices (threaded.rs runs one
// thread per rank in one process) never share scratch -- grown outside stream capture (the reference warms every shape up
// eagerly before capturing, graph.rs:471-661), or one caller-provided buffer per device set with b200_set_scratch().
struct Scratch { void* ptr = nullptr; size_t bytes = 0; bool owned = false; };
static std::mutex g_scratch_mu;
static std::map<std::pair<int, cudaStream_t>, Scratch> g_scratch;      // library-owned, per (device, stream)
static std::map<int, Scratch> g_user_scratch;                          // b200_set_scratch, per device

void* get_scratch(size_t bytes, cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error(kErrCuda, "scratch: no current device"); return nullptr; }
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    auto u = g_user_scratch.find(dev);
    if (u != g_user_scratch.end() && u->second.ptr) {
        if (bytes <= u->second.bytes) return u->second.ptr;
        set_error(kErrBadArg, "scratch: the buffer given to b200_set_scratch holds %zu bytes, %zu needed", u->second.bytes, bytes);
        return nullptr;
    }
    Scratch& sc = g_scratch[{dev, st}];
    if (bytes <= sc.bytes) return sc.ptr;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone) {
        set_error(kErrBadArg, "marlin: activation scratch too small (%zu < %zu) during stream capture; run the shape once eagerly or call b200_set_scratch", sc.bytes, bytes);
        return nullptr;
    }
    cudaStreamSynchronize(st);                               // nothing on this stream still reads the old buffer
    if (sc.ptr) cudaFree(sc.ptr);
    const size_t want = bytes < (4u << 20) ? (4u << 20) : bytes;
    if (cudaMalloc(&sc.ptr, want) != cudaSuccess) { sc = Scratch{}; cudaGetLastError(); set_error(kErrCuda, "marlin: scratch cudaMalloc(%zu) failed", want); return nullptr; }
    sc.bytes = want; sc.owned = true;
    return sc.ptr;
}

static void marlin_4bit(const void* x, const void* qweight, const void* scales, const void* qzeros, const void* g_idx, void* out,
                        int m, int k, int n, int group_size, int dtype, bool awq, int64_t stream) {
    if (m == 0 || n == 0) return;
    B200_REQUIRE(x && qweight && scales && out, kErrBadArg, "marlin_4bit: null pointer");
    B200_REQUIRE(!awq || qzeros, kErrBadArg, "marlin_awq_4bit: qzeros (marlin zero-point layout, examples/convert_awq_marlin.py) is required");
    B200_REQUIRE(m > 0 && n > 0 && k > 0, kErrBadArg, "marlin_4bit: bad sizes m=%d k=%d n=%d", m, k, n);
    // g_idx: the reference hands the checkpoint's g_idx to this symbol for EVERY quant_method == "gptq" layer
    // (linear.rs:298-337 loads it as Some(..) even with desc_act = false; gptq.rs:27-35,139-152 forwards the pointer), and only
    // repacks to Marlin when desc_act is false (linear.rs:319-325), where g_idx is the trivial k / group_size sequence.  Marlin
    // never reads it on that path and neither do we (reading a device array here would force a host sync and break graph
    // capture); real act-order checkpoints take gemm_half_q_half_alt, as in the reference (gptq.rs:182-197).
    (void)g_idx;
    if (!awq) qzeros = nullptr;                        // symmetric: zero point 8 (the reference passes qzeros but Marlin ignores them)
    B200_REQUIRE(group_size == -1 || group_size == 64 || group_size == 128, kErrUnsupported, "marlin_4bit: group size %d (64, 128, -1)", group_size);
    B200_REQUIRE(k % 256 == 0 && n % 64 == 0, kErrUnsupported, "marlin_4bit: k %% 256 and n %% 64 must be 0 (k=%d n=%d)", k, n);
    cudaStream_t st = as_stream(stream);
    // scratch: fp16 K4 copy of (up to 64 rows of) x, then the fp32 partial-sum slabs of the stream-K GEMM.  More than 64 rows
    // (prefill chunks) run 64 at a time, stream-ordered on the same scratch.
    const int mc = m < 64 ? m : 64;
    const size_t x_bytes = ((size_t)mc * k * 2 + 255) & ~(size_t)255;
    char* xs = static_cast<char*>(get_scratch(x_bytes + (size_t)wq16_slabs(n, k) * mc * n * 4, st));
    if (!xs) return;
    const size_t esz = 2;                                            // f16 / bf16
    for (int m0 = 0; m0 < m; m0 += 64) {
        const int mm = m - m0 < 64 ? m - m0 : 64;
        cast(static_cast<const char*>(x) + (size_t)m0 * k * esz, xs, (int64_t)mm * k, dtype, B200_F16_K4, stream);
        marlin_tc(xs, qweight, scales, qzeros, static_cast<char*>(out) + (size_t)m0 * n * esz, dtype, mm, n, k, group_size,
                  reinterpret_cast<float*>(xs + x_bytes), st);
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

void b200_set_scratch(void* ptr, size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error(kErrCuda, "b200_set_scratch: no current device"); return; }
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    if (ptr && bytes) g_user_scratch[dev] = Scratch{ptr, bytes, false};
    else g_user_scratch.erase(dev);
}

void marlin_checkpoint_repack(const void* marlin_b, void* out, void* scratch_gptq, int32_t k, int32_t n, int64_t stream) {
    B200_REQUIRE(marlin_b && out && scratch_gptq && k > 0 && n > 0, kErrBadArg, "marlin_checkpoint_repack: bad arguments");
    B200_REQUIRE(k % 64 == 0 && n % 64 == 0, kErrUnsupported, "marlin_checkpoint_repack: k and n must be multiples of 64 (k=%d n=%d)", k, n);
    B200_REQUIRE(upload_marlin_inv_perm(), kErrCuda, "marlin_checkpoint_repack: permutation upload failed");
    const int64_t total = (int64_t)(k / 8) * n;
    int64_t g = (total + 255) / 256;
    if (g > (int64_t)sm_count() * 16) g = (int64_t)sm_count() * 16;
    marlin_to_gptq_kernel<<<(int)g, 256, 0, as_stream(stream)>>>((const uint32_t*)marlin_b, (uint32_t*)scratch_gptq, k, n);
    count_launch();
    if (!check_launch("marlin_checkpoint_repack")) return;
    gptq_repack(scratch_gptq, out, k / 8, n, stream);
}

void gptq_repack(const void* in, void* out, int32_t k_packed, int32_t n, int64_t stream) {
    B200_REQUIRE(in && out && k_packed > 0 && n > 0, kErrBadArg, "gptq_repack: bad arguments");
    B200_REQUIRE(k_packed % 8 == 0, kErrUnsupported, "gptq_repack: K must be a multiple of 64 (k_packed=%d)", k_packed);
    const int64_t total = (int64_t)k_packed * n;
    int64_t g = (total + 255) / 256;
    if (g > (int64_t)sm_count() * 16) g = (int64_t)sm_count() * 16;
    gptq_repack_kernel<<<(int)g, 256, 0, as_stream(stream)>>>((const uint32_t*)in, (uint32_t*)out, k_packed, n);
    count_launch();
    check_launch("gptq_repack");
}

void marlin_4bit_f16(const void* x, const int32_t* qweight, const void* scales, const void* qzeros, const void* g_idx, void* out,
                     int32_t m, int32_t k, int32_t n, const void* workspace, int32_t group_size, int64_t stream) {
    (void)workspace;
    marlin_4bit(x, qweight, scales, qzeros, g_idx, out, m, k, n, group_size, B200_F16, false, stream);
}
void marlin_4bit_bf16(const void* x, const int32_t* qweight, const void* scales, const void* qzeros, const void* g_idx, void* out,
                      int32_t m, int32_t k, int32_t n, const void* workspace, int32_t group_size, int64_t stream) {
    (void)workspace;
    marlin_4bit(x, qweight, scales, qzeros, g_idx, out, m, k, n, group_size, B200_BF16, false, stream);
}
void marlin_awq_4bit_f16(const void* x, const int32_t* qweight, const void* scales, const void* qzeros, const void* g_idx, void* out,
                         int32_t m, int32_t k, int32_t n, const void* workspace, int32_t group_size, int64_t stream) {
    (void)workspace;
    marlin_4bit(x, qweight, scales, qzeros, g_idx, out, m, k, n, group_size, B200_F16, true, stream);
}
void marlin_awq_4bit_bf16(const void* x, const int32_t* qweight, const void* scales, const void* qzeros, const void* g_idx, void* out,
                          int32_t m, int32_t k, int32_t n, const void* workspace, int32_t group_size, int64_t stream) {
    (void)workspace;
    marlin_4bit(x, qweight, scales, qzeros, g_idx, out, m, k, n, group_size, B200_BF16, true, stream);
}
void awq_repack(const void* in, void* out, int32_t k, int32_t n_packed, int32_t bits, int64_t stream) {
    B200_REQUIRE(in && out && k > 0 && n_packed > 0, kErrBadArg, "awq_repack: bad arguments");
    B200_REQUIRE(bits == 4, kErrUnsupported, "awq_repack: %d-bit AWQ (4 only)", bits);
    B200_REQUIRE(k % 64 == 0, kErrUnsupported, "awq_repack: K must be a multiple of 64 (k=%d)", k);
    const int64_t total = (int64_t)(k / 8) * n_packed * 8;
    int64_t g = (total + 255) / 256;
    if (g > (int64_t)sm_count() * 16) g = (int64_t)sm_count() * 16;
    awq_repack_kernel<<<(int)g, 256, 0, as_stream(stream)>>>((const uint32_t*)in, (uint32_t*)out, k, n_packed);
    count_launch();
    check_launch("awq_repack");
}
void gemm_half_q_half_alt(const void* x, const uint32_t* qweight, const uint32_t* qzeros, const void* scales, const int32_t* g_idx, void* out,
                          int32_t m, int32_t n, int32_t k, int32_t bits, int64_t stream) {
    if (m == 0 || n == 0) return;
    B200_REQUIRE(x && qweight && qzeros && scales && out && m > 0 && n > 0 && k > 0, kErrBadArg, "gemm_half_q_half_alt: bad arguments");
    B200_REQUIRE(g_idx, kErrBadArg, "gemm_half_q_half_alt: g_idx is required (the group of every weight row)");
    B200_REQUIRE(bits == 4 || bits == 8, kErrUnsupported, "gemm_half_q_half_alt: %d-bit weights (4 or 8)", bits);
    B200_REQUIRE(k % (32 / bits) == 0 && n % (32 / bits) == 0, kErrBadArg, "gemm_half_q_half_alt: k and n must be multiples of %d", 32 / bits);
    const dim3 grid(ceil_div(n, 128), ceil_div(m, 8));
    if (bits == 4)
        gptq_alt_kernel<4><<<grid, 128, 0, as_stream(stream)>>>((const __half*)x, qweight, qzeros, (const __half*)scales, g_idx, (__half*)out, m, n, k);
    else
        gptq_alt_kernel<8><<<grid, 128, 0, as_stream(stream)>>>((const __half*)x, qweight, qzeros, (const __half*)scales, g_idx, (__half*)out, m, n, k);
    count_launch();
    check_launch("gemm_half_q_half_alt");
}

}  // extern "C"
