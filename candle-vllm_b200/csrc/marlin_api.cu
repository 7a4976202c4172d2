// marlin_api.cu -- C ABI of the GPTQ / Marlin int4 weight-only GEMM: replaces attention_rs::kernels::ffi::
// {gptq_repack, marlin_4bit_f16, marlin_4bit_bf16} (call sites /root/reference/src/backend/gptq.rs:115-178, :313-332;
// host-side preparation /root/reference/src/openai/models/linear.rs:300-413).
//
// gptq_repack turns the GPTQ checkpoint layout (u32 [K/8, N], 8 nibbles along K per word) into this library's
// row-major int4 layout (same number of words; only marlin_4bit_* reads it, so the layout is private to the
// library exactly as Marlin's tile layout is private to attention-rs).  Pre-repacked "marlin" checkpoints
// (checkpoint_format == "marlin", linear.rs:219-220) carry Marlin's own tile order and are NOT accepted.
// The GEMM runs on the tcgen05 dequant pipeline of qmatmul_tc.cu (stream-K into fp32 slabs + a finishing pass to 16 bit).
// Supported: 4-bit symmetric, group size 64 / 128 / -1, no act-order, f16 / bf16, m <= 64, k % 256 == 0 -- the set the
// reference itself repacks to Marlin (linear.rs:319-325).  Anything else records kErrUnsupported.
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

// activation scratch (fp16, K4 order): the reference ABI has no slot for it (its `workspace` is N words of locks),
// so the library owns one buffer per device, grown outside stream capture (the reference warms every shape up
// eagerly before capturing, graph.rs:471-661) or provided once with b200_set_scratch().
static void* g_scratch = nullptr;      // (fp16 activation copy + fp32 partial-sum slabs of the 16-bit-output GEMMs)
static size_t g_scratch_bytes = 0;
static bool g_scratch_owned = false;

void* get_scratch(size_t bytes, cudaStream_t st) {
    if (bytes <= g_scratch_bytes) return g_scratch;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone) {
        set_error(kErrBadArg, "marlin: activation scratch too small (%zu < %zu) during stream capture; run the shape once eagerly or call b200_set_scratch", g_scratch_bytes, bytes);
        return nullptr;
    }
    cudaStreamSynchronize(st);
    if (g_scratch_owned && g_scratch) cudaFree(g_scratch);
    const size_t want = bytes < (4u << 20) ? (4u << 20) : bytes;
    if (cudaMalloc(&g_scratch, want) != cudaSuccess) { g_scratch = nullptr; g_scratch_bytes = 0; set_error(kErrCuda, "marlin: scratch cudaMalloc(%zu) failed", want); return nullptr; }
    g_scratch_bytes = want; g_scratch_owned = true;
    return g_scratch;
}

static void marlin_4bit(const void* x, const void* qweight, const void* scales, const void* qzeros, const void* g_idx, void* out,
                        int m, int k, int n, int group_size, int dtype, int64_t stream) {
    if (m == 0 || n == 0) return;
    B200_REQUIRE(x && qweight && scales && out, kErrBadArg, "marlin_4bit: null pointer");
    B200_REQUIRE(m > 0 && n > 0 && k > 0, kErrBadArg, "marlin_4bit: bad sizes m=%d k=%d n=%d", m, k, n);
    B200_REQUIRE(g_idx == nullptr, kErrUnsupported, "marlin_4bit: act-order (g_idx) is not supported (linear.rs:319-325 never repacks it)");
    (void)qzeros;                                      // symmetric: zero point 8 (the reference passes qzeros but Marlin ignores them)
    B200_REQUIRE(group_size == -1 || group_size == 64 || group_size == 128, kErrUnsupported, "marlin_4bit: group size %d (64, 128, -1)", group_size);
    B200_REQUIRE(k % 256 == 0 && n % 64 == 0, kErrUnsupported, "marlin_4bit: k %% 256 and n %% 64 must be 0 (k=%d n=%d)", k, n);
    B200_REQUIRE(m <= 64, kErrUnsupported, "marlin_4bit: m = %d > 64 (decode batches only in this round)", m);
    cudaStream_t st = as_stream(stream);
    // scratch: fp16 K4 copy of x, then the fp32 partial-sum slabs of the stream-K GEMM
    const size_t x_bytes = ((size_t)m * k * 2 + 255) & ~(size_t)255;
    char* xs = static_cast<char*>(get_scratch(x_bytes + (size_t)wq16_slabs(n, k) * m * n * 4, st));
    if (!xs) return;
    cast(x, xs, (int64_t)m * k, dtype, B200_F16_K4, stream);
    marlin_tc(xs, qweight, scales, out, dtype, m, n, k, group_size, reinterpret_cast<float*>(xs + x_bytes), st);
}

}  // namespace b200

using namespace b200;

extern "C" {

void b200_set_scratch(void* ptr, size_t bytes) {
    if (g_scratch_owned && g_scratch) cudaFree(g_scratch);
    g_scratch = ptr; g_scratch_bytes = bytes; g_scratch_owned = false;
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
    marlin_4bit(x, qweight, scales, qzeros, g_idx, out, m, k, n, group_size, B200_F16, stream);
}
void marlin_4bit_bf16(const void* x, const int32_t* qweight, const void* scales, const void* qzeros, const void* g_idx, void* out,
                      int32_t m, int32_t k, int32_t n, const void* workspace, int32_t group_size, int64_t stream) {
    (void)workspace;
    marlin_4bit(x, qweight, scales, qzeros, g_idx, out, m, k, n, group_size, B200_BF16, stream);
}
void marlin_awq_4bit_f16(const void*, const int32_t*, const void*, const void*, const void*, void*, int32_t, int32_t, int32_t,
                         const void*, int32_t, int64_t) {
    set_error(kErrUnsupported, "marlin_awq_4bit_f16: AWQ (zero-point) int4 is not implemented in this round");
}
void marlin_awq_4bit_bf16(const void*, const int32_t*, const void*, const void*, const void*, void*, int32_t, int32_t, int32_t,
                          const void*, int32_t, int64_t) {
    set_error(kErrUnsupported, "marlin_awq_4bit_bf16: AWQ (zero-point) int4 is not implemented in this round");
}
void awq_repack(const void*, void*, int32_t, int32_t, int32_t, int64_t) {
    set_error(kErrUnsupported, "awq_repack: AWQ int4 is not implemented in this round");
}
void gemm_half_q_half_alt(const void*, const uint32_t*, const uint32_t*, const void*, const int32_t*, void*, int32_t, int32_t, int32_t,
                          int32_t, int64_t) {
    set_error(kErrUnsupported, "gemm_half_q_half_alt: act-order / asymmetric GPTQ is not implemented in this round");
}

}  // extern "C"
