// engine.cu -- host-side decode engine: GGUF-LLaMA forward for one decode step over the paged KV
// cache, captured once per batch size into a CUDA graph and replayed.
//
// Mirrors the reference's host code for this path:
//   GGUFLLaMa::forward_inner   /root/reference/src/openai/models/quantized_llama.rs:424-506
//   QuantizedAttention::forward /root/reference/src/openai/models/layers/attention.rs:910-1011
//   Mlp::forward               /root/reference/src/openai/models/quantized_llama.rs:32-44
//   GraphCapturer::capture/replay /root/reference/src/backend/graph.rs:471-661, :685-803
//     (static input buffers, metadata copied in, one cuGraphLaunch, stream sync, narrow output)
//   prepare_decode metadata    /root/reference/src/openai/pipelines/inputs.rs:376-454
//
// Dtype flow (same cast points as the reference): f32 residual stream; QMatMul activations are
// fp16 (include/b200_backend.h); q,k,v -> bf16 before the cache write / attention; attention output
// rounded to bf16, handed to wo as fp16 (exact: bf16 subset of fp16 range here); logits f32.
// The residual adds are fused into the wo / w2 GEMM epilogues (accumulate into x).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "attention.cuh"
#include "mega.cuh"
#include "qmatmul.cuh"

using namespace b200;

extern "C" long long b200_total_kernel_launches(void);
namespace b200 {
void tp_allreduce_f32(void* comm, float* buf, int64_t n, cudaStream_t st);
void tp_allgather_bytes(void* comm, const void* src, void* dst, size_t bytes_per_rank, cudaStream_t st);
size_t tp_peer_inbox_bytes(int world, int rows_max, int n);
void tp_set_timeout_ms(long long ms);
void tp_allreduce_add_norm(float* partial, float* x, const float* norm_w, void* xn_f16_k4, void* const* peers, int rank, int world,
                           int rows, int n, int rows_max, float eps, uint32_t* timeout_word, cudaStream_t st);
void rope_and_cache_slabs(float* qkv, const SlabInfo& si, void* q_out, void* key_cache, void* value_cache, const float* cos_t, const float* sin_t,
                          const int64_t* positions, const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads, int32_t num_kv_heads,
                          int32_t head_dim, int32_t interleaved, int32_t dtype, int32_t cache_dtype, int64_t stream);
void gather_logits_transpose(const float* gathered, float* out, int world, int rows, int vocab_l, int vocab, cudaStream_t st);
void argmax_pairs(const float* logits, void* pairs, int rows, int ld, int n, int chunks, int index_offset, cudaStream_t st);
void argmax_reduce_pairs(const void* gathered, int32_t* out, int rows, int world, cudaStream_t st);
}

struct b200_llama {
    b200_llama_config cfg;
    std::vector<b200_llama_layer_ex> layers;
    const float* tok_embeddings = nullptr;
    const float* norm = nullptr;
    b200_linear output{};                     // lm_head
    int rope_neox = 0;
    std::vector<void*> kc, vc;
    int64_t num_blocks = 0;
    void* comm = nullptr;
    std::vector<void*> peers;                 // every rank's inbox (CUDA IPC mappings; peers[tp_rank] is local): fused all-reduce

    // local (tensor-parallel shard) sizes
    int heads_l, kv_l, ffn_l, vocab_l, qkv_row;

    // static device buffers (graph.rs: static input buffers sized for max batch)
    char* d_meta = nullptr;                                      // one slab: tokens | positions | slots | ctx | tables
    int64_t* d_tokens = nullptr; int64_t* d_positions = nullptr; int64_t* d_slots = nullptr;
    uint32_t* d_ctx = nullptr; uint32_t* d_tables = nullptr;
    float* x = nullptr; __half* xn = nullptr; float* qkv = nullptr; __nv_bfloat16* q16 = nullptr;
    __half* attn16 = nullptr; float* gate = nullptr; float* up = nullptr; __half* act16 = nullptr;
    float* partial = nullptr; float* logits = nullptr; int32_t* next_tokens = nullptr;
    float* tp_pairs = nullptr; float* tp_gathered = nullptr;      // (max, index) per sequence: local, and gathered over ranks
    float* cos_t = nullptr; float* sin_t = nullptr;
    void* attn_ws = nullptr; size_t attn_ws_bytes = 0;

    // pinned staging for the host metadata: two slabs used alternately, each guarded by an event recorded after its H2D copy, so
    // a caller that does not read results back (both host pointers NULL) can queue the next step without overwriting a slab
    // whose copy is still in flight
    char* h_stage[2] = {nullptr, nullptr}; cudaEvent_t stage_ev[2] = {nullptr, nullptr}; int stage_idx = 0;
    size_t stage_bytes = 0;
    int32_t* h_next = nullptr;
    // fused all-reduce give-up flag: host-mapped pinned word written by the kernel (tp.cu), read by the host after a stream sync
    uint32_t* h_timeout = nullptr; uint32_t* d_timeout = nullptr;
    // vocab-parallel lm_head (distributed.rs:1448-1454): vocab padded to 64 and split; gathered logits on request
    int vocab_pad = 0;
    float* logits_gathered = nullptr; float* logits_full = nullptr;

    // persistent layer kernel (layer_mega.cu): per-tile partial-sum slabs of the GEMM phases and the grid-wide counters
    bool use_mega = false;
    int mega_mode = 0;             // B200_MEGA: 0 one launch per GEMM with atomics (legacy), 1 fused layer kernel, 2 split deterministic
    long long tp_timeout_ms = 120000;
    int mega_G = 0, s_qkv = 1, s_ro = 1, s_gu = 1;                // slabs per buffer (max CTAs sharing one tile)
    float* qkv_slabs = nullptr; float* ro_slabs = nullptr; float* gate_slabs = nullptr; float* up_slabs = nullptr;
    uint32_t* mega_counters = nullptr; size_t mega_counter_bytes = 0;
    long long* mega_trace = nullptr; int mega_trace_launch = -1;   // B200_MEGA_TRACE=<launch index>: clock64 stamps of that launch

    std::map<int, cudaGraphExec_t> graphs;       // batch size -> captured step
    std::map<int, int> launches_per_step;
    int64_t launches = 0;
    bool ok = false;
};

namespace {

template <typename T>
bool dmalloc(T*& p, size_t n) {
    void* q = nullptr;
    if (cudaMalloc(&q, n * sizeof(T) + 256) != cudaSuccess) { set_error(kErrCuda, "engine: cudaMalloc(%zu) failed", n * sizeof(T)); return false; }
    cudaMemset(q, 0, n * sizeof(T) + 256);
    p = static_cast<T*>(q);
    return true;
}

// next-step metadata on the device (fixed block tables): tokens <- argmax, pos += 1, ctx += 1,
// slot = table[pos / bs] * bs + pos % bs   (inputs.rs:410-423)
__global__ void advance_metadata_kernel(int64_t* tokens, const int32_t* next_tokens, int64_t* positions,
                                        int64_t* slots, uint32_t* ctx, const uint32_t* tables,
                                        int num_seqs, int max_blocks, int block_size, int feed_tokens) {
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_seqs) return;
    if (feed_tokens) tokens[b] = next_tokens[b];
    const int64_t pos = positions[b] + 1;
    positions[b] = pos;
    ctx[b] += 1;
    // a sequence that outgrows its table row gets the pad slot (no cache write) instead of an out-of-bounds table read
    const int64_t blk = pos / block_size;
    slots[b] = blk < max_blocks ? (int64_t)tables[(int64_t)b * max_blocks + blk] * block_size + pos % block_size : -1;
    if (blk >= max_blocks) ctx[b] = (uint32_t)max_blocks * (uint32_t)block_size;
}

__global__ void zero_f32_kernel(float* p, int64_t n) {
    pdl_wait();
    pdl_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

// One decode forward on `st` for B sequences.  Returns number of kernel launches issued.
// ---- one linear of the decode engine: y[B, n] (f32, row pitch ldy) (+)= x16[B, k] . W[n, k]^T, whatever the weight kind ---------------
// Activation format by kind (all linears of a model share it, checked in b200_llama_set_layer_ex): GGML / MARLIN4 read fp16 in K4
// order, DENSE16 reads the weights' dtype in natural order.  accumulate = 0 with a GGML weight may still split tiles over K and
// red.add: y must be zero then (the engine's consumers leave their inputs zeroed).
int act_format(const b200_llama* m) {
    const b200_linear& l = m->layers[0].wq;
    return l.kind == B200_LIN_DENSE16 ? l.type : B200_F16_K4;
}
void linear(b200_llama* m, const b200_linear& l, const void* x16, float* y, int64_t ldy, int B, int n, int k, int accumulate, cudaStream_t st) {
    switch (l.kind) {
        case B200_LIN_GGML: qmatmul_dispatch(x16, l.w, y, ldy, B, n, k, l.type, accumulate, st); break;
        case B200_LIN_MARLIN4: {
            const void* zs[1] = {l.zeros};
            marlin_tc_f32_multi(x16, 1, &l.w, &l.scales, l.type == B200_BF16, l.zeros ? zs : nullptr, &y, &n, ldy, B, k, l.group_size, accumulate, st);
            break;
        }
        case B200_LIN_DENSE16: dense_gemm_16(x16, l.w, nullptr, y, B, n, k, k, k, ldy, l.type, B200_F32, st, accumulate, /*allow_split_k=*/1); break;
        default: set_error(kErrUnsupported, "engine: linear kind %d", l.kind);
    }
}
// weight bytes of a linear the tcgen05 dequant-GEMM streams (0: another kernel serves it) -- for the L2 prefetch hint
size_t lin_stream_bytes(const b200_linear& l, int n, int k) {
    if (l.kind == B200_LIN_GGML) return l.type == B200_GGML_Q4_K ? (size_t)n * (k / 256) * 144 : (l.type == B200_GGML_Q6_K ? (size_t)n * (k / 256) * 210 : 0);
    return 0;       // int4: measured slightly slower with the hint (5.81 vs 5.74 ms per step, config 3) -- its scale loads already queue at the L2
}
void prefetch_next(std::initializer_list<std::pair<const b200_linear*, std::pair<int, int>>> next) {
    const void* ptrs[3]; size_t bytes[3]; int n = 0;
    for (const auto& e : next) { if (n == 3) break; ptrs[n] = e.first->w; bytes[n] = lin_stream_bytes(*e.first, e.second.first, e.second.second); if (bytes[n]) ++n; }
    qmatmul_tc_prefetch_next(n, ptrs, bytes);
}
// several int4 linears over the same activations (QKV, gate|up) as one launch when they share group size, scale dtype and zero-point form
bool marlin_fusable(std::initializer_list<const b200_linear*> ls) {
    const b200_linear* f = *ls.begin();
    for (const b200_linear* l : ls)
        if (l->kind != B200_LIN_MARLIN4 || l->group_size != f->group_size || l->type != f->type || (l->zeros != nullptr) != (f->zeros != nullptr)) return false;
    return true;
}
void marlin_multi(b200_llama* m, int nseg, const b200_linear* const* ls, const void* x16, float* const* ys, const int* ns, int64_t ldy, int B, int k, cudaStream_t st) {
    const void* ws[3]; const void* sc[3]; const void* zp[3];
    for (int i = 0; i < nseg; ++i) { ws[i] = ls[i]->w; sc[i] = ls[i]->scales; zp[i] = ls[i]->zeros; }
    marlin_tc_f32_multi(x16, nseg, ws, sc, ls[0]->type == B200_BF16, ls[0]->zeros ? zp : nullptr, ys, ns, ldy, B, k, ls[0]->group_size, 0, st);
}
void lm_head(b200_llama* m, int B, cudaStream_t st) {
    const int H = m->cfg.hidden;
    const b200_linear& o = m->output;
    const bool ggml_tc = o.kind == B200_LIN_GGML && qmatmul_tc_supported(B, m->vocab_l, H, o.type);
    if ((ggml_tc && qmatmul_tc_needs_zeroed_output(m->vocab_l, H)) || o.kind == B200_LIN_DENSE16) {      // split tiles meet in y through atomics
        launch_pdl(zero_f32_kernel, dim3(sm_count() * 2), dim3(256), 0, st, m->logits, (int64_t)B * m->vocab_l); count_launch();
    }
    linear(m, o, m->xn, m->logits, m->vocab_l, B, m->vocab_l, H, 0, st);
}

// ---- forward on the persistent layer kernel (layer_mega.cu) --------------------------------------------------------------------
// Per layer: {RoPE + cache write, attention, merge, ONE launch for wo -> +x, norm -> gate|up -> SiLU -> w2 -> +x, norm -> QKV of the
// next layer}.  Split-K partial sums travel in slabs and are added in slab order by their consumer: bitwise reproducible.
bool mega_usable(const b200_llama* m, int B) {
    const b200_llama_config& c = m->cfg;
    if (!m->use_mega || !m->qkv_slabs) return false;
    if (c.tp_world > 1 && (int)m->peers.size() != c.tp_world) return false;          // NCCL all-reduce: legacy launches
    if (!mega_supported(B, c.hidden, std::max(c.hidden, m->ffn_l))) return false;
    if ((m->heads_l * c.head_dim) % 256 || m->ffn_l % 256 || c.head_dim % 32) return false;
    if (m->rope_neox) return false;
    for (const auto& w : m->layers)
        for (const b200_linear* l : {&w.wq, &w.wk, &w.wv, &w.wo, &w.w1, &w.w2, &w.w3})
            if (l->kind != B200_LIN_GGML || l->type != B200_GGML_Q4_K) return false;
    return true;
}

int forward_mega(b200_llama* m, int B, cudaStream_t st, bool linear_only) {
    const b200_llama_config& c = m->cfg;
    const int64_t s = reinterpret_cast<int64_t>(st);
    const int H = c.hidden, hd = c.head_dim, L = c.num_layers;
    const int qd = m->heads_l * hd, kd = m->kv_l * hd, F = m->ffn_l;
    const long long n0 = b200_total_kernel_launches();
    const int G = m->mega_G;
    constexpr int kArgmaxChunks = 16;
    const bool fused_ar = c.tp_world > 1;
    const int64_t Bm = c.max_num_seqs;
    auto tiles = [](int n) { return (n + 127) / 128; };

    cudaMemsetAsync(m->mega_counters, 0, m->mega_counter_bytes, st);
    embedding_f32(m->tok_embeddings, m->d_tokens, m->x, B, H, s);
    rms_norm(m->x, m->layers[0].attn_norm, m->xn, B, H, c.rms_eps, B200_F16_K4, s);

    auto base_params = [&](MegaParams& P, int launch) {
        memset(&P, 0, sizeof(P));
        P.m = B; P.hidden = H; P.eps = c.rms_eps; P.x = m->x;
        P.counters = m->mega_counters + (size_t)launch * 2 * kMegaMaxPhases;
        P.error_word = m->d_timeout + 1;
        P.tp_rank = c.tp_rank; P.tp_world = c.tp_world; P.rows_max = c.max_num_seqs;
        P.tp_timeout_ns = (unsigned long long)m->tp_timeout_ms * 1000000ull;
        P.timeout_word = m->d_timeout;
        if (fused_ar) for (int i = 0; i < c.tp_world; ++i) P.peers.p[i] = static_cast<char*>(m->peers[i]);
        P.trace = launch == m->mega_trace_launch ? m->mega_trace : nullptr;
        static const int trig = [] { const char* e = getenv("B200_MEGA_TRIGGER"); return e ? atoi(e) : 0; }();     // early trigger measured 9 % slower
        P.early_trigger = m->mega_mode == 2 ? 1 : trig;      // single-phase launches behave like the per-GEMM kernels: let the next grid start early
    };
    auto qkv_phase = [&](MegaParams& P, MegaPhase& ph, const b200_llama_layer_ex& w, int map0) -> bool {
        if (!mega_make_w_map(&P.maps[map0], w.wq.w, qd, H) || !mega_make_w_map(&P.maps[map0 + 1], w.wk.w, kd, H) ||
            !mega_make_w_map(&P.maps[map0 + 2], w.wv.w, kd, H)) return false;
        ph.w_map[0] = map0; ph.w_map[1] = map0 + 1; ph.w_map[2] = map0 + 2;
        ph.n[0] = qd; ph.n[1] = kd; ph.n[2] = kd;
        ph.tile_end[0] = tiles(qd); ph.tile_end[1] = tiles(qd) + tiles(kd); ph.tile_end[2] = tiles(qd) + 2 * tiles(kd);
        ph.y[0] = m->qkv_slabs; ph.y[1] = m->qkv_slabs + qd; ph.y[2] = m->qkv_slabs + qd + kd;
        ph.ldy = m->qkv_row; ph.slab_stride = Bm * m->qkv_row;
        ph.nsb = H / 256; ph.n_tiles = ph.tile_end[2];
        return true;
    };
    SlabInfo qkv_si{};
    qkv_si.slab_stride = Bm * m->qkv_row; qkv_si.nsb = H / 256; qkv_si.n_tiles = tiles(qd) + 2 * tiles(kd);
    qkv_si.grid = std::min(G, qkv_si.n_tiles * qkv_si.nsb);                       // effective grid of the QKV phase
    qkv_si.seg_tile0[0] = 0; qkv_si.seg_tile0[1] = tiles(qd); qkv_si.seg_tile0[2] = tiles(qd) + tiles(kd);

    {   // QKV of layer 0: a one-phase launch (its activations come from the rms_norm above)
        MegaParams P;
        base_params(P, 0);
        if (!mega_make_x_map(&P.maps[0], m->xn, B, H)) return 0;
        MegaPhase& p0 = P.phase[0];
        p0.x_map = 0; p0.eop = kEopNone;
        if (!qkv_phase(P, p0, m->layers[0], 1)) return 0;
        P.n_phases = 1;
        mega_launch(P, st);
    }
    for (int l = 0; l < L; ++l) {
        const b200_llama_layer_ex& w = m->layers[l];
        if (!linear_only) {
            rope_and_cache_slabs(m->qkv_slabs, qkv_si, m->q16, m->kc[l], m->vc[l], m->cos_t, m->sin_t, m->d_positions, m->d_slots, B,
                                 m->heads_l, m->kv_l, hd, /*interleaved=*/1, B200_BF16, c.kv_dtype, s);
            paged_attention_decode(m->attn16, m->q16, m->kc[l], m->vc[l], m->d_tables, m->d_ctx, B, m->heads_l, m->kv_l, hd,
                                   c.block_size, c.max_blocks_per_seq, m->num_blocks, 1.0f / sqrtf((float)hd), 0.f, 0,
                                   B200_BF16, c.kv_dtype, B200_KV_FLASH, B200_F16_K4, m->attn_ws, m->attn_ws_bytes, s);
        }
        MegaParams P;
        base_params(P, l + 1);
        // maps: 0 attn16 [B, qd], 1 xn [B, H], 2 act16 [B, F], 3 wo, 4 w1, 5 w3, 6 w2, 7..9 q, k, v of the next layer
        if (!mega_make_x_map(&P.maps[0], m->attn16, B, qd) || !mega_make_x_map(&P.maps[1], m->xn, B, H) ||
            !mega_make_x_map(&P.maps[2], m->act16, B, F) || !mega_make_w_map(&P.maps[3], w.wo.w, H, qd) ||
            !mega_make_w_map(&P.maps[4], w.w1.w, F, H) || !mega_make_w_map(&P.maps[5], w.w3.w, F, H) || !mega_make_w_map(&P.maps[6], w.w2.w, H, F)) return 0;
        const int norm_op = fused_ar ? kEopTpNorm : kEopNorm;
        MegaPhase& a = P.phase[0];        // x (+)= wo(attn): partial sums -> ro_slabs
        a.x_map = 0; a.w_map[0] = a.w_map[1] = a.w_map[2] = 3; a.n[0] = a.n[1] = a.n[2] = H;
        a.tile_end[0] = tiles(H); a.tile_end[1] = a.tile_end[2] = 0x7fffffff;
        a.y[0] = a.y[1] = a.y[2] = m->ro_slabs; a.ldy = H; a.slab_stride = Bm * H; a.nsb = qd / 256; a.n_tiles = tiles(H); a.eop = kEopNone;
        MegaPhase& b = P.phase[1];        // gate | up on xn = norm(x + wo partials)
        b.x_map = 1; b.w_map[0] = 4; b.w_map[1] = b.w_map[2] = 5; b.n[0] = b.n[1] = b.n[2] = F;
        b.tile_end[0] = tiles(F); b.tile_end[1] = 2 * tiles(F); b.tile_end[2] = 0x7fffffff;
        b.y[0] = m->gate_slabs; b.y[1] = b.y[2] = m->up_slabs; b.ldy = F; b.slab_stride = Bm * F; b.nsb = H / 256; b.n_tiles = 2 * tiles(F);
        b.eop = norm_op; b.norm_w = w.ffn_norm; b.act_out = m->xn;
        MegaPhase& d = P.phase[2];        // w2 on act = silu(gate) * up
        d.x_map = 2; d.w_map[0] = d.w_map[1] = d.w_map[2] = 6; d.n[0] = d.n[1] = d.n[2] = H;
        d.tile_end[0] = tiles(H); d.tile_end[1] = d.tile_end[2] = 0x7fffffff;
        d.y[0] = d.y[1] = d.y[2] = m->ro_slabs; d.ldy = H; d.slab_stride = Bm * H; d.nsb = F / 256; d.n_tiles = tiles(H);
        d.eop = kEopSilu; d.act_out = m->act16;
        MegaPhase& e = P.phase[3];        // x += w2 partials, next norm; QKV of the next layer (last layer: the final norm only)
        e.x_map = 1; e.eop = norm_op; e.act_out = m->xn;
        if (l + 1 < L) {
            e.norm_w = m->layers[l + 1].attn_norm;
            if (!qkv_phase(P, e, m->layers[l + 1], 7)) return 0;
        } else {
            e.norm_w = m->norm;
            e.n_tiles = 0; e.nsb = 0; e.tile_end[0] = e.tile_end[1] = e.tile_end[2] = 0x7fffffff;
        }
        P.n_phases = 4;
        if (m->mega_mode == 2) {
            // "split" mode: the same phases and the same deterministic slab reduction, but one launch per GEMM phase and the elementwise op
            // as its own small kernel in between, all chained by programmatic dependent launch
            for (int i = 0; i < 4; ++i) {
                if (P.phase[i].eop != kEopNone) {
                    MegaParams E = P;
                    E.phase[0] = P.phase[i - 1]; E.phase[1] = P.phase[i]; E.n_phases = 2;
                    mega_eop_launch(E, P.phase[i].eop, st);
                }
                if (P.phase[i].n_tiles > 0) {
                    MegaParams S = P;
                    S.phase[0] = P.phase[i]; S.phase[0].eop = kEopNone; S.n_phases = 1; S.early_trigger = 1;
                    mega_launch(S, st);
                }
            }
        } else {
            mega_launch(P, st);
        }
    }
    lm_head(m, B, st);
    const int live_cols = std::max(0, std::min(m->vocab_l, c.vocab - c.tp_rank * m->vocab_l));
    argmax_pairs(m->logits, m->tp_pairs, B, m->vocab_l, live_cols, kArgmaxChunks, c.tp_rank * m->vocab_l, st);
    if (c.tp_world == 1) {
        argmax_reduce_pairs(m->tp_pairs, m->next_tokens, B, kArgmaxChunks, st);
    } else {
        tp_allgather_bytes(m->comm, m->tp_pairs, m->tp_gathered, (size_t)B * kArgmaxChunks * 8, st);
        argmax_reduce_pairs(m->tp_gathered, m->next_tokens, B, c.tp_world * kArgmaxChunks, st);
    }
    return (int)(b200_total_kernel_launches() - n0);
}

// linear_only: skip RoPE + cache write + attention (the measurement leg behind bench.py's roofline_gemm: the weight stream
// of all quantised projections and the small ops between them, without the KV stream)
int forward(b200_llama* m, int B, cudaStream_t st, bool linear_only = false) {
    if (mega_usable(m, B)) return forward_mega(m, B, st, linear_only);
    const b200_llama_config& c = m->cfg;
    const int64_t s = reinterpret_cast<int64_t>(st);
    const int H = c.hidden, hd = c.head_dim;
    const int qd = m->heads_l * hd, kd = m->kv_l * hd;
    const long long n0 = b200_total_kernel_launches();
    constexpr int kArgmaxChunks = 16;
    embedding_f32(m->tok_embeddings, m->d_tokens, m->x, B, H, s);
    // tensor parallel with peer inboxes: the all-reduce of the row-parallel GEMMs, the residual add and the NEXT RMSNorm are
    // one kernel over NVLink peer memory (tp.cu); `partial` is left zeroed by that kernel for the next split-K GEMM
    const bool fused_ar = c.tp_world > 1 && (int)m->peers.size() == c.tp_world && act_format(m) == B200_F16_K4;
    auto residual_fused = [&](const float* next_norm) {
        tp_allreduce_add_norm(m->partial, m->x, next_norm, m->xn, m->peers.data(), c.tp_rank, c.tp_world, B, H, c.max_num_seqs, c.rms_eps, m->d_timeout, st);
    };
    const int fmt = act_format(m);                      // what the linears read: fp16 K4 (GGML, int4) or the dense weights' dtype
    const bool all_ggml = [&] {
        for (const auto& w : m->layers)
            for (const b200_linear* q : {&w.wq, &w.wk, &w.wv, &w.w1, &w.w3}) if (q->kind != B200_LIN_GGML) return false;
        return true;
    }();
    for (int l = 0; l < c.num_layers; ++l) {
        const b200_llama_layer_ex& w = m->layers[l];
        if (!fused_ar || l == 0) rms_norm(m->x, w.attn_norm, m->xn, B, H, c.rms_eps, fmt, s);
        // QKV / gate / up accumulate (split-K) into buffers that their consumers leave zeroed.  Every GEMM of the chain tells the L2 what
        // the NEXT one will stream (prefetch_next): fetched from the tail of the launch, while HBM is otherwise idle
        prefetch_next({{&w.wo, {H, qd}}});
        if (all_ggml) {   // fused QKV: three weight matrices, one launch
            const void* ws[3] = {w.wq.w, w.wk.w, w.wv.w};
            const int ts[3] = {w.wq.type, w.wk.type, w.wv.type}, ns[3] = {qd, kd, kd};
            float* ys[3] = {m->qkv, m->qkv + qd, m->qkv + qd + kd};
            // accumulate = 0: whole tiles are plain stores, split tiles red.add into the zeroed buffer
            qmatmul_dispatch_multi(m->xn, 3, ws, ts, ys, ns, m->qkv_row, B, H, 0, st);
        } else if (marlin_fusable({&w.wq, &w.wk, &w.wv})) {
            const b200_linear* ls[3] = {&w.wq, &w.wk, &w.wv};
            const int ns[3] = {qd, kd, kd};
            float* ys[3] = {m->qkv, m->qkv + qd, m->qkv + qd + kd};
            marlin_multi(m, 3, ls, m->xn, ys, ns, m->qkv_row, B, H, st);
        } else {
            linear(m, w.wq, m->xn, m->qkv, m->qkv_row, B, qd, H, 0, st);
            linear(m, w.wk, m->xn, m->qkv + qd, m->qkv_row, B, kd, H, 0, st);
            linear(m, w.wv, m->xn, m->qkv + qd + kd, m->qkv_row, B, kd, H, 0, st);
        }
        if (linear_only) { launch_pdl(zero_f32_kernel, dim3(64), dim3(256), 0, st, m->qkv, (int64_t)B * m->qkv_row); count_launch(); } else {
        // (also re-zeroes qkv: the split-K GEMMs accumulate into it)
        rope_and_cache_impl(m->qkv, m->q16, m->kc[l], m->vc[l], m->cos_t, m->sin_t, m->d_positions, m->d_slots, B,
                            m->heads_l, m->kv_l, hd, /*interleaved=*/m->rope_neox ? 0 : 1, B200_BF16, c.kv_dtype, /*zero_src=*/true, s);
        paged_attention_decode(m->attn16, m->q16, m->kc[l], m->vc[l], m->d_tables, m->d_ctx, B, m->heads_l, m->kv_l, hd,
                               c.block_size, c.max_blocks_per_seq, m->num_blocks, 1.0f / sqrtf((float)hd), 0.f, 0,
                               B200_BF16, c.kv_dtype, B200_KV_FLASH, fmt, m->attn_ws, m->attn_ws_bytes, s);
        }
        prefetch_next({{&w.w1, {m->ffn_l, H}}, {&w.w3, {m->ffn_l, H}}});
        if (c.tp_world == 1) {
            linear(m, w.wo, m->attn16, m->x, H, B, H, qd, 1, st);                        // x += wo(attn)
        } else if (fused_ar) {
            linear(m, w.wo, m->attn16, m->partial, H, B, H, qd, 1, st);
            residual_fused(w.ffn_norm);                                                    // x += sum_ranks(partial); xn = ffn_norm(x)
        } else {
            // row-parallel: partial sums -> all-reduce -> residual add (distributed.rs:696-710)
            launch_pdl(zero_f32_kernel, dim3(64), dim3(256), 0, st, m->partial, (int64_t)B * H); count_launch();
            linear(m, w.wo, m->attn16, m->partial, H, B, H, qd, 1, st);
            tp_allreduce_f32(m->comm, m->partial, (int64_t)B * H, st);   // tp.cu (NCCL)
            add_f32(m->x, m->partial, (int64_t)B * H, s);
        }
        if (!fused_ar) rms_norm(m->x, w.ffn_norm, m->xn, B, H, c.rms_eps, fmt, s);
        prefetch_next({{&w.w2, {H, m->ffn_l}}});
        if (all_ggml) {   // fused gate | up
            const void* ws[2] = {w.w1.w, w.w3.w};
            const int ts[2] = {w.w1.type, w.w3.type}, ns[2] = {m->ffn_l, m->ffn_l};
            float* ys[2] = {m->gate, m->up};
            qmatmul_dispatch_multi(m->xn, 2, ws, ts, ys, ns, m->ffn_l, B, H, 0, st);
        } else if (marlin_fusable({&w.w1, &w.w3})) {
            const b200_linear* ls[2] = {&w.w1, &w.w3};
            const int ns[2] = {m->ffn_l, m->ffn_l};
            float* ys[2] = {m->gate, m->up};
            marlin_multi(m, 2, ls, m->xn, ys, ns, m->ffn_l, B, H, st);
        } else {
            linear(m, w.w1, m->xn, m->gate, m->ffn_l, B, m->ffn_l, H, 0, st);
            linear(m, w.w3, m->xn, m->up, m->ffn_l, B, m->ffn_l, H, 0, st);
        }
        silu_mul_zero_src_fmt(m->gate, m->up, m->act16, (int64_t)B * m->ffn_l, fmt, s);      // act = silu(gate)*up; gate/up re-zeroed
        if (l + 1 < c.num_layers) {
            const b200_llama_layer_ex& nx = m->layers[l + 1];
            prefetch_next({{&nx.wq, {qd, H}}, {&nx.wk, {kd, H}}, {&nx.wv, {kd, H}}});
        }
        if (c.tp_world == 1) {
            linear(m, w.w2, m->act16, m->x, H, B, H, m->ffn_l, 1, st);                   // x += w2(act)
        } else if (fused_ar) {
            linear(m, w.w2, m->act16, m->partial, H, B, H, m->ffn_l, 1, st);
            residual_fused(l + 1 < c.num_layers ? m->layers[l + 1].attn_norm : m->norm);
        } else {
            launch_pdl(zero_f32_kernel, dim3(64), dim3(256), 0, st, m->partial, (int64_t)B * H); count_launch();
            linear(m, w.w2, m->act16, m->partial, H, B, H, m->ffn_l, 1, st);
            tp_allreduce_f32(m->comm, m->partial, (int64_t)B * H, st);
            add_f32(m->x, m->partial, (int64_t)B * H, s);
        }
    }
    if (!fused_ar) rms_norm(m->x, m->norm, m->xn, B, H, c.rms_eps, fmt, s);
    lm_head(m, B, st);
    // greedy sampling in two stages; vocab-parallel lm_head (distributed.rs:1632-1667): gather (max, global index) pairs
    // instead of the logits
    // vocab-parallel: this rank's columns are [tp_rank * vocab_l, ...); columns at or beyond `vocab` are padding (never sampled)
    const int live_cols = std::max(0, std::min(m->vocab_l, c.vocab - c.tp_rank * m->vocab_l));
    argmax_pairs(m->logits, m->tp_pairs, B, m->vocab_l, live_cols, kArgmaxChunks, c.tp_rank * m->vocab_l, st);
    if (c.tp_world == 1) {
        argmax_reduce_pairs(m->tp_pairs, m->next_tokens, B, kArgmaxChunks, st);
    } else {
        tp_allgather_bytes(m->comm, m->tp_pairs, m->tp_gathered, (size_t)B * kArgmaxChunks * 8, st);
        argmax_reduce_pairs(m->tp_gathered, m->next_tokens, B, c.tp_world * kArgmaxChunks, st);
    }
    return (int)(b200_total_kernel_launches() - n0);
}

// captured graphs bake in weight / cache / communicator pointers: every setter that changes one drops them
void invalidate_graphs(b200_llama* m) {
    for (auto& g : m->graphs) cudaGraphExecDestroy(g.second);
    m->graphs.clear();
    m->launches_per_step.clear();
}

// a row of the fused all-reduce gave up waiting for a peer: the step's outputs are NaN -- report it instead of returning them
bool peer_timed_out(b200_llama* m, const char* who) {
    if (m->h_timeout && reinterpret_cast<volatile uint32_t*>(m->h_timeout)[1] != 0u) {
        set_error(kErrCuda, "%s: a grid-wide wait of the persistent layer kernel gave up (CTAs not co-resident?)", who);
        return true;
    }
    if (!m->h_timeout || *reinterpret_cast<volatile uint32_t*>(m->h_timeout) == 0u) return false;
    set_error(kErrCuda, "%s: the fused tensor-parallel all-reduce timed out waiting for a peer rank (B200_TP_TIMEOUT_MS); this step's outputs are invalid", who);
    return true;
}

bool ready(b200_llama* m) {
    if (!m || !m->ok) { set_error(kErrBadArg, "engine: model not initialised"); return false; }
    if (!m->tok_embeddings || !m->norm || !m->output.w) { set_error(kErrBadArg, "engine: globals not set"); return false; }
    if ((int)m->kc.size() != m->cfg.num_layers) { set_error(kErrBadArg, "engine: kv cache not set"); return false; }
    for (auto& l : m->layers) if (!l.wq.w) { set_error(kErrBadArg, "engine: layer weights not set"); return false; }
    if (m->cfg.tp_world > 1 && !m->comm) { set_error(kErrBadArg, "engine: tp_world > 1 but no communicator"); return false; }
    return true;
}

// launch the step: graph replay when enabled (capture on first use), eager otherwise
void run_step(b200_llama* m, int B, cudaStream_t st) {
    if (!m->cfg.use_graph || st == nullptr) {
        m->launches += forward(m, B, st);
        return;
    }
    auto it = m->graphs.find(B);
    if (it == m->graphs.end()) {
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            set_error(kErrCuda, "engine: begin capture failed: %s", cudaGetErrorString(cudaGetLastError()));
            return;
        }
        const int n = forward(m, B, st);
        cudaError_t e = cudaStreamEndCapture(st, &g);
        if (e != cudaSuccess || !g) { set_error(kErrCuda, "engine: end capture failed: %s", cudaGetErrorString(e)); return; }
        cudaGraphExec_t ex = nullptr;
        e = cudaGraphInstantiate(&ex, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { set_error(kErrCuda, "engine: graph instantiate failed: %s", cudaGetErrorString(e)); return; }
        m->graphs[B] = ex;
        m->launches_per_step[B] = n;
        it = m->graphs.find(B);
    }
    cudaError_t e = cudaGraphLaunch(it->second, st);
    if (e != cudaSuccess) { set_error(kErrCuda, "engine: graph launch failed: %s", cudaGetErrorString(e)); return; }
    m->launches += m->launches_per_step[B];
}

}  // namespace

extern "C" {

b200_llama* b200_llama_create(const b200_llama_config* cfg) {
    if (!cfg) { set_error(kErrBadArg, "b200_llama_create: null config"); return nullptr; }
    const b200_llama_config& c = *cfg;
    if (c.hidden <= 0 || c.num_layers <= 0 || c.num_heads <= 0 || c.num_kv_heads <= 0 || c.head_dim <= 0 || c.ffn <= 0 ||
        c.vocab <= 0 || c.block_size <= 0 || c.max_num_seqs <= 0 || c.max_blocks_per_seq <= 0 || c.max_pos <= 0 ||
        c.tp_world <= 0 || c.tp_rank < 0 || c.tp_rank >= c.tp_world) {
        set_error(kErrBadArg, "b200_llama_create: bad config");
        return nullptr;
    }
    if (c.num_heads % c.tp_world || c.ffn % c.tp_world) {
        set_error(kErrBadArg, "b200_llama_create: heads/ffn not divisible by tp_world=%d", c.tp_world);
        return nullptr;
    }
    if (c.max_blocks_per_seq > (1 << 20) || (int64_t)c.max_blocks_per_seq * c.block_size > INT32_MAX) { set_error(kErrBadArg, "b200_llama_create: table too wide"); return nullptr; }
    if (b200_device_cc() < 100) { set_error(kErrNoDevice, "b200_llama_create: no sm_100 device (cc=%d)", b200_device_cc()); return nullptr; }
    b200_llama* m = new b200_llama();
    m->cfg = c;
    m->layers.resize(c.num_layers);
    for (auto& l : m->layers) l = b200_llama_layer_ex{};
    m->heads_l = c.num_heads / c.tp_world;
    // kv_head_shard (/root/reference/src/openai/distributed.rs:725-765): split, or replicate when kvh < world
    m->kv_l = c.num_kv_heads >= c.tp_world ? c.num_kv_heads / c.tp_world : 1;
    m->ffn_l = c.ffn / c.tp_world;
    // pad_vocab_size (/root/reference/src/openai/distributed.rs:1448-1454): multiple of 64, of the world size, of 64 again; every
    // rank owns vocab_pad / world rows of the lm_head (rows past `vocab` are padding the caller fills with zeros)
    {
        const int64_t padded = ((int64_t)c.vocab + 63) / 64 * 64;
        const int64_t per_rank = (padded + c.tp_world - 1) / c.tp_world * c.tp_world;
        m->vocab_pad = c.tp_world == 1 ? c.vocab : (int)((per_rank + 63) / 64 * 64);
        if (m->vocab_pad % c.tp_world) { set_error(kErrBadArg, "b200_llama_create: padded vocab %d not divisible by tp_world %d", m->vocab_pad, c.tp_world); delete m; return nullptr; }
    }
    m->vocab_l = m->vocab_pad / c.tp_world;
    m->qkv_row = (m->heads_l + 2 * m->kv_l) * c.head_dim;
    const size_t B = c.max_num_seqs;
    constexpr size_t kChunks = 16;
    // step metadata lives in ONE device slab with the layout of the pinned staging slab: one H2D copy per step
    const size_t Bp = (B + 3) & ~(size_t)3;
    m->stage_bytes = Bp * (8 + 8 + 8 + 4) + B * c.max_blocks_per_seq * 4 + 64;
    bool ok = dmalloc(m->d_meta, m->stage_bytes);
    if (ok) {
        m->d_tokens = reinterpret_cast<int64_t*>(m->d_meta); m->d_positions = m->d_tokens + Bp; m->d_slots = m->d_positions + Bp;
        m->d_ctx = reinterpret_cast<uint32_t*>(m->d_slots + Bp); m->d_tables = m->d_ctx + Bp;
    }
    ok = ok && dmalloc(m->x, B * c.hidden) && dmalloc(m->xn, B * c.hidden) &&
              dmalloc(m->qkv, B * m->qkv_row) && dmalloc(m->q16, B * m->heads_l * c.head_dim) &&
              dmalloc(m->attn16, B * m->heads_l * c.head_dim) && dmalloc(m->gate, B * m->ffn_l) && dmalloc(m->up, B * m->ffn_l) &&
              dmalloc(m->act16, B * m->ffn_l) && dmalloc(m->partial, B * c.hidden) && dmalloc(m->logits, B * m->vocab_l) &&
              dmalloc(m->next_tokens, B) && dmalloc(m->tp_pairs, B * 2 * kChunks) && dmalloc(m->tp_gathered, B * 2 * kChunks * c.tp_world);
    m->attn_ws_bytes = paged_attention_decode_workspace_bytes((int)B, m->heads_l, c.head_dim, c.max_blocks_per_seq, c.block_size);
    char* ws = nullptr;
    ok = ok && dmalloc(ws, m->attn_ws_bytes);
    m->attn_ws = ws;
    // RoPE tables: calculate_default_inv_freq (/root/reference/src/openai/models/layers/rotary_emb.rs:14-19,31-36)
    const int half = c.head_dim / 2;
    std::vector<float> hc((size_t)c.max_pos * half), hs((size_t)c.max_pos * half);
    for (int i = 0; i < half; ++i) {
        const float inv = 1.0f / (float)std::pow((double)c.rope_theta, (double)(2 * i) / (double)c.head_dim);
        for (int p = 0; p < c.max_pos; ++p) {
            const float ang = (float)p * inv;
            hc[(size_t)p * half + i] = cosf(ang);
            hs[(size_t)p * half + i] = sinf(ang);
        }
    }
    ok = ok && dmalloc(m->cos_t, hc.size()) && dmalloc(m->sin_t, hs.size());
    if (ok) {
        cudaMemcpy(m->cos_t, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(m->sin_t, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice);
    }
    ok = ok && cudaMallocHost((void**)&m->h_stage[0], m->stage_bytes) == cudaSuccess &&
         cudaMallocHost((void**)&m->h_stage[1], m->stage_bytes) == cudaSuccess &&
         cudaEventCreateWithFlags(&m->stage_ev[0], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&m->stage_ev[1], cudaEventDisableTiming) == cudaSuccess &&
         cudaMallocHost((void**)&m->h_next, B * 4) == cudaSuccess &&
         cudaHostAlloc((void**)&m->h_timeout, 64, cudaHostAllocMapped) == cudaSuccess &&
         cudaHostGetDevicePointer((void**)&m->d_timeout, m->h_timeout, 0) == cudaSuccess;
    if (ok) { m->h_timeout[0] = 0u; m->h_timeout[1] = 0u; }
    static const long long tmo = [] { const char* e = getenv("B200_TP_TIMEOUT_MS"); return e ? atoll(e) : 120000ll; }();
    m->tp_timeout_ms = tmo > 0 ? tmo : 1;
    if (ok && c.tp_world > 1) {
        ok = dmalloc(m->logits_gathered, B * (size_t)m->vocab_pad) && dmalloc(m->logits_full, B * (size_t)c.vocab);
        tp_set_timeout_ms(m->tp_timeout_ms);
    }
    // persistent layer kernel (csrc/layer_mega.cu), opt-in with B200_MEGA=1: bitwise-deterministic split-K reduction, but its in-kernel
    // grid-wide syncs cost more than the PDL-overlapped kernel boundaries of the one-launch-per-GEMM path (measured: 5.5 vs 5.3 ms /
    // step at one GPU, 4.9 vs 4.1 ms at TP = 2; profiles/r02_layer_kernel.md), so the default stays one launch per GEMM
    const int mega_on = [] { const char* e = getenv("B200_MEGA"); return e ? atoi(e) : 0; }();      // read per model: tests toggle it
    if (ok && mega_on && c.hidden % 256 == 0 && c.hidden <= 8192 && (m->heads_l * c.head_dim) % 256 == 0 && m->ffn_l % 256 == 0) {
        auto tiles = [](int n) { return (n + 127) / 128; };
        const int G = mega_grid();
        const int qd = m->heads_l * c.head_dim, kd = m->kv_l * c.head_dim;
        m->mega_G = G;
        m->s_qkv = mega_phase_slabs(tiles(qd) + 2 * tiles(kd), c.hidden / 256, G);
        m->s_ro = std::max(mega_phase_slabs(tiles(c.hidden), qd / 256, G), mega_phase_slabs(tiles(c.hidden), m->ffn_l / 256, G));
        m->s_gu = mega_phase_slabs(2 * tiles(m->ffn_l), c.hidden / 256, G);
        m->mega_counter_bytes = (size_t)(c.num_layers + 1) * 2 * kMegaMaxPhases * sizeof(uint32_t);
        ok = dmalloc(m->qkv_slabs, (size_t)m->s_qkv * B * m->qkv_row) && dmalloc(m->ro_slabs, (size_t)m->s_ro * B * c.hidden) &&
             dmalloc(m->gate_slabs, (size_t)m->s_gu * B * m->ffn_l) && dmalloc(m->up_slabs, (size_t)m->s_gu * B * m->ffn_l) &&
             dmalloc(m->mega_counters, m->mega_counter_bytes / sizeof(uint32_t));
        m->use_mega = ok && m->s_qkv <= kMegaMaxSlabs && m->s_ro <= kMegaMaxSlabs && m->s_gu <= 4;
        m->mega_mode = mega_on;
        if (const char* tr = getenv("B200_MEGA_TRACE")) {
            m->mega_trace_launch = atoi(tr);
            ok = ok && dmalloc(m->mega_trace, (size_t)G * kMegaMaxPhases * 8);
        }
    }
    if (!ok) {
        if (!b200_last_error()) set_error(kErrCuda, "b200_llama_create: allocation failed");
        b200_llama_destroy(m);
        return nullptr;
    }
    m->ok = true;
    return m;
}

void b200_llama_destroy(b200_llama* m) {
    if (!m) return;
    for (auto& g : m->graphs) cudaGraphExecDestroy(g.second);
    void* ptrs[] = {m->d_meta, m->x, m->xn, m->qkv, m->q16, m->attn16,
                    m->gate, m->up, m->act16, m->partial, m->logits, m->next_tokens, m->cos_t, m->sin_t, m->attn_ws, m->tp_pairs, m->tp_gathered};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < 2; ++i) { if (m->h_stage[i]) cudaFreeHost(m->h_stage[i]); if (m->stage_ev[i]) cudaEventDestroy(m->stage_ev[i]); }
    if (m->h_next) cudaFreeHost(m->h_next);
    if (m->h_timeout) cudaFreeHost(m->h_timeout);
    if (m->logits_gathered) cudaFree(m->logits_gathered);
    if (m->logits_full) cudaFree(m->logits_full);
    if (m->mega_trace) cudaFree(m->mega_trace);
    for (void* p : {(void*)m->qkv_slabs, (void*)m->ro_slabs, (void*)m->gate_slabs, (void*)m->up_slabs, (void*)m->mega_counters}) if (p) cudaFree(p);
    delete m;
}

static b200_linear ggml_linear(const void* w, int type) { b200_linear l{}; l.kind = B200_LIN_GGML; l.type = type; l.w = w; return l; }

static bool check_linear(const b200_linear& l, int n, int k, const char* name) {
    if (!l.w) { set_error(kErrBadArg, "engine: %s: null weight", name); return false; }
    switch (l.kind) {
        case B200_LIN_GGML:
            if (l.type != B200_GGML_Q4_K && l.type != B200_GGML_Q6_K && l.type != B200_GGML_Q8_0) { set_error(kErrUnsupported, "engine: %s: ggml type %d", name, l.type); return false; }
            return true;
        case B200_LIN_MARLIN4:
            if (!l.scales || (l.type != B200_F16 && l.type != B200_BF16) || (l.group_size != -1 && l.group_size != 64 && l.group_size != 128) || k % 256 || n % 64) {
                set_error(kErrUnsupported, "engine: %s: marlin int4 needs f16 / bf16 scales, group 64 / 128 / -1, k %% 256 == 0, n %% 64 == 0 (n=%d k=%d g=%d)", name, n, k, l.group_size);
                return false;
            }
            return true;
        case B200_LIN_DENSE16:
            if ((l.type != B200_F16 && l.type != B200_BF16) || k % 8 || ((uintptr_t)l.w & 15)) { set_error(kErrUnsupported, "engine: %s: dense weights must be f16 / bf16, k %% 8 == 0, 16-byte aligned", name); return false; }
            return true;
        default: set_error(kErrUnsupported, "engine: %s: linear kind %d", name, l.kind); return false;
    }
}
static int fmt_of(const b200_linear& l) { return l.kind == B200_LIN_DENSE16 ? l.type : B200_F16_K4; }

void b200_llama_set_layer_ex(b200_llama* m, int32_t layer, const b200_llama_layer_ex* w) {
    B200_REQUIRE(m && w && layer >= 0 && layer < m->cfg.num_layers, kErrBadArg, "b200_llama_set_layer_ex: bad arguments");
    B200_REQUIRE(w->attn_norm && w->ffn_norm, kErrBadArg, "b200_llama_set_layer_ex: null norm in layer %d", layer);
    const b200_llama_config& c = m->cfg;
    const int H = c.hidden, qd = m->heads_l * c.head_dim, kd = m->kv_l * c.head_dim, F = m->ffn_l;
    if (!check_linear(w->wq, qd, H, "wq") || !check_linear(w->wk, kd, H, "wk") || !check_linear(w->wv, kd, H, "wv") ||
        !check_linear(w->wo, H, qd, "wo") || !check_linear(w->w1, F, H, "w1") || !check_linear(w->w2, H, F, "w2") || !check_linear(w->w3, F, H, "w3")) return;
    const int fmt = fmt_of(w->wq);
    for (const b200_linear* l : {&w->wk, &w->wv, &w->wo, &w->w1, &w->w2, &w->w3})
        B200_REQUIRE(fmt_of(*l) == fmt, kErrUnsupported, "b200_llama_set_layer_ex: the linears of a model must share one activation format");
    m->layers[layer] = *w;
    invalidate_graphs(m);
}

void b200_llama_set_layer(b200_llama* m, int32_t layer, const b200_llama_layer* w) {
    B200_REQUIRE(m && w && layer >= 0 && layer < m->cfg.num_layers, kErrBadArg, "b200_llama_set_layer: bad arguments");
    B200_REQUIRE(w->attn_norm && w->ffn_norm && w->wq && w->wk && w->wv && w->wo && w->w1 && w->w2 && w->w3, kErrBadArg,
                 "b200_llama_set_layer: null weight in layer %d", layer);
    b200_llama_layer_ex e{};
    e.attn_norm = w->attn_norm; e.ffn_norm = w->ffn_norm;
    e.wq = ggml_linear(w->wq, w->tq); e.wk = ggml_linear(w->wk, w->tk); e.wv = ggml_linear(w->wv, w->tv); e.wo = ggml_linear(w->wo, w->to);
    e.w1 = ggml_linear(w->w1, w->t1); e.w2 = ggml_linear(w->w2, w->t2); e.w3 = ggml_linear(w->w3, w->t3);
    b200_llama_set_layer_ex(m, layer, &e);
}

void b200_llama_set_globals_ex(b200_llama* m, const float* tok_embeddings, const float* norm, const b200_linear* output, int32_t rope_neox) {
    B200_REQUIRE(m && tok_embeddings && norm && output, kErrBadArg, "b200_llama_set_globals_ex: null pointer");
    if (!check_linear(*output, m->vocab_l, m->cfg.hidden, "output")) return;
    B200_REQUIRE(output->kind != B200_LIN_MARLIN4, kErrUnsupported, "b200_llama_set_globals_ex: the lm_head is GGML or dense (the reference never quantises it to int4)");
    m->tok_embeddings = tok_embeddings; m->norm = norm; m->output = *output; m->rope_neox = rope_neox ? 1 : 0;
    invalidate_graphs(m);
}

void b200_llama_set_globals(b200_llama* m, const float* tok_embeddings, const float* norm, const void* output_w, int32_t output_type) {
    B200_REQUIRE(m && tok_embeddings && norm && output_w, kErrBadArg, "b200_llama_set_globals: null pointer");
    const b200_linear o = ggml_linear(output_w, output_type);
    b200_llama_set_globals_ex(m, tok_embeddings, norm, &o, 0);
}

void b200_llama_set_kv_cache(b200_llama* m, void* const* key_caches, void* const* value_caches, int64_t num_blocks) {
    B200_REQUIRE(m && key_caches && value_caches && num_blocks > 0, kErrBadArg, "b200_llama_set_kv_cache: bad arguments");
    m->kc.assign(key_caches, key_caches + m->cfg.num_layers);
    m->vc.assign(value_caches, value_caches + m->cfg.num_layers);
    m->num_blocks = num_blocks;
    invalidate_graphs(m);                                        // pointers are baked into captured graphs
}

size_t b200_llama_peer_inbox_bytes(const b200_llama* m) {
    return m ? tp_peer_inbox_bytes(m->cfg.tp_world, m->cfg.max_num_seqs, m->cfg.hidden) : 0;
}

void b200_llama_set_peer_inboxes(b200_llama* m, void* const* inboxes, int32_t count) {
    B200_REQUIRE(m, kErrBadArg, "b200_llama_set_peer_inboxes: null model");
    if (!inboxes || count == 0) { m->peers.clear(); }
    else {
        B200_REQUIRE(count == m->cfg.tp_world && count <= 8, kErrBadArg, "b200_llama_set_peer_inboxes: %d inboxes for tp_world %d (max 8)", count, m->cfg.tp_world);
        B200_REQUIRE(m->cfg.hidden % 4 == 0 && m->cfg.hidden <= 8192, kErrUnsupported, "b200_llama_set_peer_inboxes: hidden %d", m->cfg.hidden);
        for (int i = 0; i < count; ++i) B200_REQUIRE(inboxes[i], kErrBadArg, "b200_llama_set_peer_inboxes: null inbox %d", i);
        m->peers.assign(inboxes, inboxes + count);
    }
    cudaMemset(m->partial, 0, (size_t)m->cfg.max_num_seqs * m->cfg.hidden * sizeof(float));     // the fused kernel keeps it zeroed from here on
    if (m->h_timeout) *m->h_timeout = 0u;
    invalidate_graphs(m);                                        // the forward changes shape
}

int32_t b200_llama_peer_timeouts(b200_llama* m) {
    if (!m || !m->h_timeout) return 0;
    return (int32_t)*reinterpret_cast<volatile uint32_t*>(m->h_timeout);
}

void b200_llama_set_comm(b200_llama* m, void* nccl_comm) {
    B200_REQUIRE(m, kErrBadArg, "b200_llama_set_comm: null model");
    m->comm = nccl_comm;
    invalidate_graphs(m);
}

// full-vocabulary logits [n, vocab] on the device: the local buffer at TP = 1; vocab-parallel (VocabParallelLinear::forward,
// /root/reference/src/openai/distributed.rs:1632-1667): all-gather the [n, vocab_l] shards -> [world, n, vocab_l] -> transpose ->
// [n, world * vocab_l] -> narrow to the original vocab.  A collective: every rank must ask for logits in the same step.
static const float* full_logits(b200_llama* m, int n, cudaStream_t st) {
    if (m->cfg.tp_world == 1) return m->logits;
    tp_allgather_bytes(m->comm, m->logits, m->logits_gathered, (size_t)n * m->vocab_l * sizeof(float), st);
    gather_logits_transpose(m->logits_gathered, m->logits_full, m->cfg.tp_world, n, m->vocab_l, m->cfg.vocab, st);
    m->launches += 1;
    return m->logits_full;
}

void b200_llama_decode(b200_llama* m, const uint32_t* tokens, const int64_t* positions,
                       const int64_t* slot_mapping, const uint32_t* context_lens,
                       const uint32_t* block_tables, int32_t table_width, int32_t num_seqs,
                       int32_t* next_tokens_host, float* logits_host, int64_t stream) {
    if (!ready(m)) return;
    const b200_llama_config& c = m->cfg;
    B200_REQUIRE(tokens && positions && slot_mapping && context_lens && block_tables, kErrBadArg, "b200_llama_decode: null metadata");
    B200_REQUIRE(num_seqs > 0 && num_seqs <= c.max_num_seqs, kErrBadArg, "b200_llama_decode: num_seqs %d out of (0, %d]", num_seqs, c.max_num_seqs);
    B200_REQUIRE(table_width > 0 && table_width <= c.max_blocks_per_seq, kErrBadArg,
                 "b200_llama_decode: block table width %d out of (0, %d]", table_width, c.max_blocks_per_seq);
    cudaStream_t st = as_stream(stream);
    const int B = num_seqs, W = c.max_blocks_per_seq;
    const int64_t slots_total = m->num_blocks * (int64_t)c.block_size;
    for (int b = 0; b < B; ++b) {
        B200_REQUIRE(tokens[b] < (uint32_t)c.vocab, kErrBadArg, "b200_llama_decode: token id %u >= vocab", tokens[b]);
        B200_REQUIRE(positions[b] >= 0 && positions[b] < c.max_pos, kErrBadArg, "b200_llama_decode: position out of range");
        // the attention kernel walks ceil(ctx / block_size) entries of this sequence's table row
        B200_REQUIRE(context_lens[b] >= 1 && (int64_t)context_lens[b] <= (int64_t)table_width * c.block_size, kErrBadArg,
                     "b200_llama_decode: context_lens[%d] = %u exceeds the block table (%d blocks of %d)", b, context_lens[b], table_width, c.block_size);
        B200_REQUIRE(slot_mapping[b] < slots_total, kErrBadArg, "b200_llama_decode: slot_mapping[%d] = %lld beyond the cache (%lld slots)", b,
                     (long long)slot_mapping[b], (long long)slots_total);
    }
    // stage into pinned memory (same layout as the device slab); block tables are re-padded with zeros to the static
    // width (graph.rs:732-738).  Two slabs alternate; a slab is reused only after the copy that read it has finished.
    const int si = m->stage_idx;
    m->stage_idx ^= 1;
    cudaEventSynchronize(m->stage_ev[si]);                       // no-op unless this slab's previous copy is still in flight
    char* stage = m->h_stage[si];
    const size_t Bp = ((size_t)c.max_num_seqs + 3) & ~(size_t)3;
    int64_t* h_tok = (int64_t*)stage;
    int64_t* h_pos = h_tok + Bp;
    int64_t* h_slot = h_pos + Bp;
    uint32_t* h_ctx = (uint32_t*)(h_slot + Bp);
    uint32_t* h_tab = h_ctx + Bp;
    for (int b = 0; b < B; ++b) {
        h_tok[b] = tokens[b]; h_pos[b] = positions[b]; h_slot[b] = slot_mapping[b]; h_ctx[b] = context_lens[b];
        for (int j = 0; j < W; ++j) h_tab[(size_t)b * W + j] = j < table_width ? block_tables[(size_t)b * table_width + j] : 0u;
    }
    cudaMemcpyAsync(m->d_meta, stage, (size_t)((char*)(h_tab + (size_t)B * W) - stage), cudaMemcpyHostToDevice, st);
    cudaEventRecord(m->stage_ev[si], st);
    run_step(m, B, st);
    if (logits_host) {
        const float* src = full_logits(m, B, st);
        cudaMemcpyAsync(logits_host, src, (size_t)B * c.vocab * 4, cudaMemcpyDeviceToHost, st);
    }
    if (next_tokens_host) cudaMemcpyAsync(m->h_next, m->next_tokens, 4 * B, cudaMemcpyDeviceToHost, st);
    if (next_tokens_host || logits_host) {
        cudaError_t e = cudaStreamSynchronize(st);                 // graph.rs:297-301
        if (e != cudaSuccess) { set_error(kErrCuda, "b200_llama_decode: %s", cudaGetErrorString(e)); return; }
        if (peer_timed_out(m, "b200_llama_decode")) return;
        if (next_tokens_host) for (int b = 0; b < B; ++b) next_tokens_host[b] = m->h_next[b];
    }
}

void b200_llama_decode_resident(b200_llama* m, int32_t num_seqs, int32_t advance, int64_t stream) {
    if (!ready(m)) return;
    B200_REQUIRE(num_seqs > 0 && num_seqs <= m->cfg.max_num_seqs, kErrBadArg, "b200_llama_decode_resident: bad num_seqs");
    cudaStream_t st = as_stream(stream);
    if (advance) {
        launch_pdl(advance_metadata_kernel, dim3(ceil_div(num_seqs, 128)), dim3(128), 0, st, m->d_tokens, (const int32_t*)m->next_tokens,
                   m->d_positions, m->d_slots, m->d_ctx, (const uint32_t*)m->d_tables, (int)num_seqs, (int)m->cfg.max_blocks_per_seq,
                   (int)m->cfg.block_size, 1);
        m->launches += 1;
    }
    run_step(m, num_seqs, st);
}

void b200_llama_linear_chain(b200_llama* m, int32_t num_seqs, int64_t stream) {
    if (!ready(m)) return;
    B200_REQUIRE(num_seqs > 0 && num_seqs <= m->cfg.max_num_seqs, kErrBadArg, "b200_llama_linear_chain: bad num_seqs");
    m->launches += forward(m, num_seqs, as_stream(stream), /*linear_only=*/true);
}

// profiling aid: copy the clock64 stamps of the traced launch ([CTAs][4 phases][8] int64) to the host; returns the CTA count
int32_t b200_llama_mega_trace(b200_llama* m, long long* host, int32_t max_ctas) {
    if (!m || !m->mega_trace || !host) return 0;
    const int n = std::min(max_ctas, m->mega_G);
    cudaDeviceSynchronize();
    cudaMemcpy(host, m->mega_trace, (size_t)n * kMegaMaxPhases * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    return n;
}

int32_t b200_llama_uses_layer_kernel(b200_llama* m, int32_t num_seqs) {
    return m && ready(m) && mega_usable(m, num_seqs) ? 1 : 0;
}

void b200_llama_read_next_tokens(b200_llama* m, int32_t* host, int32_t n, int64_t stream) {
    B200_REQUIRE(m && host && n > 0 && n <= m->cfg.max_num_seqs, kErrBadArg, "b200_llama_read_next_tokens: bad arguments");
    cudaMemcpyAsync(host, m->next_tokens, (size_t)n * 4, cudaMemcpyDeviceToHost, as_stream(stream));
    cudaError_t e = cudaStreamSynchronize(as_stream(stream));
    if (e != cudaSuccess) { set_error(kErrCuda, "b200_llama_read_next_tokens: %s", cudaGetErrorString(e)); return; }
    peer_timed_out(m, "b200_llama_read_next_tokens");
}

void b200_llama_read_logits(b200_llama* m, float* host, int32_t n, int64_t stream) {
    B200_REQUIRE(m && host && n > 0 && n <= m->cfg.max_num_seqs, kErrBadArg, "b200_llama_read_logits: bad arguments");
    const float* src = full_logits(m, n, as_stream(stream));
    cudaMemcpyAsync(host, src, (size_t)n * m->cfg.vocab * 4, cudaMemcpyDeviceToHost, as_stream(stream));
    cudaError_t e = cudaStreamSynchronize(as_stream(stream));
    if (e != cudaSuccess) { set_error(kErrCuda, "b200_llama_read_logits: %s", cudaGetErrorString(e)); return; }
    peer_timed_out(m, "b200_llama_read_logits");
}

const float* b200_llama_logits(b200_llama* m) { return m ? m->logits : nullptr; }
const int32_t* b200_llama_next_tokens(b200_llama* m) { return m ? m->next_tokens : nullptr; }
int64_t b200_llama_kernel_launches(b200_llama* m) { return m ? m->launches : 0; }

}  // extern "C"
