/*
 * oracle/cpu_ref.c -- C restatement of the reference's CPU (candle/GGML) arithmetic for the
 * batched-decode hot path.  TEST INFRASTRUCTURE + CPU BASELINE ONLY (see oracle/__init__.py):
 * used by tests/ as a checker, and by bench.py's cpu_baseline / --impl reference legs as the
 * timed CPU arm.  The product never links or calls this file.
 *
 * The reference's CPU path cannot be compiled here (no cargo/rustc; the arithmetic lives in
 * candle-core fork @cafd231 v0.8.3 and attention-rs @a97f519 v0.6.5, neither vendored under
 * /root/reference), so kind = "port".  PARITY UNPINNED except the block codecs (pinned against
 * gguf-py through tests/golden).  What is restated and from where:
 *   - Q4_K / Q6_K block formats and Q8_K activation quantisation + integer dot: candle
 *     k_quants.rs [UPSTREAM], itself a transliteration of llama.cpp ggml-quants.c
 *     (quantize_row_q8_K_ref, ggml_vec_dot_q4_K_q8_K, ggml_vec_dot_q6_K_q8_K); formats in
 *     SURVEY.md Appendix A.  Call site: QMatMul::forward src/openai/models/linear.rs:765-806.
 *   - op order / dtypes of one decoder layer: src/openai/models/quantized_llama.rs:424-506,
 *     src/openai/models/layers/attention.rs:910-1011, Mlp::forward quantized_llama.rs:32-44.
 *   - attention math: NaiveAttention::forward src/openai/models/mod.rs:1268-1307 over K/V
 *     gathered through block tables (src/openai/pipelines/inputs.rs:376-454).
 *   - rope_i: src/openai/models/layers/rotary_emb.rs:52-101 (interleaved pairs).
 *
 * Build: see oracle/Makefile (gcc -O3 -mavx2 -mfma -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QK_K 256
#define GGML_TYPE_Q4_K 12
#define GGML_TYPE_Q6_K 14

typedef uint16_t ggml_half;

#pragma pack(push, 1)
typedef struct { ggml_half d, dmin; uint8_t scales[12]; uint8_t qs[128]; } block_q4_K;    /* 144 */
typedef struct { uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; ggml_half d; } block_q6_K; /* 210 */
typedef struct { float d; int8_t qs[256]; int16_t bsums[16]; } block_q8_K;               /* 292 */
#pragma pack(pop)

static inline float half_to_float(ggml_half h) {
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, e = (h >> 10) & 0x1F, m = h & 0x3FF, u;
    if (e == 0) {
        if (m == 0) u = sign;
        else { e = 113; while (!(m & 0x400)) { m <<= 1; e--; } u = sign | (e << 23) | ((m & 0x3FF) << 13); }
    } else if (e == 31) u = sign | 0x7F800000u | (m << 13);
    else u = sign | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}
static inline float bf16_to_float(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
static inline uint16_t float_to_bf16(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    u += 0x7FFF + ((u >> 16) & 1);
    return (uint16_t)(u >> 16);
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py picks the thread count that is actually fastest on the box (logical CPUs can exceed the cores the container may use) */
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}


/* ---- Q8_K activation quantisation (candle BlockQ8K::from_float [UPSTREAM]) ---------------- */
void ref_quantize_row_q8_K(const float* x, void* vy, int k) {
    block_q8_K* y = (block_q8_K*)vy;
    for (int b = 0; b < k / QK_K; b++, x += QK_K) {
        float max = 0, amax = 0;
        for (int j = 0; j < QK_K; j++) { float ax = fabsf(x[j]); if (ax > amax) { amax = ax; max = x[j]; } }
        if (amax == 0) { y[b].d = 0; memset(y[b].qs, 0, QK_K); memset(y[b].bsums, 0, 32); continue; }
        const float iscale = -128.f / max;
        for (int j = 0; j < QK_K; j++) {
            int v = (int)roundf(iscale * x[j]);
            y[b].qs[j] = (int8_t)(v > 127 ? 127 : v);
        }
        for (int j = 0; j < 16; j++) {
            int s = 0;
            for (int i = 0; i < 16; i++) s += y[b].qs[j * 16 + i];
            y[b].bsums[j] = (int16_t)s;
        }
        y[b].d = 1.f / iscale;
    }
}

static inline void get_scale_min_k4(int j, const uint8_t* q, uint8_t* d, uint8_t* m) {
    if (j < 4) { *d = q[j] & 63; *m = q[j + 4] & 63; }
    else { *d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); *m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

/* ---- integer dots (ggml_vec_dot_q4_K_q8_K / ggml_vec_dot_q6_K_q8_K, scalar form) ---------- */
float ref_vec_dot_q4_K_q8_K(int n, const void* vx, const void* vy) {
    const block_q4_K* x = (const block_q4_K*)vx;
    const block_q8_K* y = (const block_q8_K*)vy;
    float sumf = 0;
    for (int b = 0; b < n / QK_K; b++) {
        const uint8_t* q4 = x[b].qs;
        const int8_t* q8 = y[b].qs;
        int32_t isum = 0, msum = 0;
        for (int c = 0; c < 4; c++) {
            uint8_t sc0, m0, sc1, m1;
            get_scale_min_k4(2 * c, x[b].scales, &sc0, &m0);
            get_scale_min_k4(2 * c + 1, x[b].scales, &sc1, &m1);
            int32_t s0 = 0, s1 = 0;
            for (int l = 0; l < 32; l++) {
                s0 += (int16_t)(q4[l] & 0xF) * (int16_t)q8[l];
                s1 += (int16_t)(q4[l] >> 4) * (int16_t)q8[32 + l];
            }
            isum += s0 * sc0 + s1 * sc1;
            msum += m0 * (y[b].bsums[4 * c] + y[b].bsums[4 * c + 1]) +
                    m1 * (y[b].bsums[4 * c + 2] + y[b].bsums[4 * c + 3]);
            q4 += 32; q8 += 64;
        }
        const float d = y[b].d * half_to_float(x[b].d), dm = y[b].d * half_to_float(x[b].dmin);
        sumf += d * (float)isum - dm * (float)msum;
    }
    return sumf;
}

float ref_vec_dot_q6_K_q8_K(int n, const void* vx, const void* vy) {
    const block_q6_K* x = (const block_q6_K*)vx;
    const block_q8_K* y = (const block_q8_K*)vy;
    float sumf = 0;
    for (int b = 0; b < n / QK_K; b++) {
        int8_t a[QK_K];
        const uint8_t* ql = x[b].ql; const uint8_t* qh = x[b].qh;
        for (int h = 0; h < 2; h++, ql += 64, qh += 32) {
            int8_t* o = a + 128 * h;
            for (int l = 0; l < 32; l++) {
                o[l]      = (int8_t)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                o[l + 32] = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                o[l + 64] = (int8_t)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                o[l + 96] = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
            }
        }
        int32_t isum = 0;
        for (int g = 0; g < 16; g++) {
            int32_t s = 0;
            for (int l = 0; l < 16; l++) s += (int16_t)a[16 * g + l] * (int16_t)y[b].qs[16 * g + l];
            isum += s * x[b].scales[g];
        }
        sumf += y[b].d * half_to_float(x[b].d) * (float)isum;
    }
    return sumf;
}

/* y[m,n] = QMatMul(x[m,k]): activations -> Q8_K per row, integer dot per weight row. */
int ref_qmatmul_q8k(const float* x, int m, const uint8_t* w, int ggml_type, int n, int k, float* y) {
    if (k % QK_K) return -1;
    const int nbk = k / QK_K;
    const size_t wrow = (size_t)nbk * (ggml_type == GGML_TYPE_Q4_K ? 144 : 210);
    if (ggml_type != GGML_TYPE_Q4_K && ggml_type != GGML_TYPE_Q6_K) return -2;
    block_q8_K* xq = (block_q8_K*)malloc((size_t)m * nbk * sizeof(block_q8_K));
    #pragma omp parallel for schedule(static)
    for (int r = 0; r < m; r++) ref_quantize_row_q8_K(x + (size_t)r * k, xq + (size_t)r * nbk, k);
    #pragma omp parallel for schedule(static)
    for (int j = 0; j < n; j++) {
        const uint8_t* wr = w + (size_t)j * wrow;
        for (int r = 0; r < m; r++) {
            const block_q8_K* xr = xq + (size_t)r * nbk;
            y[(size_t)r * n + j] = ggml_type == GGML_TYPE_Q4_K ? ref_vec_dot_q4_K_q8_K(k, wr, xr)
                                                              : ref_vec_dot_q6_K_q8_K(k, wr, xr);
        }
    }
    free(xq);
    return 0;
}

/* ---- elementwise pieces -------------------------------------------------------------------- */
void ref_rms_norm(const float* x, const float* w, float eps, int rows, int n, float* y) {
    #pragma omp parallel for
    for (int r = 0; r < rows; r++) {
        double ss = 0;
        for (int i = 0; i < n; i++) ss += (double)x[(size_t)r * n + i] * x[(size_t)r * n + i];
        const float sc = 1.0f / sqrtf((float)(ss / n) + eps);
        for (int i = 0; i < n; i++) y[(size_t)r * n + i] = x[(size_t)r * n + i] * sc * w[i];
    }
}

/* interleaved RoPE (rope_i) in place on x[T, h, hd]; cos/sin [max_pos, hd/2] */
void ref_rope_i(float* x, const float* cos_t, const float* sin_t, const int64_t* pos, int T, int h, int hd) {
    #pragma omp parallel for
    for (int t = 0; t < T; t++)
        for (int j = 0; j < h; j++) {
            float* p = x + ((size_t)t * h + j) * hd;
            const float* c = cos_t + (size_t)pos[t] * (hd / 2);
            const float* s = sin_t + (size_t)pos[t] * (hd / 2);
            for (int i = 0; i < hd / 2; i++) {
                const float a = p[2 * i], b = p[2 * i + 1];
                p[2 * i] = a * c[i] - b * s[i];
                p[2 * i + 1] = a * s[i] + b * c[i];
            }
        }
}

void ref_silu_mul(const float* g, const float* u, size_t n, float* y) {
    #pragma omp parallel for
    for (size_t i = 0; i < n; i++) y[i] = g[i] / (1.f + expf(-g[i])) * u[i];
}

/* k,v f32 [T,kvh,hd] -> bf16 caches [nb,bs,kvh,hd]; slot<0 skipped */
void ref_reshape_and_cache_flash_bf16(const float* k, const float* v, uint16_t* kc, uint16_t* vc,
                                      const int64_t* slots, int T, int kvh, int hd) {
    const size_t row = (size_t)kvh * hd;
    for (int t = 0; t < T; t++) {
        if (slots[t] < 0) continue;
        for (size_t i = 0; i < row; i++) {
            kc[(size_t)slots[t] * row + i] = float_to_bf16(k[(size_t)t * row + i]);
            vc[(size_t)slots[t] * row + i] = float_to_bf16(v[(size_t)t * row + i]);
        }
    }
}

/* decode attention: q f32 (bf16-representable) [B,H,hd]; caches bf16 [nb,bs,kvh,hd];
 * block_tables u32 [B,max_blocks]; context_lens u32 [B]; out f32 [B,H,hd] rounded to bf16. */
void ref_paged_attention_decode_bf16(const float* q, const uint16_t* kc, const uint16_t* vc,
                                     const uint32_t* block_tables, const uint32_t* context_lens,
                                     int B, int H, int kvh, int hd, int bs, int max_blocks,
                                     float scale, float* out) {
    const int rep = H / kvh;
    #pragma omp parallel for collapse(2) schedule(dynamic)
    for (int b = 0; b < B; b++)
        for (int g = 0; g < kvh; g++) {
            const int L = (int)context_lens[b];
            float m[16], l[16];
            float* acc = (float*)calloc((size_t)rep * hd, sizeof(float));
            for (int r = 0; r < rep; r++) { m[r] = -INFINITY; l[r] = 0; }
            for (int t = 0; t < L; t++) {
                const size_t slot = (size_t)block_tables[(size_t)b * max_blocks + t / bs] * bs + t % bs;
                const uint16_t* kr = kc + (slot * kvh + g) * hd;
                const uint16_t* vr = vc + (slot * kvh + g) * hd;
                float kf[512], vf[512];
                for (int d = 0; d < hd; d++) { kf[d] = bf16_to_float(kr[d]); vf[d] = bf16_to_float(vr[d]); }
                for (int r = 0; r < rep; r++) {
                    const float* qr = q + ((size_t)b * H + g * rep + r) * hd;
                    float s = 0;
                    for (int d = 0; d < hd; d++) s += qr[d] * kf[d];
                    s *= scale;
                    const float mn = s > m[r] ? s : m[r];
                    const float corr = expf(m[r] - mn), p = expf(s - mn);
                    float* a = acc + (size_t)r * hd;
                    for (int d = 0; d < hd; d++) a[d] = a[d] * corr + p * vf[d];
                    l[r] = l[r] * corr + p; m[r] = mn;
                }
            }
            for (int r = 0; r < rep; r++) {
                float* o = out + ((size_t)b * H + g * rep + r) * hd;
                for (int d = 0; d < hd; d++)
                    o[d] = L > 0 ? bf16_to_float(float_to_bf16(acc[(size_t)r * hd + d] / l[r])) : 0.f;
            }
            free(acc);
        }
}

/* copy_blocks over host pointers (src/backend/cache.rs:103-162 semantics) */
void ref_copy_blocks(void** key_ptrs, void** val_ptrs, const int64_t* mapping, int layers, int pairs,
                     size_t bytes_per_block) {
    for (int l = 0; l < layers; l++)
        for (int p = 0; p < pairs; p++) {
            const int64_t s = mapping[2 * p], d = mapping[2 * p + 1];
            memcpy((char*)key_ptrs[l] + d * bytes_per_block, (char*)key_ptrs[l] + s * bytes_per_block, bytes_per_block);
            memcpy((char*)val_ptrs[l] + d * bytes_per_block, (char*)val_ptrs[l] + s * bytes_per_block, bytes_per_block);
        }
}

/* ---- one decoder layer / lm head, the timed CPU arm ---------------------------------------- */
typedef struct {
    int hidden, heads, kv_heads, head_dim, ffn, block_size, max_blocks;
    float rms_eps;
} ref_cfg;

typedef struct {
    const float *attn_norm, *ffn_norm;
    const uint8_t *wq, *wk, *wv, *wo, *w1, *w2, *w3;   /* Q4_K rows */
} ref_layer;

/* x f32 [B,hidden] updated in place.  scratch: >= B*(3*hidden + 2*ffn + 2*kv) floats */
int ref_llama_layer_decode(const ref_cfg* c, const ref_layer* w, float* x, int B,
                           const int64_t* positions, const int64_t* slots,
                           const uint32_t* block_tables, const uint32_t* context_lens,
                           uint16_t* kcache, uint16_t* vcache, const float* cos_t, const float* sin_t,
                           float* scratch) {
    const int H = c->hidden, qd = c->heads * c->head_dim, kd = c->kv_heads * c->head_dim, F = c->ffn;
    float* h = scratch;               float* q = h + (size_t)B * H;
    float* k = q + (size_t)B * qd;    float* v = k + (size_t)B * kd;
    float* att = v + (size_t)B * kd;  float* g = att + (size_t)B * qd;
    float* u = g + (size_t)B * F;     float* o = u + (size_t)B * F;
    ref_rms_norm(x, w->attn_norm, c->rms_eps, B, H, h);
    if (ref_qmatmul_q8k(h, B, w->wq, GGML_TYPE_Q4_K, qd, H, q)) return -1;
    ref_qmatmul_q8k(h, B, w->wk, GGML_TYPE_Q4_K, kd, H, k);
    ref_qmatmul_q8k(h, B, w->wv, GGML_TYPE_Q4_K, kd, H, v);
    ref_rope_i(q, cos_t, sin_t, positions, B, c->heads, c->head_dim);
    ref_rope_i(k, cos_t, sin_t, positions, B, c->kv_heads, c->head_dim);
    for (size_t i = 0; i < (size_t)B * qd; i++) q[i] = bf16_to_float(float_to_bf16(q[i]));
    ref_reshape_and_cache_flash_bf16(k, v, kcache, vcache, slots, B, c->kv_heads, c->head_dim);
    ref_paged_attention_decode_bf16(q, kcache, vcache, block_tables, context_lens, B, c->heads,
                                    c->kv_heads, c->head_dim, c->block_size, c->max_blocks,
                                    1.0f / sqrtf((float)c->head_dim), att);
    ref_qmatmul_q8k(att, B, w->wo, GGML_TYPE_Q4_K, H, qd, o);
    for (size_t i = 0; i < (size_t)B * H; i++) x[i] += o[i];
    ref_rms_norm(x, w->ffn_norm, c->rms_eps, B, H, h);
    ref_qmatmul_q8k(h, B, w->w1, GGML_TYPE_Q4_K, F, H, g);
    ref_qmatmul_q8k(h, B, w->w3, GGML_TYPE_Q4_K, F, H, u);
    ref_silu_mul(g, u, (size_t)B * F, g);
    ref_qmatmul_q8k(g, B, w->w2, GGML_TYPE_Q4_K, H, F, o);
    for (size_t i = 0; i < (size_t)B * H; i++) x[i] += o[i];
    return 0;
}

size_t ref_layer_scratch_floats(const ref_cfg* c, int B) {
    return (size_t)B * (2 * (size_t)c->hidden + 2 * (size_t)c->heads * c->head_dim +
                        2 * (size_t)c->kv_heads * c->head_dim + 2 * (size_t)c->ffn) + 64;
}
