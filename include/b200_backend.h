/*
 * b200_backend.h -- C ABI of libb200backend.so: the B200-native (sm_100a) batched-decode backend
 * that drops in behind candle-vllm's `src/backend` / attention-rs FFI surface.
 *
 * Conventions (identical to the reference's `attention_rs::kernels::ffi`):
 *   - extern "C", plain pointers and sizes, `int64_t stream` = the caller's CUstream
 *     (`*dev.cu_stream() as i64`, /root/reference/src/backend/cache.rs:135, gptq.rs:130);
 *   - every entry point is stream-ordered, allocation-free, host-sync-free and therefore
 *     CUDA-graph capture safe (/root/reference/src/backend/graph.rs:271-274); scratch memory is
 *     a caller-provided `workspace` whose size comes from the matching `*_workspace_bytes`;
 *   - functions return void like the reference's; argument errors (the reference validates on the
 *     Rust side and `bail!`s, cache.rs:26-39, gptq.rs:232-238) are recorded per thread and read
 *     with b200_last_error(); device errors stay sticky CUDA errors, as in the reference;
 *   - outputs are caller-allocated, possibly uninitialised (gptq.rs:66), inputs may be views.
 *
 * There is no CPU fallback: on a machine without an sm_100 GPU every launch fails loudly
 * (b200_last_error() != 0 / CUDA error), it never computes on the host.
 */
#ifndef B200_BACKEND_H_
#define B200_BACKEND_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library / error state ------------------------------------------------------------- */
int         b200_abi_version(void);
/* 0 = ok.  Last argument/launch error recorded on THIS thread; reading clears it. */
int         b200_last_error(void);
const char* b200_last_error_message(void);
/* SM count / compute capability of the current device (0 if no usable device). */
int         b200_device_sm_count(void);
int         b200_device_cc(void);

/* data-type tags used by the entry points below */
enum { B200_F32 = 0, B200_F16 = 1, B200_BF16 = 2, B200_U8 = 3, B200_FP8_E4M3 = 4,
       /* fp16 in "K4" order: within every aligned group of 4 elements along the last dimension the middle
        * two are swapped ([k0,k2,k1,k3]).  This is the activation layout the tcgen05 dequant-GEMM consumes
        * (it matches the order in which two nibbles fall out of one 32-bit word of a GGML block, so the
        * weights never need a byte permute); rms_norm / silu_mul / cast / paged_attention_decode can emit it. */
       B200_F16_K4 = 5 };
/* GGML tensor types (values = ggml_type, as stored in GGUF files) */
enum { B200_GGML_Q8_0 = 8, B200_GGML_Q4_K = 12, B200_GGML_Q6_K = 14 };
/* KV layouts (/root/reference/src/scheduler/cache_engine.rs:298-341) */
enum { B200_KV_FLASH = 0 /* K,V [nb, bs, kvh, hd] */,
       B200_KV_PAGED = 1 /* K [nb, kvh, hd/x, bs, x], V [nb, kvh, hd, bs], x = 16/elem */ };

/* ---- K4: copy_blocks -- replaces attention_rs::kernels::ffi::copy_blocks_{bf16,f16,f32} ----
 * Call site: /root/reference/src/backend/cache.rs:127-162.  The first three arguments are HOST
 * pointers (Vec::as_mut_ptr, cache.rs:112-114): u64 device addresses of each layer's K / V cache,
 * and i64 (src,dst) block pairs.  For every layer, K and V: dst block <- src block,
 * numel_per_block elements each.  copy_blocks_u8 adds the arm the reference lacks for FP8 KV
 * (cache.rs:95-97 bails). */
void copy_blocks_bf16(void* key_cache_ptrs, void* value_cache_ptrs, const void* block_mapping,
                      int32_t num_layers, int32_t num_pairs, int32_t numel_per_block, int64_t stream);
void copy_blocks_f16 (void* key_cache_ptrs, void* value_cache_ptrs, const void* block_mapping,
                      int32_t num_layers, int32_t num_pairs, int32_t numel_per_block, int64_t stream);
void copy_blocks_f32 (void* key_cache_ptrs, void* value_cache_ptrs, const void* block_mapping,
                      int32_t num_layers, int32_t num_pairs, int32_t numel_per_block, int64_t stream);
void copy_blocks_u8  (void* key_cache_ptrs, void* value_cache_ptrs, const void* block_mapping,
                      int32_t num_layers, int32_t num_pairs, int32_t numel_per_block, int64_t stream);

/* ---- K5: swap_blocks -- replaces attention_rs::cache::swap_blocks(src, dst, &HashMap) -------
 * Call site: /root/reference/src/scheduler/cache_engine.rs:527-535 (swap_in :345-363, swap_out
 * :365-385).  dst[dst_block] <- src[src_block], bytes_per_block each.  `mapping` is a HOST array
 * of i64 (src,dst) pairs.  One of src/dst may be host memory (pinned or pageable); the copy kind
 * is inferred (cudaMemcpyDefault).  Adjacent pairs are coalesced into single copies. */
void swap_blocks(const void* src, void* dst, const int64_t* mapping, int32_t num_pairs,
                 int64_t bytes_per_block, int64_t stream);

/* ---- FlashInfer-style page tables (the reference's default build): indptr u32 [B + 1], indices u32 [nnz] (physical block ids, sequence by
 * sequence), last_len u32 [B] (tokens in each sequence's last page), as built at /root/reference/src/openai/pipelines/inputs.rs:477-506 ->
 * block_tables u32 [B, max_blocks_per_seq] (0-padded) and context_lens u32 [B] = (pages - 1) * block_size + last_len (inputs.rs:523-531), on
 * the device and stream-ordered, so a host that only has CSR metadata (graph.rs:471-803 static buffers) can drive every entry point below. */
void flashinfer_csr_to_paged(const uint32_t* indptr, const uint32_t* indices, const uint32_t* last_len, uint32_t* block_tables, uint32_t* context_lens,
                             int32_t num_seqs, int32_t max_blocks_per_seq, int32_t block_size, int64_t stream);

/* ---- K3: reshape_and_cache -- the cache write inside PagedAttention::forward ----------------
 * Call sites: /root/reference/src/openai/models/layers/attention.rs:707-718, :983-994; slot math
 * /root/reference/src/openai/pipelines/inputs.rs:180-194, :410-423.
 * key/value: [num_tokens, num_kv_heads, head_dim] of `in_dtype` (B200_F32/F16/BF16), row strides
 * in elements.  slot_mapping i64[num_tokens]: flat slot = block*block_size + offset; negative
 * (pad, -1) = skip.  cache_dtype: same 16-bit type as the model dtype, or B200_FP8_E4M3 (stored as
 * U8, saturating RNE cast, scale 1.0 -- the reference passes no scale).  layout: B200_KV_*. */
void reshape_and_cache(const void* key, const void* value, void* key_cache, void* value_cache,
                       const int64_t* slot_mapping, int32_t num_tokens, int32_t num_kv_heads,
                       int32_t head_dim, int32_t block_size, int64_t key_stride, int64_t value_stride,
                       int32_t in_dtype, int32_t cache_dtype, int32_t layout, int64_t stream);

/* ---- K1: paged attention, decode -- PagedAttention::forward with is_prefill = false ---------
 * Call sites as above; metadata contract /root/reference/src/openai/pipelines/inputs.rs:552-568.
 * q, out: [num_seqs, num_heads, head_dim] of `dtype` (B200_BF16 / B200_F16), contiguous.
 * block_tables u32 [num_seqs, max_blocks_per_seq] (0-padded), context_lens u32 [num_seqs]
 * INCLUDING the token being decoded.  The kernel reads only device-side context_lens (never a
 * host max_context_len), so a graph captured with padded tables replays correctly (graph.rs:604).
 * softcap <= 0: none.  sliding_window <= 0: none.  out_dtype: `dtype`, or B200_F16 to hand the
 * result straight to a quantised GEMM (values are rounded to `dtype` first).
 * workspace: >= paged_attention_decode_workspace_bytes(...) bytes of device memory, zero-filled once
 * before its first use (it holds the work-queue head, which every call leaves at zero again). */
size_t paged_attention_decode_workspace_bytes(int32_t num_seqs, int32_t num_heads, int32_t head_dim,
                                              int32_t max_blocks_per_seq, int32_t block_size);
void paged_attention_decode(void* out, const void* q, const void* key_cache, const void* value_cache,
                            const uint32_t* block_tables, const uint32_t* context_lens,
                            int32_t num_seqs, int32_t num_heads, int32_t num_kv_heads, int32_t head_dim,
                            int32_t block_size, int32_t max_blocks_per_seq, int64_t num_blocks,
                            float scale, float softcap, int32_t sliding_window,
                            int32_t dtype, int32_t cache_dtype, int32_t layout, int32_t out_dtype,
                            void* workspace, size_t workspace_bytes, int64_t stream);

/* ---- K2: paged attention, (chunked) prefill -- is_prefill = true ----------------------------
 * Metadata /root/reference/src/openai/pipelines/inputs.rs:133-148, :351-367.  Varlen causal
 * attention; ALL keys/values (cached prefix + this chunk) are read from the paged cache, which the
 * caller has already written with reshape_and_cache.  Sequence i owns q rows
 * cu_seqlens_q[i]..cu_seqlens_q[i+1] = the last q_len positions of its k_len context. */
void paged_attention_prefill(void* out, const void* q, const void* key_cache, const void* value_cache,
                             const uint32_t* block_tables, const uint32_t* cu_seqlens_q,
                             const uint32_t* cu_seqlens_k, int32_t num_seqs, int32_t total_q,
                             int32_t max_seqlen_q, int32_t num_heads, int32_t num_kv_heads,
                             int32_t head_dim, int32_t block_size, int32_t max_blocks_per_seq,
                             float scale, float softcap, int32_t sliding_window,
                             int32_t dtype, int32_t cache_dtype, int32_t layout, int64_t stream);

/* ---- K6: QMatMul::forward on GGUF tensors -- replaces candle QMatMul (quantized.cu) ---------
 * Call sites: /root/reference/src/openai/models/linear.rs:765-806,
 * /root/reference/src/openai/models/quantized_llama.rs:33-37, layers/attention.rs:920-922,1004.
 * y[m,n] (f32) = x[m,k] (f32) . dequant(W[n,k])^T.  W = GGML blocks, row-major over n, verbatim
 * GGUF bytes (k % 256 == 0 for K-quants, % 32 for Q8_0).  Activations are rounded to fp16
 * (saturating) and products accumulate in fp32 on the tensor cores; the reference quantises
 * activations to 8 bit (Q8_1 / Q8_K), which is coarser.
 * accumulate != 0: y += result (used for the residual add; implies fp32 atomics).
 * m <= 64 (decode) runs the dequant-into-TMEM kernel; 64 < m < 512 runs it 64 rows per pass; m >= 512 (prefill chunks) dequantises W once
 * into the workspace (fp16, same single rounding) and runs the dense tcgen05 GEMM, so every m-tile reuses the dequantised tile. */
size_t qmatmul_workspace_bytes(int32_t m, int32_t n, int32_t k);
void qmatmul_f32(const float* x, const void* w, float* y, int32_t m, int32_t n, int32_t k,
                 int32_t ggml_type, int32_t accumulate, void* workspace, size_t workspace_bytes,
                 int64_t stream);
/* same with activations already in fp16 [m,k] in B200_F16_K4 order (what the fused decode layer feeds).
 * When accumulate == 0 and the product is split over K, y is zero-filled first (contiguous y only). */
void qmatmul_f16act(const void* x_f16, const void* w, float* y, int32_t m, int32_t n, int32_t k,
                    int32_t ggml_type, int32_t accumulate, int64_t stream);
/* Atomic-free, bitwise-deterministic form used by the fused decode layer: when the product of a small-n matrix is split
 * over K between SMs, every SM that shares a 128-row tile stores its partial sum to its own slab instead of red.add-ing
 * into y.  y_slabs = [slabs][m][n] f32 (slab stride m*n); returns the number of slabs S <= slabs_avail that now hold
 * partial sums (every element of slabs 0..S-1 is written, nothing needs pre-zeroing); the product is their sum, which the
 * consumer kernel (RMSNorm, RoPE, SiLU) folds into its own load.  qmatmul_slab_count() gives the S this (n, k) needs.
 * Returns 0 and records an error on bad arguments. */
int32_t qmatmul_slab_count(int32_t m, int32_t n, int32_t k, int32_t ggml_type);
int32_t qmatmul_f16act_slabs(const void* x_f16, const void* w, float* y_slabs, int32_t slabs_avail, int32_t m, int32_t n,
                             int32_t k, int32_t ggml_type, int64_t stream);
/* QTensor::dequantize: W -> f32 [n,k] (linear.rs:808-842 forward_via_dequant) */
void dequantize_f32(const void* w, float* out, int64_t n, int64_t k, int32_t ggml_type, int64_t stream);

/* ---- K7 / K9: GPTQ -> Marlin int4 weight-only GEMM -- replaces attention_rs::kernels::ffi::{gptq_repack,
 * marlin_4bit_f16, marlin_4bit_bf16, marlin_awq_4bit_*, awq_repack, gemm_half_q_half_alt} with the reference's exact
 * signatures (/root/reference/src/backend/gptq.rs:115-197, :313-332).
 * gptq_repack: qweight u32 [k_packed = K/8, n] (8 nibbles along K per word) -> same word count in this library's
 *   private int4 layout (the reference reshapes the result to [K/16, 2n], gptq.rs:283-297; only marlin_4bit_* reads it).
 * marlin_4bit_{f16,bf16}: out[m,n] = x[m,k] . ((q - 8) * scale)^T; scales [k/group, n] in the order produced by the
 *   reference's marlin_permute_scales (/root/reference/src/openai/models/linear.rs:354-379); qzeros ignored (symmetric),
 *   g_idx is accepted and NOT read (the reference passes the checkpoint's trivial k / group_size sequence on this path, linear.rs:298-337;
 *   act-order goes to gemm_half_q_half_alt); group_size 64 / 128 / -1; k % 256 == 0; n % 64 == 0; any m (64 rows per tensor-core pass).  `workspace` (n zeroed u32 of
 *   locks in Marlin) is unused.  The fp16 copy of x lives in a library-owned scratch buffer per (device, stream), grown outside
 *   stream capture, or in the per-device buffer handed over with b200_set_scratch().
 * awq_repack: AWQ qweight u32 [k, n_packed = N/8] (nibble i of a word = column 8j + [0,2,4,6,1,3,5,7][i]) -> the same private layout.
 * marlin_awq_4bit_{f16,bf16}: as marlin_4bit_* with zero points: out = x . ((q - z) * scale)^T; qzeros u32 [k/group, n/8] in the
 *   layout the reference's offline converter writes (examples/convert_awq_marlin.py:75-113: scale_perm inside 64-column blocks,
 *   [0,2,4,6,1,3,5,7] interleave inside 8, packed along columns).
 * gemm_half_q_half_alt: conventional GPTQ (act-order and / or asymmetric, 4 or 8 bit; no repack): x f16 [m,k], qweight u32
 *   [k/pack, n] packed along k, qzeros u32 [G, n/pack] packed along n and stored minus one, scales f16 [G, n], g_idx i32 [k]
 *   (required), out f16 [m,n]; shape-generic SIMT kernel. */
void gptq_repack(const void* in, void* out, int32_t k_packed, int32_t n, int64_t stream);
/* Pre-repacked "marlin" checkpoints (checkpoint_format == "marlin": tensors `B` u32 [K/16, 2N] and `s`, linear.rs:219-251) carry the Marlin project's
 * tile order, which the reference hands to marlin_4bit_* unchanged.  This library's GEMM reads its own row-major int4 layout, so such a checkpoint
 * needs ONE extra load-time call: marlin_checkpoint_repack(B) -> the layout marlin_4bit_* reads (same word count; scratch_gptq = k/8 * n words of
 * scratch; `s` is used as is: it already is in marlin_permute_scales order).  k % 64 == 0, n % 64 == 0. */
void marlin_checkpoint_repack(const void* marlin_b, void* out, void* scratch_gptq, int32_t k, int32_t n, int64_t stream);
void awq_repack(const void* in, void* out, int32_t k, int32_t n_packed, int32_t bits, int64_t stream);
void marlin_4bit_f16(const void* x, const int32_t* qweight, const void* scales, const void* qzeros, const void* g_idx,
                     void* out, int32_t m, int32_t k, int32_t n, const void* workspace, int32_t group_size, int64_t stream);
void marlin_4bit_bf16(const void* x, const int32_t* qweight, const void* scales, const void* qzeros, const void* g_idx,
                      void* out, int32_t m, int32_t k, int32_t n, const void* workspace, int32_t group_size, int64_t stream);
void marlin_awq_4bit_f16(const void* x, const int32_t* qweight, const void* scales, const void* qzeros, const void* g_idx,
                         void* out, int32_t m, int32_t k, int32_t n, const void* workspace, int32_t group_size, int64_t stream);
void marlin_awq_4bit_bf16(const void* x, const int32_t* qweight, const void* scales, const void* qzeros, const void* g_idx,
                          void* out, int32_t m, int32_t k, int32_t n, const void* workspace, int32_t group_size, int64_t stream);
void gemm_half_q_half_alt(const void* x, const uint32_t* qweight, const uint32_t* qzeros, const void* scales,
                          const int32_t* g_idx, void* out, int32_t m, int32_t n, int32_t k, int32_t bits, int64_t stream);
void b200_set_scratch(void* device_ptr, size_t bytes);

/* ---- dense 16-bit Linear::forward -- replaces candle's cuBLAS matmul for unquantised weights (/root/reference/src/openai/models/linear.rs:124-172)
 * out[m,n] = x[m,k] . weight[n,k]^T (+ bias[n]); x, weight, bias, out of `dtype` (B200_F16 / B200_BF16), contiguous, fp32 accumulation on
 * tcgen05 (128 x 256 tiles for m > 64; swap-AB weight stream for m <= 64).  k % 8 == 0, n % 8 == 0 (16-byte rows). */
void linear_16bit(const void* x, const void* weight, const void* bias, void* out, int32_t m, int32_t n, int32_t k, int32_t dtype, int64_t stream);

/* ---- K10 / K11 / K12: weight-only low-precision float linears -- replace attention_rs::{fp8_linear::fp8_matmul,
 * nvfp4_linear::nvfp4_matmul, mxfp4_linear::mxfp4_matmul} (call sites /root/reference/src/openai/models/linear.rs:1190-1221,
 * :1913-1943, :1717-1757; tensor layouts :944-973, :1812-1853, :1686-1700).
 * out[m,n] = x[m,k] . dequant(W[n,k])^T (+ bias[n]); x, bias, out are f16 or bf16 (`dtype`), fp32 accumulation.
 *   fp8_matmul : weight e4m3 bytes [n,k]; weight_scale f32 [ceil(n/block_y), ceil(k/block_x)] multiplies its tile
 *                (default tile [128,128]).  k % 256 == 0, block_x % 64 == 0 run on the tcgen05 pipeline, 64 rows per pass (weights
 *                scaled and rounded to fp16, activations fp16); other shapes on a SIMT kernel with exact fp32 weights.
 *   nvfp4_matmul: blocks u8 [n,k/2] (two e2m1 per byte, low nibble = even k), scales e4m3 [n,k/16], global_scale f32
 *                (the reciprocal / weight_scale_2 value the reference computes, linear.rs:1829-1853); input_scale is ignored
 *                (weight-only product).  k % 32 == 0.
 *   mxfp4_matmul: blocks u8 [n,k/2], scales e8m0 [n,k/32]: w = e2m1 * 2^(scale - 127).  k % 32 == 0.
 * Parity is unpinned upstream (no fixtures; the kernels live in attention-rs): the oracle (oracle/fp_formats.py) follows the
 * OCP / NVFP4 format definitions and torch.float8_e4m3fn. */
void fp8_matmul(const void* x, const void* weight, const float* weight_scale, const void* bias, void* out, int32_t m, int32_t n,
                int32_t k, int32_t block_y, int32_t block_x, int32_t dtype, int64_t stream);
void nvfp4_matmul(const void* x, const void* blocks, const void* scales, float global_scale, float input_scale, const void* bias,
                  void* out, int32_t m, int32_t n, int32_t k, int32_t dtype, int64_t stream);
void mxfp4_matmul(const void* x, const void* blocks, const void* scales, const void* bias, void* out, int32_t m, int32_t n,
                  int32_t k, int32_t dtype, int64_t stream);

/* ---- multi-head latent attention (DeepSeek-V2 / V3) -- replaces attention_rs::mla::{concat_and_cache_mla, mla_paged_decode, mla_paged_prefill}
 * (call sites /root/reference/src/openai/models/layers/mla_attention.rs:479-552; cache shapes /root/reference/src/scheduler/cache_engine.rs:172-185).
 *   concat_and_cache_mla: ckv [T, kv_lora_rank], k_pe [T, qk_rope_head_dim] -> ckv_cache [nb, bs, 1, kv_lora_rank], kpe_cache [nb, bs, 1, rope] at
 *     slot_mapping[t] (negative = pad, skipped); 16-bit `dtype`, bit copy.
 *   mla_paged_attention: "absorbed" MLA: q_absorbed [rows, H, kv_lora_rank] (W_uk already folded in), q_pe [rows, H, rope];
 *     score = (q_abs . ckv + q_pe . kpe) * sm_scale over the keys of the sequence, out [rows, H, kv_lora_rank] = softmax . ckv (the caller applies
 *     W_uv).  cu_seqlens_q == NULL: decode (one row per sequence, context_lens INCLUDING the new token); otherwise causal chunked prefill where
 *     sequence i owns rows cu_seqlens_q[i] .. cu_seqlens_q[i+1] = the last positions of its context_lens[i] keys.  Decode with latent 512 / rope 64 /
 *     block 64 runs the tensor-core split-KV kernel (workspace of mla_paged_decode_workspace_bytes, num_blocks = blocks in the cache); everything
 *     else a shape-generic kernel. */
void concat_and_cache_mla(const void* ckv, const void* k_pe, void* ckv_cache, void* kpe_cache, const int64_t* slot_mapping, int32_t num_tokens,
                          int32_t kv_lora_rank, int32_t qk_rope_head_dim, int32_t dtype, int64_t stream);
size_t mla_paged_decode_workspace_bytes(int32_t num_seqs, int32_t num_heads, int32_t max_blocks_per_seq, int32_t block_size);
void mla_paged_attention(void* out, const void* q_absorbed, const void* q_pe, const void* ckv_cache, const void* kpe_cache, const uint32_t* block_tables,
                         const uint32_t* context_lens, const uint32_t* cu_seqlens_q, int32_t num_seqs, int32_t num_rows, int32_t num_heads,
                         int32_t kv_lora_rank, int32_t qk_rope_head_dim, int32_t block_size, int32_t max_blocks_per_seq, int64_t num_blocks, float sm_scale,
                         int32_t dtype, void* workspace, size_t workspace_bytes, int64_t stream);

/* ---- fused mixture-of-experts on GGUF expert tensors -- replaces attention_rs::{topk::topk_softmax, moe::moe_gemm_gguf} and the host-side
 * sort (call sites /root/reference/src/openai/models/layers/moe.rs:35-45, :425-480, :1429-1482; quantized_qwen3_moe.rs:70-141).
 *   topk_softmax: router_logits f32 [T, E] -> topk_weights f32 [T, k] (softmax probabilities of the k largest, NOT renormalised: the
 *     caller applies norm_topk_prob / routed_scaling_factor like the reference), topk_ids u32 [T, k]; ties -> smaller expert id.
 *   sort_expert_assignments: flattened topk_ids u32 [P = T * k] -> expert_ids u32 [P] ascending, sorted_token_ids u32 [P] = the pair index
 *     (token * k + slot) at each sorted position (moe.rs:35-45; stable).
 *   moe_gemm_gguf: out f32 [P, n]; row p = W[expert of pair p] (experts = stacked GGML blocks [E, n, k]) . x[row(p)] (* topk_weights[p] when
 *     given); x f32 [size_m, k] with size_m == P (one row per pair: the down projection) or size_m * topk == P (one row per token: gate / up).
 *     Activations are rounded to fp16 like QMatMul.  Q4_K / Q6_K with k % 256 == 0 (Q6_K: k % 2048 == 0) run grouped on the tcgen05
 *     dequant-GEMM -- every expert that was hit is streamed once -- given a 256-byte aligned workspace of moe_gemm_workspace_bytes();
 *     other shapes run a shape-generic kernel.  is_prefill is accepted for signature parity (the sort is the same here). */
void topk_softmax(const float* router_logits, float* topk_weights, uint32_t* topk_ids, int32_t num_tokens, int32_t num_experts, int32_t topk,
                  int64_t stream);
void sort_expert_assignments(const uint32_t* topk_ids, uint32_t* expert_ids, uint32_t* sorted_token_ids, int32_t num_pairs, int32_t num_experts,
                             int64_t stream);
size_t moe_gemm_workspace_bytes(int32_t num_pairs, int32_t n, int32_t k, int32_t num_experts);
void moe_gemm_gguf(const float* x, const void* experts, const float* topk_weights, const uint32_t* sorted_token_ids, const uint32_t* expert_ids,
                   float* out, int32_t num_experts, int32_t topk, int32_t size_m, int32_t num_pairs, int32_t n, int32_t k, int32_t ggml_type,
                   int32_t is_prefill, void* workspace, size_t workspace_bytes, int64_t stream);

/* moe_gemm_fp8 -- replaces attention_rs::moe::moe_gemm_fp8 (call sites /root/reference/src/openai/models/layers/moe.rs:1447-1473, block-FP8
 * experts as in DeepSeek-V3 / Qwen3 FP8 checkpoints): x [size_m, k] and out [num_pairs, n] of `dtype` (B200_F16 / B200_BF16); experts e4m3
 * [E, n, k]; scale f32 [E, ceil(n / block_y), ceil(k / block_x)] multiplies the weight tile; routing arguments as moe_gemm_gguf.  Weight-only:
 * activations are rounded to fp16, weights decoded exactly, fp32 accumulation, one rounding to `dtype`.  k % 256 == 0, n % block_y == 0,
 * block_x % 64 == 0 run grouped on the tcgen05 pipeline; other shapes on a shape-generic kernel.  workspace: moe_gemm_fp8_workspace_bytes(),
 * 256-byte aligned (required). */
size_t moe_gemm_fp8_workspace_bytes(int32_t num_pairs, int32_t n, int32_t k, int32_t num_experts);
void moe_gemm_fp8(const void* x, const void* experts, const float* scale, const float* topk_weights, const uint32_t* sorted_token_ids,
                  const uint32_t* expert_ids, void* out, int32_t num_experts, int32_t topk, int32_t size_m, int32_t num_pairs, int32_t n, int32_t k,
                  int32_t block_y, int32_t block_x, int32_t dtype, int32_t is_prefill, void* workspace, size_t workspace_bytes, int64_t stream);

/* ---- K15 / K21: the elementwise ops between the big ones ------------------------------------
 * rms_norm: candle_nn::ops::rms_norm (layers/qrmsnorm.rs:28-31).  out_dtype F32 or F16. */
void rms_norm(const float* x, const float* weight, void* out, int32_t rows, int32_t n, float eps,
              int32_t out_dtype, int64_t stream);
/* FusedRope::apply_inplace (layers/rotary_emb.rs:52-69): q [T,h,hd], k [T,kvh,hd] f32 in place,
 * cos/sin f32 [max_pos, hd/2], positions i64[T]; interleaved != 0 => rope_i (GGUF llama). */
void fused_rope_f32(float* q, float* k, const float* cos_t, const float* sin_t, const int64_t* positions,
                    int32_t num_tokens, int32_t num_heads, int32_t num_kv_heads, int32_t head_dim,
                    int32_t interleaved, int64_t stream);
/* silu(gate)*up (quantized_llama.rs:32-37); out_dtype F32 or F16 */
void silu_mul(const float* gate, const float* up, void* out, int64_t numel, int32_t out_dtype, int64_t stream);
void add_f32(float* x, const float* y, int64_t numel, int64_t stream);                 /* x += y */
void cast(const void* src, void* dst, int64_t numel, int32_t src_dtype, int32_t dst_dtype, int64_t stream);
void embedding_f32(const float* table, const int64_t* ids, float* out, int32_t num_tokens, int32_t dim, int64_t stream);
void argmax_f32(const float* logits, int32_t* out, int32_t rows, int32_t n, int64_t stream);

/* ---- fused decode layer pieces used by the engine (also exported for tests) ------------------
 * rope (interleaved or NeoX) on q,k + cast q -> dtype + reshape_and_cache of k,v, from the packed
 * f32 [T, (h + 2*kvh)*hd] output of the fused QKV projection. */
void rope_and_cache(const float* qkv, void* q_out, void* key_cache, void* value_cache,
                    const float* cos_t, const float* sin_t, const int64_t* positions,
                    const int64_t* slot_mapping, int32_t num_tokens, int32_t num_heads,
                    int32_t num_kv_heads, int32_t head_dim, int32_t block_size, int32_t interleaved,
                    int32_t dtype, int32_t cache_dtype, int64_t stream);

/* ---- decode engine: GGUFLLaMa::forward_inner for one decode step ------------------------------
 * Mirrors /root/reference/src/openai/models/quantized_llama.rs:424-506 (+ attention.rs:910-1011,
 * Mlp::forward :32-44) with the reference's CUDA-graph replay protocol (backend/graph.rs:685-803):
 * static device buffers, metadata copied in, one graph launch per step. */
typedef struct {
    int32_t hidden, num_layers, num_heads, num_kv_heads, head_dim, ffn, vocab;
    int32_t block_size, max_num_seqs, max_blocks_per_seq, max_pos;
    float   rms_eps, rope_theta;
    int32_t kv_dtype;          /* B200_BF16 or B200_FP8_E4M3 */
    int32_t tp_rank, tp_world; /* tensor-parallel shard (heads, kv heads, ffn, vocab split) */
    int32_t use_graph;
} b200_llama_config;

typedef struct {
    const float* attn_norm; const float* ffn_norm;          /* f32 [hidden] */
    const void *wq, *wk, *wv, *wo, *w1, *w2, *w3;           /* GGML blocks (device) */
    int32_t tq, tk, tv, to, t1, t2, t3;                     /* ggml types */
} b200_llama_layer;

/* A linear layer of the decode engine.  GGUF models use GGML tensors (b200_llama_layer above is the short form); the safetensors
 * models of the reference use Linear / GPTQ-Marlin / block-FP8 layers (/root/reference/src/openai/models/llama.rs:47-64,
 * linear.rs:124-172, :300-413): BASELINE configs 2 (dense BF16) and 3 (GPTQ/Marlin int4 + FP8 KV). */
enum { B200_LIN_GGML = 0,      /* w = GGML blocks, type = ggml type */
       B200_LIN_MARLIN4 = 1,   /* w = gptq_repack / awq_repack output, scales = marlin_permute_scales output (type = B200_F16 / B200_BF16 of the
                                  scales), zeros = NULL (symmetric GPTQ) or the converter's AWQ zero points, group_size 64 / 128 / -1 */
       B200_LIN_DENSE16 = 2 }; /* w = dense [n, k] of type B200_F16 / B200_BF16 */
typedef struct {
    int32_t kind, type;
    const void* w;
    const void* scales;
    const void* zeros;
    int32_t group_size, reserved;
} b200_linear;

typedef struct {
    const float* attn_norm; const float* ffn_norm;          /* f32 [hidden] */
    b200_linear wq, wk, wv, wo, w1, w2, w3;
} b200_llama_layer_ex;

typedef struct b200_llama b200_llama;

b200_llama* b200_llama_create(const b200_llama_config* cfg);
void        b200_llama_destroy(b200_llama* m);
/* all pointers are device pointers owned by the caller and must outlive the model */
void b200_llama_set_layer(b200_llama* m, int32_t layer, const b200_llama_layer* w);
void b200_llama_set_globals(b200_llama* m, const float* tok_embeddings /*f32 [vocab,hidden]*/,
                            const float* norm, const void* output_w, int32_t output_type);
/* general forms: any linear kind per weight; all linears of a model must share one activation format (GGML / MARLIN4: fp16 K4;
 * DENSE16: the weights' dtype).  rope_neox != 0 selects the NeoX rotation (i, i + hd/2) of the safetensors models (llama.rs:222)
 * instead of the interleaved one of GGUF (quantized_llama.rs:313-318).  The persistent layer kernel serves all-Q4_K layers only;
 * other kinds run one launch per GEMM. */
void b200_llama_set_layer_ex(b200_llama* m, int32_t layer, const b200_llama_layer_ex* w);
void b200_llama_set_globals_ex(b200_llama* m, const float* tok_embeddings, const float* norm, const b200_linear* output, int32_t rope_neox);
/* KV caches: flash layout, one K and one V device pointer per layer (cache_engine.rs:122-294) */
void b200_llama_set_kv_cache(b200_llama* m, void* const* key_caches, void* const* value_caches,
                             int64_t num_blocks);
/* tensor-parallel all-reduce hook: an NCCL communicator (ncclComm_t) created by the host */
void b200_llama_set_comm(b200_llama* m, void* nccl_comm);
/* Tensor parallel over NVLink peer memory: every rank allocates an inbox of b200_llama_peer_inbox_bytes() with
 * b200_ipc_alloc (cudaMalloc + zero + CUDA IPC handle), the handles travel over the host control plane, every rank maps its
 * peers' inboxes with b200_ipc_open and hands all `tp_world` pointers (its own at index tp_rank) to the engine.  The
 * row-parallel linears then run {GEMM, one fused kernel}: push the partial row to every peer, epoch flags, local sum, residual
 * add and the next RMSNorm (replaces AllReduce::cuda_fwd + add + RmsNorm, distributed.rs:572-653).  Without inboxes the engine
 * uses NCCL on the communicator of b200_llama_set_comm.  count = 0 switches back. */
size_t b200_llama_peer_inbox_bytes(const b200_llama* m);
void b200_llama_set_peer_inboxes(b200_llama* m, void* const* inboxes, int32_t count);
/* 1 when a row of the fused all-reduce gave up waiting for a peer (B200_TP_TIMEOUT_MS, default 120 s: a rank died).  The flag is a
 * host-mapped word: b200_llama_decode / read_next_tokens / read_logits check it after their stream sync and record an error
 * (b200_last_error) instead of handing back the poisoned step's tokens. */
int32_t b200_llama_peer_timeouts(b200_llama* m);
void* b200_ipc_alloc(size_t bytes, void* handle_out_64_bytes);
void* b200_ipc_open(const void* handle_64_bytes);
void b200_ipc_close(void* mapped);
void b200_ipc_free(void* allocated);
/* One decode step from HOST metadata (what prepare_decode builds, inputs.rs:376-454):
 * tokens u32[B], positions i64[B], slot_mapping i64[B], context_lens u32[B],
 * block_tables u32[B, table_width].  Copies them into the static device buffers (async, pinned
 * staging), replays the graph, and (if next_tokens_host != NULL) returns greedy argmax token ids
 * after synchronising the stream, like GraphCapturer::replay (graph.rs:297-301).
 * logits_host (optional): f32 [B, vocab] copied back -- the FULL vocabulary at any tp_world: vocab-parallel shards are all-gathered,
 * transposed and narrowed like VocabParallelLinear::forward (distributed.rs:1632-1667), so with tp_world > 1 asking for logits is a
 * collective (every rank asks in the same step).  Vocab-parallel lm_head: each rank holds pad_vocab_size(vocab, world) / world rows
 * (distributed.rs:1448-1454; rows past `vocab` are zero padding supplied by the caller and are never sampled).
 * context_lens[b] must fit the table (<= table_width * block_size) and slot_mapping[b] the cache; violations are argument errors. */
void b200_llama_decode(b200_llama* m, const uint32_t* tokens, const int64_t* positions,
                       const int64_t* slot_mapping, const uint32_t* context_lens,
                       const uint32_t* block_tables, int32_t table_width, int32_t num_seqs,
                       int32_t* next_tokens_host, float* logits_host, int64_t stream);
/* Device-resident variant: metadata already in the static buffers; advance positions/slots on the
 * device (fixed block tables) and replay.  Used for the HBM-resident `value` measurement. */
void b200_llama_decode_resident(b200_llama* m, int32_t num_seqs, int32_t advance, int64_t stream);
/* Measurement aid: one forward WITHOUT RoPE / cache write / attention, eagerly on `stream` -- every quantised projection of every
 * layer, the lm_head and the small ops between them (bench.py's roofline_gemm times the weight stream with it).  Leaves the
 * activations of the static buffers meaningless; never call it between decode steps whose results matter. */
void b200_llama_linear_chain(b200_llama* m, int32_t num_seqs, int64_t stream);
/* 1 when a step of `num_seqs` sequences runs on the persistent layer kernel (csrc/layer_mega.cu, opt-in with B200_MEGA=1: one launch per layer
 * for wo -> norm -> gate|up -> SiLU -> w2 -> norm -> next QKV, split-K sums reduced in a fixed order => bitwise reproducible logits, ~4 %
 * slower), 0 when it runs one launch per GEMM (the default: split-K sums meet in fp32 atomics; the run-to-run spread of the logits is bounded
 * and tested, tests/test_llama_gpu.py). */
int32_t b200_llama_uses_layer_kernel(b200_llama* m, int32_t num_seqs);
/* Profiling aid (env B200_MEGA_TRACE=<launch index> at model creation): clock64 stamps [CTA][phase 0..3][8] of that launch of the
 * persistent layer kernel, copied to `host`; returns the number of CTAs.  tools/mega_trace.py prints the timeline. */
int32_t b200_llama_mega_trace(b200_llama* m, long long* host, int32_t max_ctas);
const float* b200_llama_logits(b200_llama* m);        /* device f32 [max_num_seqs, vocab_local]: this rank's shard */
const int32_t* b200_llama_next_tokens(b200_llama* m); /* device i32 [max_num_seqs] */
int64_t b200_llama_kernel_launches(b200_llama* m);    /* kernels launched by this model so far */
/* copy the last step's greedy tokens (i32[n]) / full-vocabulary logits (f32 [n, vocab]; a collective when tp_world > 1) to the host;
 * syncs the stream */
void b200_llama_read_next_tokens(b200_llama* m, int32_t* host, int32_t n, int64_t stream);
void b200_llama_read_logits(b200_llama* m, float* host, int32_t n, int64_t stream);

/* library-wide count of kernels launched by this library (evidence for bench.py's gpu_launches) */
long long b200_total_kernel_launches(void);
/* NCCL all-reduce(sum) in place on f32, on the caller's communicator (tensor-parallel residual path,
 * /root/reference/src/openai/distributed.rs:547-654) */
void b200_allreduce_f32(void* nccl_comm, float* buf, int64_t n, int64_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_BACKEND_H_ */
