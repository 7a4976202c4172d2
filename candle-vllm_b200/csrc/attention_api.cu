// attention_api.cu -- C ABI of paged attention (K1 decode, K2 prefill) and dispatch between the
// TMA-staged split-KV decode kernel and the shape-generic kernel.
#include <cstdlib>

#include "attention.cuh"

using namespace b200;

extern "C" {

size_t paged_attention_decode_workspace_bytes(int32_t num_seqs, int32_t num_heads, int32_t head_dim,
                                              int32_t max_blocks_per_seq, int32_t block_size) {
    return paged_attention_decode_tma_workspace(num_seqs, num_heads, head_dim, max_blocks_per_seq, block_size);
}

void paged_attention_decode(void* out, const void* q, const void* key_cache, const void* value_cache,
                            const uint32_t* block_tables, const uint32_t* context_lens,
                            int32_t num_seqs, int32_t num_heads, int32_t num_kv_heads, int32_t head_dim,
                            int32_t block_size, int32_t max_blocks_per_seq, int64_t num_blocks,
                            float scale, float softcap, int32_t sliding_window,
                            int32_t dtype, int32_t cache_dtype, int32_t layout, int32_t out_dtype,
                            void* workspace, size_t workspace_bytes, int64_t stream) {
    if (num_seqs == 0) return;
    B200_REQUIRE(out && q && key_cache && value_cache && block_tables && context_lens, kErrBadArg, "paged_attention_decode: null pointer");
    B200_REQUIRE(num_seqs > 0 && num_heads > 0 && num_kv_heads > 0 && head_dim > 0 && block_size > 0 && max_blocks_per_seq > 0,
                 kErrBadArg, "paged_attention_decode: bad sizes");
    B200_REQUIRE(num_heads % num_kv_heads == 0, kErrBadArg, "paged_attention_decode: heads %d %% kv heads %d != 0", num_heads, num_kv_heads);
    B200_REQUIRE(head_dim <= 256, kErrUnsupported, "paged_attention_decode: head_dim %d > 256", head_dim);
    B200_REQUIRE(dtype == B200_BF16 || dtype == B200_F16, kErrUnsupported, "paged_attention_decode: dtype %d (bf16/f16 only)", dtype);
    B200_REQUIRE(out_dtype == dtype || out_dtype == B200_F16_K4 || (dtype == B200_BF16 && out_dtype == B200_F16), kErrUnsupported,
                 "paged_attention_decode: out dtype %d", out_dtype);
    B200_REQUIRE(out_dtype != B200_F16_K4 || head_dim % 4 == 0, kErrBadArg, "paged_attention_decode: K4 output needs head_dim %% 4 == 0");
    const bool fp8 = cache_dtype == B200_FP8_E4M3 || cache_dtype == B200_U8;
    B200_REQUIRE(fp8 || cache_dtype == dtype, kErrUnsupported, "paged_attention_decode: cache dtype %d vs %d", cache_dtype, dtype);
    B200_REQUIRE(layout == B200_KV_FLASH || layout == B200_KV_PAGED, kErrBadArg, "paged_attention_decode: layout %d", layout);
    cudaStream_t st = as_stream(stream);
    DecodeArgs d{out, q, key_cache, value_cache, block_tables, context_lens, num_seqs, num_heads, num_kv_heads,
                 head_dim, block_size, max_blocks_per_seq, num_blocks, scale, dtype, out_dtype, workspace, workspace_bytes};
    d.fp8 = fp8;
    if (paged_attention_decode_tma_supported(d, softcap, sliding_window, cache_dtype, layout)) {
        paged_attention_decode_tma(d, st);
        return;
    }
    GenericAttnArgs a{block_tables, context_lens, nullptr, nullptr, num_seqs, num_heads, num_kv_heads, head_dim,
                      block_size, max_blocks_per_seq, scale, softcap, sliding_window, layout, 0};
    paged_attention_generic(out, q, key_cache, value_cache, a, num_seqs, dtype, cache_dtype, out_dtype, st);
}

void paged_attention_prefill(void* out, const void* q, const void* key_cache, const void* value_cache,
                             const uint32_t* block_tables, const uint32_t* cu_seqlens_q,
                             const uint32_t* cu_seqlens_k, int32_t num_seqs, int32_t total_q,
                             int32_t max_seqlen_q, int32_t num_heads, int32_t num_kv_heads,
                             int32_t head_dim, int32_t block_size, int32_t max_blocks_per_seq,
                             float scale, float softcap, int32_t sliding_window,
                             int32_t dtype, int32_t cache_dtype, int32_t layout, int64_t stream) {
    (void)max_seqlen_q;
    if (num_seqs == 0 || total_q == 0) return;
    B200_REQUIRE(out && q && key_cache && value_cache && block_tables && cu_seqlens_q && cu_seqlens_k, kErrBadArg, "paged_attention_prefill: null pointer");
    B200_REQUIRE(num_heads % num_kv_heads == 0 && head_dim <= 256 && head_dim > 0, kErrBadArg, "paged_attention_prefill: bad head config");
    B200_REQUIRE(dtype == B200_BF16 || dtype == B200_F16, kErrUnsupported, "paged_attention_prefill: dtype %d", dtype);
    const bool fp8 = cache_dtype == B200_FP8_E4M3 || cache_dtype == B200_U8;
    B200_REQUIRE(fp8 || cache_dtype == dtype, kErrUnsupported, "paged_attention_prefill: cache dtype");
    static const bool force_generic = [] { const char* e = getenv("B200_PREFILL_GENERIC"); return e && atoi(e) != 0; }();
    if (!force_generic && paged_attention_prefill_tc_supported(head_dim, block_size, dtype, cache_dtype, layout, softcap, 1, q, key_cache, value_cache)) {
        // the ABI carries no block count (the reference's prefill entry has none either): the tensor map only needs an upper bound
        paged_attention_prefill_tc(out, q, key_cache, value_cache, block_tables, cu_seqlens_q, cu_seqlens_k, num_seqs, total_q, num_heads, num_kv_heads,
                                   max_blocks_per_seq, (int64_t)1 << 22, scale, sliding_window, dtype, as_stream(stream));
        return;
    }
    GenericAttnArgs a{block_tables, nullptr, cu_seqlens_q, cu_seqlens_k, num_seqs, num_heads, num_kv_heads, head_dim,
                      block_size, max_blocks_per_seq, scale, softcap, sliding_window, layout, 1};
    paged_attention_generic(out, q, key_cache, value_cache, a, total_q, dtype, cache_dtype, dtype, as_stream(stream));
}

}  // extern "C"
