// attention.cuh -- internal interfaces between the attention translation units.
#pragma once
#include "common.cuh"

namespace b200 {

struct GenericAttnArgs {
    const uint32_t* block_tables;
    const uint32_t* context_lens;   // decode
    const uint32_t* cu_q;           // prefill
    const uint32_t* cu_k;
    int num_seqs, num_heads, num_kv_heads, head_dim, block_size, max_blocks;
    float scale, softcap;
    int window, layout, prefill;
};

// catch-all path (attention_generic.cu)
void paged_attention_generic(void* out, const void* q, const void* kc, const void* vc, const GenericAttnArgs& a,
                             int rows, int dtype, int cache_dtype, int out_dtype, cudaStream_t st);

// TMA-staged split-KV decode (attention_decode.cu); returns false if the shape is not covered
struct DecodeArgs {
    void* out; const void* q; const void* kc; const void* vc;
    const uint32_t* block_tables; const uint32_t* context_lens;
    int num_seqs, num_heads, num_kv_heads, head_dim, block_size, max_blocks;
    int64_t num_blocks;
    float scale;
    int dtype, out_dtype;
    void* workspace; size_t workspace_bytes;
    bool fp8 = false;                 // the cache holds e4m3 bytes (scale 1.0)
};
bool paged_attention_decode_tma_supported(const DecodeArgs& a, float softcap, int window, int cache_dtype, int layout);
size_t paged_attention_decode_tma_workspace(int num_seqs, int num_heads, int head_dim, int max_blocks, int block_size);
void paged_attention_decode_tma(const DecodeArgs& a, cudaStream_t st);

// tensor-core (chunked) prefill over the paged cache (attention_prefill.cu): flash layout, head_dim 128, block 64, 16-bit cache
bool paged_attention_prefill_tc_supported(int head_dim, int block_size, int dtype, int cache_dtype, int layout, float softcap, int64_t num_blocks,
                                          const void* q, const void* kc, const void* vc);
void paged_attention_prefill_tc(void* out, const void* q, const void* kc, const void* vc, const uint32_t* block_tables, const uint32_t* cu_q,
                                const uint32_t* cu_k, int num_seqs, int total_q, int num_heads, int num_kv_heads, int max_blocks, int64_t num_blocks,
                                float scale, int window, int dtype, cudaStream_t st);

}  // namespace b200
