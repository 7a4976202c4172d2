// attention_decode.cu -- placeholder until the TMA-staged split-KV kernel lands (next commit).
#include "attention.cuh"
namespace b200 {
bool paged_attention_decode_tma_supported(const DecodeArgs&, float, int, int, int) { return false; }
size_t paged_attention_decode_tma_workspace(int, int, int, int, int) { return 256; }
void paged_attention_decode_tma(const DecodeArgs&, cudaStream_t) {}
}  // namespace b200
