// tp.cu -- tensor-parallel collectives on the residual path (AllReduce after the row-parallel
// o_proj / down_proj: /root/reference/src/openai/distributed.rs:547-654, :696-710).
// NCCL is resolved at run time from the already-loaded torch-bundled libnccl.so.2 (the host
// process creates the communicator); no link-time dependency.
#include <dlfcn.h>

#include "common.cuh"

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

}  // namespace b200

extern "C" void b200_allreduce_f32(void* nccl_comm, float* buf, int64_t n, int64_t stream) {
    b200::tp_allreduce_f32(nccl_comm, buf, n, b200::as_stream(stream));
}
