// lib.cu -- library state: per-thread error record, launch counter, device probing.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace b200 {

static thread_local int t_err = 0;
static thread_local char t_msg[512] = {0};
static std::atomic<long long> g_launches{0};

void set_error(int code, const char* fmt, ...) {
    t_err = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_msg, sizeof(t_msg), fmt, ap);
    va_end(ap);
}

bool check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error(kErrCuda, "%s: %s", what, cudaGetErrorString(e));
        return false;
    }
    return true;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
    static const int on = [] { const char* e = getenv("B200_PDL"); return e ? atoi(e) : 1; }();
    return on != 0;
}

constexpr int kMaxDevices = 64;

int sm_count() {
    static std::atomic<int> cached[kMaxDevices];          // zero-initialised: 0 = not probed yet
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return 148; }
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            return 148;   // B200; only used to size grids
        }
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

void ensure_dynamic_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> done;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
    std::lock_guard<std::mutex> lk(mu);
    if (done.count({dev, kernel})) return;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess) done.insert({dev, kernel});
    else cudaGetLastError();
}

}  // namespace b200

extern "C" {

int b200_abi_version(void) { return 1; }

int b200_last_error(void) {
    int e = b200::t_err;
    b200::t_err = 0;
    return e;
}

const char* b200_last_error_message(void) { return b200::t_msg; }

int b200_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b200_device_cc(void) {
    int dev = 0, ma = 0, mi = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return ma * 10 + mi;
}

long long b200_total_kernel_launches(void) { return b200::g_launches.load(); }

}  // extern "C"
