// launch.cuh -- one launch helper for every kernel of the encoder chain: adds the programmatic-dependent-launch attribute so
// kernel i+1's prologue (and its launch latency) overlaps kernel i's tail, in eager streams and inside the captured CUDA graph
// alike. Device side: ptx.cuh pdl_wait() / pdl_launch_dependents(). Q2W_PDL=0 turns the attribute off (plain stream order).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace q2w {

// Function attributes (the > 48 KB dynamic shared memory opt-in) and the SM count are PER DEVICE: a process may hold contexts on
// several GPUs (q2w_model_create takes any ordinal, q2w_multi_* drives all of them), so every latch below is keyed by the device
// that is current at launch time, never by "first call in this process".
constexpr int kMaxDevices = 64;

struct DeviceInfo {
    int dev = -1;
    int num_sms = 0;
};
inline cudaError_t current_device_info(DeviceInfo& out) {
    static std::atomic<int> sms[kMaxDevices] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    int n = sms[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        sms[dev].store(n, std::memory_order_relaxed);
    }
    out.dev = dev;
    out.num_sms = n;
    return cudaSuccess;
}

// one opt-in per (kernel, device); `done` is a per-kernel bit mask owned by the caller (a function-local static)
template <typename K>
inline cudaError_t smem_optin_once(K kernel, int bytes, int dev, std::atomic<unsigned long long>& done) {
    const unsigned long long bit = 1ull << dev;
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("Q2W_PDL");
        return !(e && std::strcmp(e, "0") == 0);
    }();
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace q2w
