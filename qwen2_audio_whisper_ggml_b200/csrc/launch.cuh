// launch.cuh -- one launch helper for every kernel of the encoder chain: adds the programmatic-dependent-launch attribute so
// kernel i+1's prologue (and its launch latency) overlaps kernel i's tail, in eager streams and inside the captured CUDA graph
// alike. Device side: ptx.cuh pdl_wait() / pdl_launch_dependents(). Q2W_PDL=0 turns the attribute off (plain stream order).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

namespace q2w {

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
