// cudaLaunchKernelEx wrapper shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>

namespace amc {

// SM count of the current device (cached per device): grid caps are multiples of it, not of a hard-coded 148
inline int device_sm_count() {
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, int block, size_t smem, cudaStream_t s, bool pdl,
                             Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace amc
