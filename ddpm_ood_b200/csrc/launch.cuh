// Kernel launch with programmatic dependent launch (PDL) enabled: the launched grid may begin while the previous grid
// in the stream is still draining; every such kernel calls ptx::pdl_wait() before it touches global memory.
// DDPM_PDL=0 in the environment turns the attribute off (plain stream order) for A/B measurements.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

namespace ddpm {

// Function attributes (cudaFuncSetAttribute) and the SM count are PER DEVICE: one process may drive several GPUs (the
// module re-creates its engine when moved), so "already set" state is kept per device ordinal.
constexpr int kMaxDevices = 64;
inline int device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev % kMaxDevices;
}

inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("DDPM_PDL");
        on = (e && atoi(e) == 0) ? 0 : 1;
    }
    return on != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace ddpm
