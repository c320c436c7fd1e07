// Programmatic dependent launch (PDL) for the model-stage kernel chain: a kernel launched through launchPdl may become resident
// while its predecessor in the stream drains (the predecessor calls pdlLaunchDependents() at its start), runs its prologue
// (barrier init, TMEM allocation, constant loads) and then blocks in pdlWait() until the predecessor has completed and its
// writes are visible.  Rule for every kernel launched this way: pdlWait() is executed by every thread before the first access
// to memory that any earlier kernel writes or still reads.  W2X_NO_PDL=1 falls back to plain stream order (W2X_DEV build only).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

#include "../hostutil.h"

namespace w2x {

#ifdef __CUDACC__
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdlLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

inline bool pdlEnabled() {
    static const bool on = devEnv("W2X_NO_PDL") == nullptr;
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launchPdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdlEnabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace w2x
