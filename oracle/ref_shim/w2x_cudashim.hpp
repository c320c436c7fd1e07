// TEST INFRASTRUCTURE ONLY -- part of oracle/_ref.  Host-memory stand-in for the handful of CUDA runtime calls the
// reference's src/tensorrt/*.cpp makes (helper.h:13-56, img2img_load.cpp:137-243, img2img_infer.cpp:73-76,
// img2img_base.cpp:6-10): "device" memory is malloc'd host memory, streams are synchronous, there is one fake device whose
// name the driver sets (it feeds getConfigHash and the json sidecar).
#pragma once
#include <cstdlib>
#include <cstring>
#include <string>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidDevice = 101, cudaErrorMemoryAllocation = 2 };
typedef struct w2xshimStream* cudaStream_t;
enum { cudaStreamDefault = 0, cudaStreamNonBlocking = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
struct cudaDeviceProp {
    char name[256];
};

namespace w2xshim {
inline std::string& deviceName() {
    static std::string n = "NVIDIA B200";
    return n;
}
}  // namespace w2xshim

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorInvalidDevice ? "invalid device ordinal" : "cuda shim error"; }
inline cudaError_t cudaSetDevice(int id) { return id == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int id) {
    std::memset(p, 0, sizeof(*p));
    if (id != 0) return cudaErrorInvalidDevice;
    std::strncpy(p->name, w2xshim::deviceName().c_str(), sizeof(p->name) - 1);
    return cudaSuccess;
}
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T>
inline cudaError_t cudaMallocAsync(T** p, size_t n, cudaStream_t) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(dst, src, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
