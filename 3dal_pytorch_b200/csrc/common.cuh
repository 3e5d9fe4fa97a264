// Shared helpers for libal3d.so (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstdint>

namespace al3d {

char *error_buffer();                       // thread-local, 512 bytes (defined in capi.cu)
int set_error(const char *fmt, ...);        // formats into error_buffer(), returns 1

#define AL3D_CHECK_ARG(cond, ...)                                   \
    do { if (!(cond)) return ::al3d::set_error(__VA_ARGS__); } while (0)

#define AL3D_CHECK_LAUNCH(name)                                                         \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return ::al3d::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

#define AL3D_CHECK_CUDA(call)                                                            \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return ::al3d::set_error("%s failed: %s", #call, cudaGetErrorString(e__));   \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace al3d
