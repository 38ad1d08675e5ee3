// Shared helpers for the r2d2_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/r2d2_b200.h"

namespace r2d2 {

void set_error(const char* fmt, ...);

#define R2D2_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::r2d2::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                              cudaGetErrorString(_e));                                     \
            return R2D2_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define R2D2_REQUIRE(cond, msg)                                                            \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            ::r2d2::set_error("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond);           \
            return R2D2_ERR_ARG;                                                           \
        }                                                                                  \
    } while (0)

#define R2D2_LAUNCH_CHECK() R2D2_CUDA_CHECK(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember per device (bit = ordinal) that a kernel
// has been opted in, so a process that drives several GPUs configures each of them once.
template <class Kernel>
static inline cudaError_t ensure_dynamic_smem(Kernel kern, int bytes, unsigned long long* configured_devices) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (*configured_devices & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) *configured_devices |= bit;
    return e;
}

}  // namespace r2d2
