// Shared helpers for libbgnn.so (sm_100a).  Not part of the public ABI (see include/bgnn.h).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "bgnn.h"

void bgnn_set_error(const char* fmt, ...);
void bgnn_prof_start(int slot, cudaStream_t s);
void bgnn_prof_stop(cudaStream_t s);
struct ProfScope {          // records an event pair around the kernel launches in its scope (bench aid)
    cudaStream_t s;
    ProfScope(int slot, cudaStream_t st) : s(st) { bgnn_prof_start(slot, st); }
    ~ProfScope() { bgnn_prof_stop(s); }
};

#define BGNN_CHECK_ARG(cond, ...)                      \
    do {                                               \
        if (!(cond)) {                                 \
            bgnn_set_error(__VA_ARGS__);               \
            return BGNN_E_INVALID;                     \
        }                                              \
    } while (0)

#define BGNN_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            bgnn_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                           cudaGetErrorString(e__));                                           \
            return BGNN_E_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define BGNN_LAUNCH_CHECK()  BGNN_CUDA(cudaGetLastError())

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t  align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// numpy.nan_to_num(x, nan=0.0) on float32: NaN -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX
__device__ __forceinline__ float nan_to_num_f32(float v) {
    if (isnan(v)) return 0.0f;
    if (isinf(v)) return v > 0 ? 3.402823466e+38f : -3.402823466e+38f;
    return v;
}
