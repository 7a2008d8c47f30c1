// Shared helpers for the MRSSM sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mrssm_b200.h"

void mrssm_set_error(const char* fmt, ...);

#define MRSSM_CHECK(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            mrssm_set_error(__VA_ARGS__);      \
            return 1;                          \
        }                                      \
    } while (0)

#define MRSSM_CUDA(expr)                                                                    \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            mrssm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

#define MRSSM_LAUNCH_CHECK() MRSSM_CUDA(cudaGetLastError())

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == MRSSM_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == MRSSM_ACT_ELU) return v > 0.f ? v : expm1f(v);
    return v;
}
// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
    if (act == MRSSM_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (act == MRSSM_ACT_ELU) return y > 0.f ? 1.f : y + 1.f;
    return 1.f;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// torch F.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
