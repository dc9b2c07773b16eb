// Shared helpers for the rf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rf_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "rf_b200 is written for sm_100a (B200) only"
#endif

void rf_set_error(const char* fmt, ...);

#define RF_CHECK_ARG(cond, ...)      \
    do {                             \
        if (!(cond)) {               \
            rf_set_error(__VA_ARGS__); \
            return 1;                \
        }                            \
    } while (0)

#define RF_CUDA_OK(call)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            rf_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

#define RF_LAUNCH_OK(name)                                                               \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            rf_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));      \
            return 3;                                                                    \
        }                                                                                \
    } while (0)

static inline int rf_cdiv(long a, long b) { return (int)((a + b - 1) / b); }
static inline long rf_cdivl(long a, long b) { return (a + b - 1) / b; }

// Grid for an elementwise kernel using a grid-stride loop: enough CTAs to fill
// 148 SMs a few times over, never more than the work needs.
static inline int rf_grid_1d(long n, int block, int max_ctas = 148 * 16) {
    long g = (n + block - 1) / block;
    if (g > max_ctas) g = max_ctas;
    if (g < 1) g = 1;
    return (int)g;
}

__device__ __forceinline__ float rf_act(float v, int act, float slope) {
    if (act == RF_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == RF_ACT_LEAKY) return v > 0.f ? v : v * slope;
    if (act == RF_ACT_TANH) return tanhf(v);
    return v;
}

// Activation of a register vector with ONE switch on the (runtime) activation kind: rf_act per element makes
// the compiler expand the tanh path next to every element and branch around it 16 times per group, which
// turned the tensor-core epilogues into ~900 instructions per 16 columns.
template <int N>
__device__ __forceinline__ void rf_act_vec(float (&v)[N], int act, float slope) {
    if (act == RF_ACT_RELU) {
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] = fmaxf(v[e], 0.f);
    } else if (act == RF_ACT_LEAKY) {
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * slope;
    } else if (act == RF_ACT_TANH) {
#pragma unroll  // (a rolled loop would index v dynamically and push the whole vector to local memory)
        for (int e = 0; e < N; ++e) v[e] = tanhf(v[e]);
    }
}
