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

// One-time, PER-DEVICE opt-in to more than 48 KB of dynamic shared memory for `kernel` (function attributes belong to
// a device's context; a process may drive several devices).  rf_create(device) runs every kernel's opt-in eagerly;
// the launch sites repeat it lazily, so the library also works without a handle.  The flags are idempotent caches,
// not state: a race between threads merely sets the attribute twice.
#define RF_SMEM_OPT_IN(kernel, bytes)                                                                        \
    do {                                                                                                     \
        static bool done__[64] = {};                                                                         \
        int dev__ = 0;                                                                                       \
        RF_CUDA_OK(cudaGetDevice(&dev__));                                                                   \
        if (dev__ < 0 || dev__ >= 64 || !done__[dev__]) {                                                    \
            RF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (dev__ >= 0 && dev__ < 64) done__[dev__] = true;                                              \
        }                                                                                                    \
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

// Correctly rounded a / b for a divisor that is fixed per launch (the (x - mean) / std normalisation must be the
// same two rounded fp32 operations numpy performs).  __fdiv_rn costs ~20 instructions per element and made the
// re-indexing kernels issue-bound; with y = RN(1 / b) from the host, two Newton steps on the quotient give a
// faithful q1 and Markstein's theorem (q faithful, r = a - b q exact by FMA, y = RN(1/b)  =>  RN(q + r y) = RN(a / b))
// makes the third rounding exact: 5 FMA-class instructions.  Guarded ranges keep every intermediate normal
// (|a| in [2^-20, 2^20], |b| in [2^-10, 2^10], checked on the host: y == 0 means "use __fdiv_rn"); zeros, tiny and
// huge values take the IEEE division.
static inline float rf_host_rcp_for_div(float b) {
    const float ab = b < 0.f ? -b : b;
    return (ab >= 0.0009765625f && ab <= 1024.f) ? 1.0f / b : 0.f;
}
__device__ __forceinline__ float rf_div_rn_fixed(float a, float b, float y) {
    const float aa = fabsf(a);
    if (y != 0.f && aa >= 9.5367431640625e-07f && aa <= 1048576.f) {
        const float q0 = __fmul_rn(a, y);
        const float r0 = __fmaf_rn(-b, q0, a);
        const float q1 = __fmaf_rn(r0, y, q0);
        const float r1 = __fmaf_rn(-b, q1, a);
        return __fmaf_rn(r1, y, q1);
    }
    return __fdiv_rn(a, b);
}

// Exact division of a 32-bit index by a run-time divisor with one 64-bit high multiply: m = ceil(2^64 / d) gives
// floor(n / d) for every n < 2^32 (error term n * (m * d - 2^64) / (d * 2^64) < 1 / d).  The re-indexing kernels
// decompose one linear index per 16-byte access, so the ~20-instruction hardware-less integer division matters.
struct FastDiv {
    unsigned long long m;
    unsigned d;
};
static inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = (unsigned)d;
    f.m = d <= 1 ? 0ull : (~0ull) / (unsigned long long)d + 1ull;
    return f;
}
__device__ __forceinline__ unsigned fd_div(unsigned n, const FastDiv& f) {
    return f.d == 1u ? n : (unsigned)__umul64hi((unsigned long long)n, f.m);
}
// n -> n / d, returns n % d
__device__ __forceinline__ unsigned fd_divmod(unsigned& n, const FastDiv& f) {
    const unsigned q = fd_div(n, f);
    const unsigned r = n - q * f.d;
    n = q;
    return r;
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
        if (slope > 0.f && slope <= 1.f) {
            // max(v, v * slope): the same value for every input (v * slope <= v iff v >= 0 when 0 < slope <= 1, rounding is
            // monotonic, signed zeros, infinities and NaN included), one instruction fewer per element than compare + select
#pragma unroll
            for (int e = 0; e < N; ++e) v[e] = fmaxf(v[e], v[e] * slope);
        } else {
#pragma unroll
            for (int e = 0; e < N; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * slope;
        }
    } else if (act == RF_ACT_TANH) {
#pragma unroll  // (a rolled loop would index v dynamically and push the whole vector to local memory)
        for (int e = 0; e < N; ++e) v[e] = tanhf(v[e]);
    }
}
