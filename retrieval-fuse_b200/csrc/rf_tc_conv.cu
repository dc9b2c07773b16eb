// Tensor-core 3D convolution for the U-Nets (model/unet.py 'gcr' blocks) and
// the conv patch encoders, on channels-last activations.
//
// Pipeline of one SingleConv (GroupNorm -> Conv3d -> ReLU):
//   cl_gn_partial_*       per (sample, channel) fp64 sums of the fp32 channels-last input
//   + cl_gn_finalize      (the skip concat and the nearest 2x upsampling are virtual: two
//                         source tensors) -> per (sample, group) mean / rstd.  One warp per
//                         sample for many small samples, a cp.async.bulk ring through shared
//                         memory otherwise, scalar loads as the fallback (rf_cl_gn_stats)
//   cl_norm_split_kernel  y = (x-mu)*a+beta, split into fp16 hi/lo, channels
//                         padded to a multiple of 8 -> [N,D,H,W,Cp] x 2.  The
//                         normalisation runs ONCE per element here instead of
//                         27x inside the convolution's operand gather.
//   tc_conv3d_kernel      implicit GEMM on tcgen05: M = output voxels (128 per
//                         CTA), N = Cout (padded to 16), K = taps x padded
//                         channels in blocks of 64.  With channels last, one
//                         16-byte chunk of the A operand (8 fp16 channels of one
//                         input voxel and tap) is one aligned 16-byte global
//                         load; 8 lanes fetch the 8 chunks of a row, zero
//                         padding is a predicate.  fp16 hi/lo split, three
//                         products per K step, fp32 accumulators in TMEM.
//                         Epilogue: bias + activation, fp32 channels-last
//                         (next layer) or NCDHW (module boundary) stores.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "rf_common.cuh"

namespace {

constexpr int TM = 128, IMG = TM * 128, NTHREADS = 256;
constexpr int MAX_A_STAGES = 2, MAX_B_STAGES = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 26)) {  // a pipeline bug must never hang the GPU
            printf("rf_tc_conv: mbarrier wait timed out (block %d thread %d bar %u)\n", blockIdx.x, threadIdx.x, bar);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
          "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
          "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {  // K-major SWIZZLE_128B (see rf_knn_tc.cu)
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_f16(int n) {  // D f32 (bit 4), A/B f16 (format 0), K-major, N>>3 @17, M>>4 @24
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
// x = hi + lo + r with hi, lo fp16: |r| <= 2^-24 |x| for |x| in fp16's normal range (the
// operands here are GroupNorm-ed activations, TSDF patches and weights, all O(1)); hi is
// saturated so that even |x| up to 1.3e5 splits without producing inf.
__device__ __forceinline__ void split_f16(float x, uint32_t& hi, uint32_t& lo) {
    const __half h = __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}

// ----------------------------------------------------------------- GroupNorm statistics (channels last)
// x [N,D,H,W,C1] (+ optional x2 [N,D/2,H/2,W/2,C2], virtually upsampled and concatenated after x).
// Pass 1: every CTA reduces a slice of voxels for ALL channels (coalesced rows of C floats; x2 is
// read at its own resolution and weighted by 8, since each coarse voxel appears 8 times in the
// upsampled volume) into per-(sample, channel) fp64 sum / sum of squares (one atomicAdd pair per
// channel per CTA).  Pass 2 folds channels into groups: mean, 1/sqrt(var+eps), expanded per channel.
// fp64 moments of fp32 data: E[x^2]-E[x]^2 loses nothing that matters unless var/mean^2 < 1e-10.
__global__ void __launch_bounds__(256) cl_gn_partial_kernel(const float* __restrict__ x, long S, int C, int slices, double weight,
                                                            int c_off, int c_tot, double* __restrict__ sums) {
    const int n = blockIdx.x / slices, sl = blockIdx.x % slices;
    const long per = (S + slices - 1) / slices;
    const long v0 = sl * per, v1 = v0 + per < S ? v0 + per : S;
    // thread -> channel (tid % C) when C <= 256, rows strided by 256 / C; else loop over channels
    extern __shared__ double sh[];  // [2][C]
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sh[c] = 0.0;
    __syncthreads();
    const long total = (v1 - v0) * C;
    const float* base = x + ((long)n * S + v0) * C;
    if (C <= 256) {  // the first T = floor(256/C)*C threads each keep ONE channel: no per-element modulo, coalesced rows
        const int T = (256 / C) * C;
        if ((int)threadIdx.x < T) {
            const int c = threadIdx.x % C;
            double s1 = 0.0, s2 = 0.0;
            long e = threadIdx.x;
            // four independent loads in flight per thread (one load per trip left the kernel waiting on DRAM latency:
            // 1.5 TB/s), two accumulator pairs to shorten the fp64 dependency chains
            double t1 = 0.0, t2 = 0.0;
            for (; e + 3L * T < total; e += 4L * T) {
                const float f0 = __ldg(base + e), f1 = __ldg(base + e + T), f2 = __ldg(base + e + 2L * T), f3 = __ldg(base + e + 3L * T);
                const double v0 = (double)f0, v1 = (double)f1, v2 = (double)f2, v3 = (double)f3;
                s1 += v0; s2 += v0 * v0;
                t1 += v1; t2 += v1 * v1;
                s1 += v2; s2 += v2 * v2;
                t1 += v3; t2 += v3 * v3;
            }
            for (; e < total; e += T) {
                const double v = (double)__ldg(base + e);
                s1 += v; s2 += v * v;
            }
            s1 += t1; s2 += t2;
            if (C <= 32 && (32 % C) == 0) {
                // lanes l, l + C, l + 2C, ... of a warp hold the same channel (T is a multiple of 32 here): fold them
                // with shuffles so that only C lanes per warp touch the shared accumulators (a 1-channel tensor made
                // all 256 threads hit one shared-memory word)
                for (int o = 16; o >= C; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if ((int)(threadIdx.x & 31) < C) {
                    atomicAdd(&sh[c], s1);
                    atomicAdd(&sh[C + c], s2);
                }
            } else {
                atomicAdd(&sh[c], s1);
                atomicAdd(&sh[C + c], s2);
            }
        }
    } else {
        for (long e = threadIdx.x; e < total; e += 256) {
            const double v = (double)__ldg(base + e);
            const int c = (int)(e % C);
            atomicAdd(&sh[c], v);
            atomicAdd(&sh[C + c], v * v);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(&sums[((long)n * c_tot + c_off + c) * 2], sh[c] * weight);
        atomicAdd(&sums[((long)n * c_tot + c_off + c) * 2 + 1], sh[C + c] * weight);
    }
}

// Same reduction for C % 4 == 0 (every GroupNorm of the U-Nets): a thread keeps one channel QUAD and reads float4s, four
// of them in flight.  The scalar kernel above has 2048 threads x 4 loads x 4 B = 32 KiB in flight per SM, less than the
// ~52 KiB HBM's latency-bandwidth product asks of an SM (it measured 2.5 - 3.3 TB/s on the 0.5 - 1 GB tensors of a
// 64-chunk step); 16-byte loads put 4x the bytes behind the same number of requests.
__global__ void __launch_bounds__(256) cl_gn_partial_vec4_kernel(const float* __restrict__ x, long S, int C, int slices, double weight,
                                                                 int c_off, int c_tot, double* __restrict__ sums) {
    const int n = blockIdx.x / slices, sl = blockIdx.x % slices;
    const long per = (S + slices - 1) / slices;
    const long v0 = sl * per, v1 = v0 + per < S ? v0 + per : S;
    extern __shared__ double sh[];  // [2][C]
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sh[c] = 0.0;
    __syncthreads();
    const int C4 = C >> 2;
    const int T = (256 / C4) * C4;  // active threads: whole rows of channel quads
    if ((int)threadIdx.x < T && v1 > v0) {
        const long total = (v1 - v0) * C4;
        const float4* base = reinterpret_cast<const float4*>(x + ((long)n * S + v0) * C);
        const int q = threadIdx.x % C4;
        double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
        long e = threadIdx.x;
        auto acc = [&](const float4& f) {
            const double a0 = (double)f.x, a1 = (double)f.y, a2 = (double)f.z, a3 = (double)f.w;
            s1[0] += a0; s2[0] += a0 * a0;
            s1[1] += a1; s2[1] += a1 * a1;
            s1[2] += a2; s2[2] += a2 * a2;
            s1[3] += a3; s2[3] += a3 * a3;
        };
        for (; e + 3L * T < total; e += 4L * T) {
            float4 f0 = __ldg(base + e);
            const float4 f1 = __ldg(base + e + T), f2 = __ldg(base + e + 2L * T), f3 = __ldg(base + e + 3L * T);
            // ptxas sinks each load to right above its first use (one load in flight per thread) to save registers; a
            // value-preserving dependency of the first use on ALL four loads keeps them in flight together
            f0.x = fmaf(0.f, f1.w, fmaf(0.f, f2.w, fmaf(0.f, f3.w, f0.x)));
            acc(f0); acc(f1); acc(f2); acc(f3);
        }
        for (; e < total; e += T) acc(__ldg(base + e));
        if (C4 <= 32 && (32 % C4) == 0) {
            // T is a multiple of 32 here: lanes l, l + C4, ... of a warp hold the same quad - fold them with shuffles
            for (int o = 16; o >= C4; o >>= 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                    s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
                }
            }
            if ((int)(threadIdx.x & 31) < C4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    atomicAdd(&sh[4 * q + j], s1[j]);
                    atomicAdd(&sh[C + 4 * q + j], s2[j]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&sh[4 * q + j], s1[j]);
                atomicAdd(&sh[C + 4 * q + j], s2[j]);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(&sums[((long)n * c_tot + c_off + c) * 2], sh[c] * weight);
        atomicAdd(&sums[((long)n * c_tot + c_off + c) * 2 + 1], sh[C + c] * weight);
    }
}

// The same reduction fed by bulk copies (cp.async.bulk global -> shared, mbarrier completion): the CTA's slice streams
// through a two-stage ring of 16 KiB chunks that one thread keeps full, the 256 threads reduce each chunk from shared
// memory (conflict-free 16-byte reads).  No load sits in a register while it is in flight, so neither the register
// allocator nor the instruction scheduler decides how many bytes an SM has outstanding: up to six CTAs x 32 KiB.
constexpr int GN_CHUNK4 = 1024, GN_STAGES = 2;  // float4s per stage
__global__ void __launch_bounds__(256) cl_gn_partial_bulk_kernel(const float* __restrict__ x, long S, int C, int slices, double weight,
                                                                 int c_off, int c_tot, double* __restrict__ sums) {
    extern __shared__ __align__(128) uint8_t gn_smem[];
    const float4* buf = reinterpret_cast<const float4*>(gn_smem);                              // [GN_STAGES][GN_CHUNK4]
    const uint32_t bar0 = smem_u32(gn_smem + (size_t)GN_STAGES * GN_CHUNK4 * 16);
    const uint32_t sbuf = smem_u32(gn_smem);
    const int n = blockIdx.x / slices, sl = blockIdx.x % slices;
    const long per = (S + slices - 1) / slices;
    const long v0 = sl * per, v1 = v0 + per < S ? v0 + per : S;
    const int C4 = C >> 2;
    const int T = (256 / C4) * C4;               // active threads: whole rows of channel quads
    const int chunk4 = (GN_CHUNK4 / C4) * C4;    // chunks start on a row boundary: a thread keeps ONE quad
    const long total4 = v1 > v0 ? (v1 - v0) * C4 : 0;
    const int n_chunks = (int)((total4 + chunk4 - 1) / chunk4);
    const float4* base = reinterpret_cast<const float4*>(x + ((long)n * S + v0) * C);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int st = 0; st < GN_STAGES; ++st) mbar_init(bar0 + 8 * st, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int i) {
        const int st = i % GN_STAGES;
        const long left = total4 - (long)i * chunk4;
        const uint32_t bytes = (uint32_t)(left < chunk4 ? left : chunk4) * 16u;
        mbar_arrive_expect_tx(bar0 + 8 * st, bytes);
        bulk_g2s(sbuf + (uint32_t)st * GN_CHUNK4 * 16u, base + (long)i * chunk4, bytes, bar0 + 8 * st);
    };
    if (tid == 0)
        for (int i = 0; i < GN_STAGES && i < n_chunks; ++i) issue(i);
    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < n_chunks; ++i) {
        const int st = i % GN_STAGES;
        mbar_wait(bar0 + 8 * st, (uint32_t)(i / GN_STAGES) & 1u);
        const long left = total4 - (long)i * chunk4;
        const int n4 = (int)(left < chunk4 ? left : chunk4);
        const float4* b = buf + st * GN_CHUNK4;
        if (tid < T) {
#pragma unroll 4
            for (int f = tid; f < n4; f += T) {
                const float4 v = b[f];
                const double a0 = (double)v.x, a1 = (double)v.y, a2 = (double)v.z, a3 = (double)v.w;
                s1[0] += a0; s2[0] += a0 * a0;
                s1[1] += a1; s2[1] += a1 * a1;
                s1[2] += a2; s2[2] += a2 * a2;
                s1[3] += a3; s2[3] += a3 * a3;
            }
        }
        __syncthreads();  // the stage is consumed: refill it
        if (tid == 0 && i + GN_STAGES < n_chunks) issue(i + GN_STAGES);
    }
    // Per-thread sums -> per-channel sums without shared-memory atomics (fp64 atomicAdd on shared memory is a CAS loop):
    // "owners" park their 8 sums in slots of the (consumed) chunk ring, slot % C4 == the owner's quad; thread t < 2C
    // then adds up the slots of its channel.
    double* part = reinterpret_cast<double*>(gn_smem);  // [owners][8]
    int n_owners = 0;
    if (n_chunks > 0) {
        const bool fold = C4 <= 32 && (32 % C4) == 0;   // T = 256: lanes l, l + C4, ... of a warp hold the same quad
        if (fold) {
            for (int o = 16; o >= C4; o >>= 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                    s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
                }
            }
        }
        n_owners = fold ? 8 * C4 : T;
        const int slot = fold ? ((tid & 31) < C4 ? (tid >> 5) * C4 + (tid & 31) : -1) : (tid < T ? tid : -1);
        if (slot >= 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                part[slot * 8 + j] = s1[j];
                part[slot * 8 + 4 + j] = s2[j];
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < 2 * C && n_owners > 0; t += blockDim.x) {
        const int which = t >= C, c = which ? t - C : t, q = c >> 2, j = c & 3;
        double acc = 0.0;
        for (int k = q; k < n_owners; k += C4) acc += part[k * 8 + which * 4 + j];
        atomicAdd(&sums[((long)n * c_tot + c_off + c) * 2 + which], acc * weight);
    }
}

// Small samples (one CTA per sample is mostly launch, barrier and reduction latency: 0.3 - 1.7 TB/s on the 4^3 and 2^3
// levels): one WARP per sample, eight float4 loads in flight per lane, no shared memory and no block barrier.  Needs
// 32 % (C / 4) == 0 so that a lane keeps one channel quad.
template <int U>
__device__ __forceinline__ void gn_warp_batch(const float4* __restrict__ p, double (&s1)[4], double (&s2)[4]) {
    float4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldg(p + 32 * k);
    // value-preserving dependency of the first use on all U loads (see cl_gn_partial_vec4_kernel)
    float d = v[0].x;
#pragma unroll
    for (int k = 1; k < U; ++k) d = fmaf(0.f, v[k].w, d);
    v[0].x = d;
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const double a0 = (double)v[k].x, a1 = (double)v[k].y, a2 = (double)v[k].z, a3 = (double)v[k].w;
        s1[0] += a0; s2[0] += a0 * a0;
        s1[1] += a1; s2[1] += a1 * a1;
        s1[2] += a2; s2[2] += a2 * a2;
        s1[3] += a3; s2[3] += a3 * a3;
    }
}

// direct != 0: this launch is the only writer of its (sample, channel) sums - plain stores, no memset needed before.
__global__ void __launch_bounds__(256) cl_gn_partial_warp_kernel(const float* __restrict__ x, long S, int C, long N, double weight,
                                                                 int c_off, int c_tot, double* __restrict__ sums, int direct) {
    const long n = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    const int C4 = C >> 2;
    const long total4 = S * C4;
    const float4* base = reinterpret_cast<const float4*>(x + n * S * C);
    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
    long f = lane;
    for (; f + 7 * 32 < total4; f += 8 * 32) gn_warp_batch<8>(base + f, s1, s2);
    if (f + 3 * 32 < total4) { gn_warp_batch<4>(base + f, s1, s2); f += 4 * 32; }
    if (f + 32 < total4) { gn_warp_batch<2>(base + f, s1, s2); f += 2 * 32; }
    if (f < total4) gn_warp_batch<1>(base + f, s1, s2);
    for (int o = 16; o >= C4; o >>= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
        }
    }
    if (lane < C4) {
        double* dst = sums + (n * c_tot + c_off + 4 * lane) * 2;
        if (direct) {
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(dst + 2 * j) = make_double2(s1[j] * weight, s2[j] * weight);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(dst + 2 * j, s1[j] * weight);
                atomicAdd(dst + 2 * j + 1, s2[j] * weight);
            }
        }
    }
}

// One thread per (sample, group): adds the group's channel sums once (a thread per channel re-added them C / G times:
// 58 us on the 96-channel join of a 64-chunk step) and writes mean and rstd * gamma of its channels.
__global__ void __launch_bounds__(256) cl_gn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                                             float* __restrict__ mu_out, float* __restrict__ a_out, int N, int C,
                                                             int G, double count_per_channel, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (n, g)
    if (i >= N * G) return;
    const int n = i / G, g = i % G, cpg = C / G;
    const double2* sp = reinterpret_cast<const double2*>(sums) + (long)n * C + g * cpg;  // (sum, sum of squares) per channel
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < cpg; ++j) {
        const double2 v = sp[j];
        s1 += v.x;
        s2 += v.y;
    }
    const double cnt = count_per_channel * cpg;
    const double mean = s1 / cnt;
    double var = s2 / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)eps));
    const long o = (long)n * C + g * cpg;
    for (int j = 0; j < cpg; ++j) {
        mu_out[o + j] = mu;
        a_out[o + j] = rstd * __ldg(gamma + g * cpg + j);
    }
}

// x [N,S,C] fp32 -> (x - mu[n, c_off+c]) * a[n, c_off+c] + beta[c_off+c] -> fp16 hi/lo [N,S,Cp]; pad channels are zero.
// mu == nullptr: no normalisation (plain split).
__global__ void __launch_bounds__(256) cl_norm_split_kernel(const float* __restrict__ x, const float* __restrict__ mu,
                                                            const float* __restrict__ a, const float* __restrict__ beta,
                                                            int c_off, int c_tot, uint16_t* __restrict__ hi,
                                                            uint16_t* __restrict__ lo, long NS, long S, int C, int Cp,
                                                            float scale) {
    const int chunks = Cp / 8;
    const long total = NS * chunks;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long v = i / chunks;
        const int cc = (int)(i % chunks);
        const long n = v / S;
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            float f[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int c = cc * 8 + e + t;
                float val = 0.f;
                if (c < C) {
                    val = __ldg(x + v * C + c);
                    if (mu) {
                        const long si = n * c_tot + c_off + c;
                        val = fmaf(val - __ldg(mu + si), __ldg(a + si), __ldg(beta + c_off + c));
                    }
                }
                f[t] = val * scale;  // power of two: keeps the lo parts in fp16's normal range
            }
            uint32_t h0, l0, h1, l1;
            split_f16(f[0], h0, l0);
            split_f16(f[1], h1, l1);
            h[e / 2] = h0 | (h1 << 16);
            l[e / 2] = l0 | (l1 << 16);
        }
        *reinterpret_cast<uint4*>(hi + v * Cp + cc * 8) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + v * Cp + cc * 8) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// Single-channel tensors (the 1-channel TSDF inputs) stay unpadded: [N,S] fp16 hi / lo.
__global__ void __launch_bounds__(256) cl_norm_split_c1_kernel(const float* __restrict__ x, const float* __restrict__ mu,
                                                               const float* __restrict__ a, const float* __restrict__ beta,
                                                               uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, long NS, long S,
                                                               float scale) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < NS; i += (long)gridDim.x * blockDim.x) {
        float val = __ldg(x + i);
        if (mu) {
            const long n = i / S;
            val = fmaf(val - __ldg(mu + n), __ldg(a + n), __ldg(beta));
        }
        uint32_t h, l;
        split_f16(val * scale, h, l);
        hi[i] = (uint16_t)h;
        lo[i] = (uint16_t)l;
    }
}

__global__ void __launch_bounds__(256) cl_maxpool2_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int D,
                                                          int H, int W, int C, FastDiv fC, FastDiv fW, FastDiv fH, FastDiv fD) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    const long total = (long)N * Do * Ho * Wo * C;  // < 2^32 (checked on the host): multiply-high index decomposition
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        unsigned t = (unsigned)i;
        const int c = (int)fd_divmod(t, fC);
        const int w = (int)fd_divmod(t, fW);
        const int h = (int)fd_divmod(t, fH);
        const int d = (int)fd_divmod(t, fD);
        float m = -3.402823466e38f;
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx)
                    m = fmaxf(m, __ldg(x + (((((long)t * D + 2 * d + dz) * H + 2 * h + dy) * (long)W + 2 * w + dx) * C + c)));
        y[i] = m;
    }
}

// NCDHW <-> channels-last (small tensors at module boundaries)
// The same pooling for C % 4 == 0: a thread owns a channel quad of one output voxel - eight 16-byte loads, all in flight
// (the scalar kernel's running maximum lets ptxas sink each load to its use: 3.1 TB/s on the 8^3 level's 537 MB).
// fmaxf of the eight taps in the scalar kernel's order; NaN handling and signed zeros as fmaxf gives them there.
__global__ void __launch_bounds__(256) cl_maxpool2_vec4_kernel(const float4* __restrict__ x, float4* __restrict__ y, int N, int D,
                                                               int H, int W, int C4, FastDiv fC4, FastDiv fW, FastDiv fH, FastDiv fD) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    const long total = (long)N * Do * Ho * Wo * C4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        unsigned t = (unsigned)i;
        const int c = (int)fd_divmod(t, fC4);
        const int w = (int)fd_divmod(t, fW);
        const int h = (int)fd_divmod(t, fH);
        const int d = (int)fd_divmod(t, fD);
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dz = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
            v[k] = __ldg(x + (((((long)t * D + 2 * d + dz) * H + 2 * h + dy) * (long)W + 2 * w + dx) * C4 + c));
        }
        // first use depends on all eight loads (value-preserving: 0 * clamp(.) + x)
        float dep = v[0].x;
#pragma unroll
        for (int k = 1; k < 8; ++k) dep = fmaf(0.f, fminf(fmaxf(v[k].w, -1.f), 1.f), dep);
        v[0].x = dep;
        float4 m = make_float4(-3.402823466e38f, -3.402823466e38f, -3.402823466e38f, -3.402823466e38f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            m.x = fmaxf(m.x, v[k].x); m.y = fmaxf(m.y, v[k].y); m.z = fmaxf(m.z, v[k].z); m.w = fmaxf(m.w, v[k].w);
        }
        y[i] = m;
    }
}

__global__ void __launch_bounds__(256) cl_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, long N, long S,
                                                           int C, int to_cl) {
    const long total = N * S * C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        // i indexes the OUTPUT
        if (to_cl) {
            const int c = (int)(i % C);
            const long v = (i / C) % S, n = i / ((long)C * S);
            out[i] = __ldg(in + (n * C + c) * S + v);
        } else {
            const long v = i % S;
            const int c = (int)((i / S) % C);
            const long n = i / (S * (long)C);
            out[i] = __ldg(in + (n * S + v) * C + c);
        }
    }
}

// ----------------------------------------------------------------- weight image
// w [Cout, C1+C2, KS,KS,KS] -> [n_kb][hi|lo][Npad rows][128 B]; K index = chunk q * 8 + e,
// q = tap * CC + cc, padded channel slot cc*8+e: [0,Cp1) -> x channels, [Cp1, Cp1+Cp2) -> x2 channels.
__global__ void __launch_bounds__(256) tc_conv_weight_image_kernel(const float* __restrict__ w, int Cout, int C1, int C2,
                                                                   int Cp1, int Cp2, int KS, int Npad, int n_kb,
                                                                   float scale, uint8_t* __restrict__ img) {
    const int taps = KS * KS * KS, CC = (Cp1 + Cp2) / 8, Cin = C1 + C2;
    const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;  // (row n, chunk q) incl. padding
    const long total = (long)Npad * n_kb * 8;
    if (gid >= total) return;
    const int n = (int)(gid / (n_kb * 8)), q = (int)(gid % (n_kb * 8));
    uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
    if (Cin == 1) {  // tap-major mode: K index = tap (no channel padding)
        if (n < Cout) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int tap = q * 8 + e;
                uint32_t h, l;
                split_f16(tap < taps ? w[(long)n * taps + tap] * scale : 0.f, h, l);
                hi[e / 2] |= h << (16 * (e & 1));
                lo[e / 2] |= l << (16 * (e & 1));
            }
        }
    } else if (n < Cout && q < taps * CC) {
        const int tap = q / CC, cc = q % CC;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int cs = cc * 8 + e;
            int ci = -1;
            if (cs < Cp1) { if (cs < C1) ci = cs; }
            else if (cs - Cp1 < C2) ci = C1 + (cs - Cp1);
            const float val = ci >= 0 ? w[((long)n * Cin + ci) * taps + tap] * scale : 0.f;
            uint32_t h, l;
            split_f16(val, h, l);
            hi[e / 2] |= h << (16 * (e & 1));
            lo[e / 2] |= l << (16 * (e & 1));
        }
    }
    const int kb = q / 8, c = q % 8;
    uint8_t* base = img + (long)kb * (2L * Npad * 128) + (long)n * 128 + ((c ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + (long)Npad * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct ConvArgs {
    const uint16_t *x_hi, *x_lo, *x2_hi, *x2_lo;
    const uint8_t* wimg;
    const float* bias;
    float* y;
    int N, Di, Hi, Wi, Do, Ho, Wo, KS, stride, pad, Cp1, Cp2, Cout, Npad, act, out_ncdhw;
    float slope, out_scale;  // out_scale = 1 / (activation scale * weight scale), applied to the accumulator
    int M, n_kb, total_chunks, a_stages, b_stages, cin1;
};

__global__ void __launch_bounds__(NTHREADS, 2) tc_conv3d_kernel(const ConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int A_STAGES = a.a_stages, B_STAGES = a.b_stages;
    const uint32_t b_bytes = 2u * (uint32_t)a.Npad * 128u;
    const uint32_t sA = base;
    const uint32_t sB = base + A_STAGES * 2 * IMG;
    const uint32_t bars = sB + B_STAGES * b_bytes;  // b_bytes is a multiple of 4 KiB
    const uint32_t bar_afull = bars, bar_aempty = bars + 8 * MAX_A_STAGES;
    const uint32_t bar_bfull = bars + 16 * MAX_A_STAGES, bar_bempty = bar_bfull + 8 * MAX_B_STAGES;
    const uint32_t bar_dfull = bar_bempty + 8 * MAX_B_STAGES;
    const uint32_t tmem_slot = bar_dfull + 8;
    uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TM;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < a.Npad) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(bar_afull + 8 * s, 128); mbar_init(bar_aempty + 8 * s, 1); }
        for (int s = 0; s < B_STAGES; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
        mbar_init(bar_dfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0) {  // ---- weight producer
            for (int kb = 0; kb < a.n_kb; ++kb) {
                const int s = kb % B_STAGES;
                mbar_wait(bar_bempty + 8 * s, ((uint32_t)(kb / B_STAGES) & 1u) ^ 1u);
                mbar_arrive_expect_tx(bar_bfull + 8 * s, b_bytes);
                bulk_g2s(sB + s * b_bytes, a.wimg + (long)kb * b_bytes, b_bytes, bar_bfull + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            const uint32_t idesc = idesc_f16(a.Npad);
            for (int kb = 0; kb < a.n_kb; ++kb) {
                const int sa = kb % A_STAGES, s = kb % B_STAGES;
                mbar_wait(bar_afull + 8 * sa, (uint32_t)(kb / A_STAGES) & 1u);
                mbar_wait(bar_bfull + 8 * s, (uint32_t)(kb / B_STAGES) & 1u);
                tc_fence_after();
                const uint32_t a_hi = sA + sa * 2 * IMG, a_lo = a_hi + IMG;
                const uint32_t b_hi = sB + s * b_bytes, b_lo = b_hi + (uint32_t)a.Npad * 128u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc_mma(tmem_base, umma_desc(a_hi + k * 32), umma_desc(b_hi + k * 32), idesc, (kb | k) ? 1u : 0u);
                    tc_mma(tmem_base, umma_desc(a_hi + k * 32), umma_desc(b_lo + k * 32), idesc, 1u);
                    tc_mma(tmem_base, umma_desc(a_lo + k * 32), umma_desc(b_hi + k * 32), idesc, 1u);
                }
                tc_commit(bar_bempty + 8 * s);
                tc_commit(bar_aempty + 8 * sa);
            }
            tc_commit(bar_dfull);
        }
    } else if (warp >= 4) {
        const int pw = warp - 4;
        // ---- A producer.  Lane -> (chunk j = lane % 8, row sub-slot lane / 8); 8 passes of 4 rows per warp.
        const int j = lane & 7;
        int r_n[8], r_d[8], r_h[8], r_w[8];  // per-row sample index and input origin (after stride / padding)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int m = m0 + pw * 32 + p * 4 + (lane >> 3);
            if (m < a.M) {
                int t = m;
                const int ow = t % a.Wo; t /= a.Wo;
                const int oh = t % a.Ho; t /= a.Ho;
                const int od = t % a.Do; t /= a.Do;
                r_n[p] = t; r_d[p] = od * a.stride - a.pad; r_h[p] = oh * a.stride - a.pad; r_w[p] = ow * a.stride - a.pad;
            } else {
                r_n[p] = -1; r_d[p] = r_h[p] = r_w[p] = 0;
            }
        }
        const int CC1 = a.Cp1 >> 3, CC = (a.Cp1 + a.Cp2) >> 3;
        const int D2 = a.Di >> 1, H2 = a.Hi >> 1, W2 = a.Wi >> 1;
        if (a.cin1) {
            // Single input channel: K index = tap.  x_hi / x_lo are unpadded [N,D,H,W] fp16; chunk q of a row holds
            // taps 8q .. 8q+7 of that output voxel's receptive field (scalar gathers, 27 / 125 taps in total).
            const int taps = a.KS * a.KS * a.KS;
            for (int kb = 0; kb < a.n_kb; ++kb) {
                const int sa = kb % A_STAGES;
                const int q = kb * 8 + j;
                mbar_wait(bar_aempty + 8 * sa, ((uint32_t)(kb / A_STAGES) & 1u) ^ 1u);
                uint8_t* img_hi = smem_al + (sA - base) + sa * 2 * IMG;
#pragma unroll 2
                for (int p = 0; p < 8; ++p) {
                    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
                    if (r_n[p] >= 0) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int tap = q * 8 + e;
                            const int kw = tap % a.KS, kh = (tap / a.KS) % a.KS, kd = tap / (a.KS * a.KS);
                            const int id = r_d[p] + kd, ih = r_h[p] + kh, iw = r_w[p] + kw;
                            if (tap < taps && id >= 0 && id < a.Di && ih >= 0 && ih < a.Hi && iw >= 0 && iw < a.Wi) {
                                const long off = (((long)r_n[p] * a.Di + id) * a.Hi + ih) * a.Wi + iw;
                                h[e >> 1] |= (uint32_t)__ldg(a.x_hi + off) << (16 * (e & 1));
                                l[e >> 1] |= (uint32_t)__ldg(a.x_lo + off) << (16 * (e & 1));
                            }
                        }
                    }
                    const int r = pw * 32 + p * 4 + (lane >> 3);
                    const uint32_t off = (uint32_t)r * 128u + ((uint32_t)(j ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(img_hi + IMG + off) = make_uint4(l[0], l[1], l[2], l[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(bar_afull + 8 * sa);
            }
        } else
        for (int kb = 0; kb < a.n_kb; ++kb) {
            const int sa = kb % A_STAGES;
            const int q = kb * 8 + j;
            const bool active = q < a.total_chunks;
            const int tap = active ? q / CC : 0, cc = active ? q % CC : 0;
            const int kw = tap % a.KS, kh = (tap / a.KS) % a.KS, kd = tap / (a.KS * a.KS);
            const bool from_x2 = cc >= CC1;
            uint4 vh[8], vl[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                vh[p] = make_uint4(0, 0, 0, 0);
                vl[p] = make_uint4(0, 0, 0, 0);
                const int id = r_d[p] + kd, ih = r_h[p] + kh, iw = r_w[p] + kw;
                if (active && r_n[p] >= 0 && id >= 0 && id < a.Di && ih >= 0 && ih < a.Hi && iw >= 0 && iw < a.Wi) {
                    long off;
                    const uint16_t *ph, *pl;
                    if (!from_x2) {
                        off = ((((long)r_n[p] * a.Di + id) * a.Hi + ih) * a.Wi + iw) * a.Cp1 + cc * 8;
                        ph = a.x_hi; pl = a.x_lo;
                    } else {
                        off = ((((long)r_n[p] * D2 + (id >> 1)) * H2 + (ih >> 1)) * W2 + (iw >> 1)) * a.Cp2 + (cc - CC1) * 8;
                        ph = a.x2_hi; pl = a.x2_lo;
                    }
                    vh[p] = __ldg(reinterpret_cast<const uint4*>(ph + off));
                    vl[p] = __ldg(reinterpret_cast<const uint4*>(pl + off));
                }
            }
            mbar_wait(bar_aempty + 8 * sa, ((uint32_t)(kb / A_STAGES) & 1u) ^ 1u);
            uint8_t* img_hi = smem_al + (sA - base) + sa * 2 * IMG;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int r = pw * 32 + p * 4 + (lane >> 3);
                const uint32_t off = (uint32_t)r * 128u + ((uint32_t)(j ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(img_hi + off) = vh[p];
                *reinterpret_cast<uint4*>(img_hi + IMG + off) = vl[p];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
            mbar_arrive(bar_afull + 8 * sa);
        }
        // ---- epilogue: thread <-> output voxel
        mbar_wait(bar_dfull, 0);
        tc_fence_after();
        const int m = m0 + pw * 32 + lane;
        const int So = a.Do * a.Ho * a.Wo;
        for (int c0 = 0; c0 < a.Npad; c0 += 32) {
            float v[32];
            tc_ld32(tmem_base + ((uint32_t)(pw * 32) << 16) + (uint32_t)c0, v);
            if (m < a.M) {
#pragma unroll
                for (int t = 0; t < 32; ++t) v[t] = fmaf(v[t], a.out_scale, (a.bias && c0 + t < a.Cout) ? __ldg(a.bias + c0 + t) : 0.f);
                rf_act_vec(v, a.act, a.slope);
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const int co = c0 + t;
                    if (co < a.Cout) {
                        if (a.out_ncdhw) a.y[((long)(m / So) * a.Cout + co) * So + (m % So)] = v[t];
                        else a.y[(long)m * a.Cout + co] = v[t];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// K blocks of 64: (taps x padded channel chunks) / 8, or taps / 64 in the single-channel tap-major mode
int conv_total_chunks(int C1, int C2, int KS) {
    const int taps = KS * KS * KS;
    if (C1 + C2 == 1) return (taps + 7) / 8;
    return taps * ((C1 + 7) / 8 + (C2 + 7) / 8);
}
int conv_smem(int a_stages, int b_stages, int npad) { return 1024 + a_stages * 2 * IMG + b_stages * 2 * npad * 128 + 256; }
int round_up(int v, int m) { return (v + m - 1) / m * m; }
int conv_npad(int cout) { return round_up(cout, 16) < 32 ? 32 : round_up(cout, 16); }

}  // namespace

extern "C" size_t rf_cl_gn_stats_workspace_bytes(int N, int C) { return (size_t)N * C * 2 * sizeof(double); }

extern "C" int rf_cl_gn_stats(const float* x, const float* x2, int C2, const float* gamma, float* gn_mu, float* gn_a, int N,
                              int C, int D, int H, int W, int groups, float eps, void* workspace, void* stream) {
    RF_CHECK_ARG(gamma && gn_mu && gn_a && workspace && (x || C2 == C), "rf_cl_gn_stats: null pointer");
    RF_CHECK_ARG(N > 0 && C > 0 && groups > 0 && C % groups == 0 && C2 >= 0 && C2 <= C && (C2 == 0 || x2), "rf_cl_gn_stats: bad channels");
    RF_CHECK_ARG(C2 == 0 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "rf_cl_gn_stats: upsampled input needs even extents");
    cudaStream_t s = (cudaStream_t)stream;
    double* sums = (double*)workspace;
    const long S = (long)D * H * W;
    static const int mode = [] { const char* e = getenv("RF_GN_MODE"); return e ? atoi(e) : 2; }();  // tuning aid: 0 scalar, 1 float4 loads, 2 warp per small sample / bulk copies, 3 bulk copies
    // many small samples: one warp per sample (a lane must keep one channel quad: 32 % (C / 4) == 0)
    auto warp_ok = [&](const float* src, long Ssrc, int Csrc) {
        const int C4src = Csrc >> 2;
        return mode == 2 && Csrc > 0 && (Csrc & 3) == 0 && C4src <= 32 && (32 % C4src) == 0 && ((uintptr_t)src & 15) == 0 &&
               N >= 148 * 8 && Ssrc * Csrc * 4 <= 64 * 1024;
    };
    const int C1 = C - C2;
    // every (sample, channel) sum has exactly one writer when all sources take the warp kernel: plain stores, no memset
    const bool direct = (C1 == 0 || warp_ok(x, S, C1)) && (C2 == 0 || warp_ok(x2, S / 8, C2));
    if (!direct) RF_CUDA_OK(cudaMemsetAsync(sums, 0, rf_cl_gn_stats_workspace_bytes(N, C), s));
    auto launch = [&](const float* src, long Ssrc, int Csrc, double weight, int c_off) -> int {
        // enough CTAs to fill the chip, but at least ~4k elements per CTA
        long slices = (148L * 8 + N - 1) / N;
        const long max_slices = (Ssrc * Csrc + 4095) / 4096;
        if (slices > max_slices) slices = max_slices;
        if (slices < 1) slices = 1;
        if (warp_ok(src, Ssrc, Csrc)) {
            cl_gn_partial_warp_kernel<<<(unsigned)rf_cdivl((long)N * 32, 256), 256, 0, s>>>(src, Ssrc, Csrc, N, weight, c_off, C, sums, direct ? 1 : 0);
            RF_LAUNCH_OK("cl_gn_partial_warp_kernel");
            return 0;
        }
        if (mode >= 2 && (Csrc & 3) == 0 && Csrc <= 512 && ((uintptr_t)src & 15) == 0) {
            const size_t smem = (size_t)GN_STAGES * GN_CHUNK4 * 16 + 8 * GN_STAGES;
            cl_gn_partial_bulk_kernel<<<(unsigned)(N * slices), 256, smem, s>>>(src, Ssrc, Csrc, (int)slices, weight, c_off, C, sums);
            RF_LAUNCH_OK("cl_gn_partial_bulk_kernel");
            return 0;
        }
        if (mode >= 1 && (Csrc & 3) == 0 && Csrc <= 1024 && ((uintptr_t)src & 15) == 0) {
            cl_gn_partial_vec4_kernel<<<(unsigned)(N * slices), 256, 2 * Csrc * sizeof(double), s>>>(src, Ssrc, Csrc, (int)slices, weight, c_off, C, sums);
            RF_LAUNCH_OK("cl_gn_partial_vec4_kernel");
            return 0;
        }
        cl_gn_partial_kernel<<<(unsigned)(N * slices), 256, 2 * Csrc * sizeof(double), s>>>(src, Ssrc, Csrc, (int)slices, weight, c_off, C, sums);
        RF_LAUNCH_OK("cl_gn_partial_kernel");
        return 0;
    };
    if (C1 > 0) { const int rc = launch(x, S, C1, 1.0, 0); if (rc) return rc; }
    if (C2 > 0) { const int rc = launch(x2, S / 8, C2, 8.0, C1); if (rc) return rc; }
    cl_gn_finalize_kernel<<<rf_cdiv((long)N * groups, 256), 256, 0, s>>>(sums, gamma, gn_mu, gn_a, N, C, groups, (double)S, eps);
    RF_LAUNCH_OK("cl_gn_finalize_kernel");
    return 0;
}

extern "C" int rf_cl_norm_split(const float* x, const float* gn_mu, const float* gn_a, const float* gn_beta, int c_off,
                                int c_tot, void* hi, void* lo, long N, long S, int C, int Cp, float scale, void* stream) {
    if (C == 1 && Cp == 1) {  // single-channel input: unpadded fp16 hi / lo, consumed by the conv's tap-major mode
        RF_CHECK_ARG(x && hi && lo && N > 0 && S > 0 && c_off == 0 && c_tot == 1, "rf_cl_norm_split: bad arguments (C = 1)");
        cl_norm_split_c1_kernel<<<rf_grid_1d(N * S, 256), 256, 0, (cudaStream_t)stream>>>(x, gn_mu, gn_a, gn_beta, (uint16_t*)hi,
                                                                                         (uint16_t*)lo, N * S, S, scale);
        RF_LAUNCH_OK("cl_norm_split_c1_kernel");
        return 0;
    }
    RF_CHECK_ARG(x && hi && lo && N > 0 && S > 0 && C > 0 && Cp >= C && Cp % 8 == 0, "rf_cl_norm_split: bad arguments");
    RF_CHECK_ARG((gn_mu == nullptr) == (gn_a == nullptr) && (gn_mu == nullptr) == (gn_beta == nullptr), "rf_cl_norm_split: partial GroupNorm arguments");
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0, "rf_cl_norm_split: outputs must be 16-byte aligned");
    const long total = N * S * (Cp / 8);
    cl_norm_split_kernel<<<rf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(x, gn_mu, gn_a, gn_beta, c_off, c_tot, (uint16_t*)hi,
                                                                                  (uint16_t*)lo, N * S, S, C, Cp, scale);
    RF_LAUNCH_OK("cl_norm_split_kernel");
    return 0;
}

extern "C" int rf_cl_maxpool3d_2(const float* x, float* y, int N, int D, int H, int W, int C, void* stream) {
    RF_CHECK_ARG(x && y && N > 0 && C > 0 && D >= 2 && H >= 2 && W >= 2, "rf_cl_maxpool3d_2: bad arguments");
    const long total = (long)N * (D / 2) * (H / 2) * (W / 2) * C;
    RF_CHECK_ARG(total < (1L << 32), "rf_cl_maxpool3d_2: more than 2^32 elements");
    if ((C & 3) == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
        cl_maxpool2_vec4_kernel<<<rf_grid_1d(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(
            (const float4*)x, (float4*)y, N, D, H, W, C / 4, make_fastdiv(C / 4), make_fastdiv(W / 2), make_fastdiv(H / 2), make_fastdiv(D / 2));
        RF_LAUNCH_OK("cl_maxpool2_vec4_kernel");
        return 0;
    }
    cl_maxpool2_kernel<<<rf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, N, D, H, W, C, make_fastdiv(C), make_fastdiv(W / 2),
                                                                                 make_fastdiv(H / 2), make_fastdiv(D / 2));
    RF_LAUNCH_OK("cl_maxpool2_kernel");
    return 0;
}

extern "C" int rf_cl_transpose(const float* in, float* out, long N, long S, int C, int to_channels_last, void* stream) {
    RF_CHECK_ARG(in && out && N > 0 && S > 0 && C > 0, "rf_cl_transpose: bad arguments");
    cl_transpose_kernel<<<rf_grid_1d(N * S * C, 256), 256, 0, (cudaStream_t)stream>>>(in, out, N, S, C, to_channels_last);
    RF_LAUNCH_OK("cl_transpose_kernel");
    return 0;
}

extern "C" size_t rf_tc_conv_weight_image_bytes(int Cout, int C1, int C2, int KS) {
    if (Cout < 1 || Cout > 128 || C1 < 0 || C2 < 0 || C1 + C2 < 1 || KS < 1 || KS > 5) return 0;
    const int n_kb = (conv_total_chunks(C1, C2, KS) + 7) / 8;
    return (size_t)n_kb * 2 * conv_npad(Cout) * 128;
}

extern "C" int rf_tc_conv_weight_image(const float* w, int Cout, int C1, int C2, int KS, float scale, void* image,
                                       void* stream) {
    RF_CHECK_ARG(w && image, "rf_tc_conv_weight_image: null pointer");
    RF_CHECK_ARG(rf_tc_conv_weight_image_bytes(Cout, C1, C2, KS) > 0, "rf_tc_conv_weight_image: unsupported shape Cout=%d C1=%d C2=%d KS=%d", Cout, C1, C2, KS);
    RF_CHECK_ARG(((uintptr_t)image & 1023) == 0, "rf_tc_conv_weight_image: image must be 1024-byte aligned");
    const int Cp1 = round_up(C1, 8), Cp2 = round_up(C2, 8);
    const int n_kb = (conv_total_chunks(C1, C2, KS) + 7) / 8, npad = conv_npad(Cout);
    const long threads = (long)npad * n_kb * 8;
    tc_conv_weight_image_kernel<<<(unsigned)rf_cdivl(threads, 256), 256, 0, (cudaStream_t)stream>>>(w, Cout, C1, C2, Cp1, Cp2, KS, npad, n_kb,
                                                                                                   scale, (uint8_t*)image);
    RF_LAUNCH_OK("tc_conv_weight_image_kernel");
    return 0;
}

int rf_tc_conv_init() {
    RF_SMEM_OPT_IN(tc_conv3d_kernel, conv_smem(MAX_A_STAGES, MAX_B_STAGES, 128));
    return 0;
}

extern "C" int rf_tc_conv3d_fwd(const void* x_hi, const void* x_lo, int C1, const void* x2_hi, const void* x2_lo, int C2,
                                const void* weight_image, const float* bias, float* y, int N, int Di, int Hi, int Wi, int Cout,
                                int KS, int stride, int pad, int act, float slope, float out_scale, int out_ncdhw, void* stream) {
    RF_CHECK_ARG(weight_image && y && (C1 == 0 || (x_hi && x_lo)) && (C2 == 0 || (x2_hi && x2_lo)), "rf_tc_conv3d_fwd: null pointer");
    RF_CHECK_ARG(rf_tc_conv_weight_image_bytes(Cout, C1, C2, KS) > 0, "rf_tc_conv3d_fwd: unsupported shape Cout=%d C1=%d C2=%d KS=%d", Cout, C1, C2, KS);
    RF_CHECK_ARG(N > 0 && Di > 0 && Hi > 0 && Wi > 0 && stride >= 1 && pad >= 0, "rf_tc_conv3d_fwd: bad geometry");
    RF_CHECK_ARG(C2 == 0 || (Di % 2 == 0 && Hi % 2 == 0 && Wi % 2 == 0), "rf_tc_conv3d_fwd: upsampled input needs even extents");
    RF_CHECK_ARG(((uintptr_t)weight_image & 1023) == 0, "rf_tc_conv3d_fwd: weight image must be 1024-byte aligned");
    ConvArgs a;
    a.x_hi = (const uint16_t*)x_hi; a.x_lo = (const uint16_t*)x_lo; a.x2_hi = (const uint16_t*)x2_hi; a.x2_lo = (const uint16_t*)x2_lo;
    a.wimg = (const uint8_t*)weight_image; a.bias = bias; a.y = y;
    a.N = N; a.Di = Di; a.Hi = Hi; a.Wi = Wi; a.KS = KS; a.stride = stride; a.pad = pad;
    a.Do = (Di + 2 * pad - KS) / stride + 1; a.Ho = (Hi + 2 * pad - KS) / stride + 1; a.Wo = (Wi + 2 * pad - KS) / stride + 1;
    RF_CHECK_ARG(a.Do > 0 && a.Ho > 0 && a.Wo > 0, "rf_tc_conv3d_fwd: empty output");
    a.Cp1 = round_up(C1, 8); a.Cp2 = round_up(C2, 8); a.Cout = Cout; a.Npad = conv_npad(Cout); a.act = act; a.slope = slope;
    a.out_scale = out_scale; a.out_ncdhw = out_ncdhw;
    const long M = (long)N * a.Do * a.Ho * a.Wo;
    RF_CHECK_ARG(M < (1L << 31) - TM, "rf_tc_conv3d_fwd: too many output voxels");
    a.M = (int)M;
    a.total_chunks = conv_total_chunks(C1, C2, KS);
    a.n_kb = (a.total_chunks + 7) / 8;
    a.cin1 = (C1 + C2 == 1) ? 1 : 0;
    RF_CHECK_ARG(!a.cin1 || C1 == 1, "rf_tc_conv3d_fwd: a single input channel must come from x");
    a.a_stages = MAX_A_STAGES;
    a.b_stages = 2;  // Npad <= 64: 97 KiB per CTA -> two CTAs per SM; Npad = 128: 129 KiB, one CTA
    if (int rc = rf_tc_conv_init()) return rc;
    tc_conv3d_kernel<<<(unsigned)rf_cdivl(M, TM), NTHREADS, conv_smem(a.a_stages, a.b_stages, a.Npad), (cudaStream_t)stream>>>(a);
    RF_LAUNCH_OK("tc_conv3d_kernel");
    return 0;
}
