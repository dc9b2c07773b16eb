// fp32 implicit-GEMM 3D convolution / linear layer with fused input GroupNorm,
// fused (virtual) nearest-upsample + channel concat on the input side and a
// fused bias / per-channel affine / activation epilogue; GroupNorm statistics.
//
// GEMM view: Out[m, co] = sum_k A[m, k] * Wt[k, co]
//   m = (n, od, oh, ow) flattened, k = (ci, kd, kh, kw) flattened, A gathered
//   on the fly from the NCDHW input (never materialised: no im2col buffer).
// CTA tile 64(m) x 64(co) x 16(k), 256 threads, 4x4 outputs per thread, fp32
// FMA accumulation - this is the accuracy-first path (1e-4 vs the reference's
// fp32 CPU result); the tcgen05 split-bf16 path replaces it layer by layer.
#include "rf_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct ConvArgs {
    const float *x, *x2, *wt, *bias, *oscale, *oshift, *gn_mu, *gn_a, *gn_beta;
    float* y;
    int N, Cin, C1, C2, Di, Hi, Wi, Cout, Do, Ho, Wo, stride, pad, act;
    float slope;
    int M, Kg;
};

template <int KS>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvArgs a) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tm = tid % 16, tn = tid / 16;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // ---- A-tile loader: this thread always gathers for the same output voxel
    const int am = tid % BM, ak0 = tid / BM;  // ak0 in 0..3
    const int m = m0 + am;
    const bool valid_m = m < a.M;
    int n_idx = 0, id0 = 0, ih0 = 0, iw0 = 0;
    if (valid_m) {
        int t = m;
        const int ow = t % a.Wo; t /= a.Wo;
        const int oh = t % a.Ho; t /= a.Ho;
        const int od = t % a.Do; t /= a.Do;
        n_idx = t;
        id0 = od * a.stride - a.pad; ih0 = oh * a.stride - a.pad; iw0 = ow * a.stride - a.pad;
    }
    const long x_n = (long)n_idx * a.C1;
    const long x2_n = (long)n_idx * a.C2;
    const int D2 = a.Di >> 1, H2 = a.Hi >> 1, W2 = a.Wi >> 1;
    const int bn = tid % BN, bk0 = tid / BN;
    const int co_load = n0 + bn;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.Kg; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kl = ak0 + 4 * i;
            const int k = k0 + kl;
            float v = 0.f;
            if (valid_m && k < a.Kg) {
                const int kw = k % KS;
                const int kh = (k / KS) % KS;
                const int kd = (k / (KS * KS)) % KS;
                const int ci = k / (KS * KS * KS);
                const int id = id0 + kd, ih = ih0 + kh, iw = iw0 + kw;
                if (id >= 0 && id < a.Di && ih >= 0 && ih < a.Hi && iw >= 0 && iw < a.Wi) {
                    if (ci < a.C1) {
                        v = __ldg(a.x + (((x_n + ci) * a.Di + id) * a.Hi + ih) * (long)a.Wi + iw);
                    } else {
                        v = __ldg(a.x2 + (((x2_n + (ci - a.C1)) * D2 + (id >> 1)) * H2 + (ih >> 1)) * (long)W2 + (iw >> 1));
                    }
                    if (a.gn_mu != nullptr) {
                        const long gi = (long)n_idx * a.Cin + ci;
                        v = fmaf(v - __ldg(a.gn_mu + gi), __ldg(a.gn_a + gi), __ldg(a.gn_beta + ci));
                    }
                }
            }
            As[kl][am] = v;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kl = bk0 + 4 * i;
            const int k = k0 + kl;
            Bs[kl][bn] = (k < a.Kg && co_load < a.Cout) ? __ldg(a.wt + (long)k * a.Cout + co_load) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float ar[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ar[i] = As[kk][tm + 16 * i];
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tn * 4]);
            const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int spatial = a.Do * a.Ho * a.Wo;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int mo = m0 + tm + 16 * i;
        if (mo >= a.M) continue;
        const int nn = mo / spatial, sp = mo - nn * spatial;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tn * 4 + j;
            if (co >= a.Cout) continue;
            float v = acc[i][j];
            if (a.bias) v += __ldg(a.bias + co);
            if (a.oscale) v *= __ldg(a.oscale + co);
            if (a.oshift) v += __ldg(a.oshift + co);
            a.y[((long)nn * a.Cout + co) * spatial + sp] = rf_act(v, a.act, a.slope);
        }
    }
}

// One CTA per (sample, group): mean and rstd over the group's elements of the
// virtual input (x ++ upsample2(x2)); two passes, fp64 accumulation.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, const float* __restrict__ x2, int C1,
                                                       int C2, const float* __restrict__ gamma, float* __restrict__ mu_out,
                                                       float* __restrict__ a_out, int C, int D, int H, int W, int G,
                                                       float eps) {
    const int n = blockIdx.x / G, g = blockIdx.x % G;
    const int cpg = C / G;
    const long sp = (long)D * H * W;
    const long count = cpg * sp;
    const int D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    __shared__ double red[32];
    __shared__ double bcast;

    auto load = [&](long e) -> float {
        const int c = g * cpg + (int)(e / sp);
        const long s = e % sp;
        if (c < C1) return __ldg(x + ((long)n * C1 + c) * sp + s);
        const int w = (int)(s % W), h = (int)((s / W) % H), d = (int)(s / ((long)W * H));
        return __ldg(x2 + ((((long)n * C2 + (c - C1)) * D2 + (d >> 1)) * H2 + (h >> 1)) * (long)W2 + (w >> 1));
    };
    auto block_sum = [&](double v) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();  // protect red/bcast from the previous use
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) bcast = t;
        }
        __syncthreads();
        return bcast;
    };

    double s1 = 0.0;
    for (long e = threadIdx.x; e < count; e += blockDim.x) s1 += (double)load(e);
    const double mean = block_sum(s1) / (double)count;
    double s2 = 0.0;
    for (long e = threadIdx.x; e < count; e += blockDim.x) {
        const double dlt = (double)load(e) - mean;
        s2 += dlt * dlt;
    }
    const double var = block_sum(s2) / (double)count;  // biased, as torch.nn.GroupNorm
    const double rstd = 1.0 / sqrt(var + (double)eps);
    for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
        const int ch = g * cpg + c;
        mu_out[(long)n * C + ch] = (float)mean;
        a_out[(long)n * C + ch] = (float)rstd * __ldg(gamma + ch);
    }
}

}  // namespace

extern "C" int rf_conv3d_fwd(const float* x, const float* x2, int C2, const float* wt, const float* bias,
                             const float* oscale, const float* oshift, const float* gn_mu, const float* gn_a,
                             const float* gn_beta, float* y, int N, int Cin, int Di, int Hi, int Wi, int Cout, int KS,
                             int stride, int pad, int act, float slope, void* stream) {
    RF_CHECK_ARG(wt && y && (x || C2 == Cin), "rf_conv3d_fwd: null pointer");
    RF_CHECK_ARG(N > 0 && Cin > 0 && Cout > 0 && Di > 0 && Hi > 0 && Wi > 0, "rf_conv3d_fwd: bad shape");
    RF_CHECK_ARG(KS >= 1 && KS <= 5 && stride >= 1 && pad >= 0, "rf_conv3d_fwd: kernel size %d / stride %d / pad %d unsupported", KS, stride, pad);
    RF_CHECK_ARG(C2 >= 0 && C2 <= Cin && (C2 == 0 || x2 != nullptr), "rf_conv3d_fwd: bad concat split C2=%d Cin=%d", C2, Cin);
    RF_CHECK_ARG(C2 == 0 || (Di % 2 == 0 && Hi % 2 == 0 && Wi % 2 == 0), "rf_conv3d_fwd: upsampled input needs even extents");
    RF_CHECK_ARG((gn_mu == nullptr) == (gn_a == nullptr) && (gn_mu == nullptr) == (gn_beta == nullptr),
                 "rf_conv3d_fwd: gn_mu / gn_a / gn_beta must be given together");
    RF_CHECK_ARG(act >= RF_ACT_NONE && act <= RF_ACT_TANH, "rf_conv3d_fwd: bad activation %d", act);
    ConvArgs a;
    a.x = x; a.x2 = x2; a.wt = wt; a.bias = bias; a.oscale = oscale; a.oshift = oshift;
    a.gn_mu = gn_mu; a.gn_a = gn_a; a.gn_beta = gn_beta; a.y = y;
    a.N = N; a.Cin = Cin; a.C2 = C2; a.C1 = Cin - C2; a.Di = Di; a.Hi = Hi; a.Wi = Wi; a.Cout = Cout;
    a.Do = (Di + 2 * pad - KS) / stride + 1; a.Ho = (Hi + 2 * pad - KS) / stride + 1; a.Wo = (Wi + 2 * pad - KS) / stride + 1;
    RF_CHECK_ARG(a.Do > 0 && a.Ho > 0 && a.Wo > 0, "rf_conv3d_fwd: empty output");
    a.stride = stride; a.pad = pad; a.act = act; a.slope = slope;
    const long M = (long)N * a.Do * a.Ho * a.Wo;
    RF_CHECK_ARG(M < (1L << 31) - BM, "rf_conv3d_fwd: too many output voxels (%ld)", M);
    a.M = (int)M; a.Kg = Cin * KS * KS * KS;
    dim3 grid(rf_cdiv(M, BM), rf_cdiv(Cout, BN));
    cudaStream_t s = (cudaStream_t)stream;
    switch (KS) {
        case 1: conv_igemm_kernel<1><<<grid, 256, 0, s>>>(a); break;
        case 2: conv_igemm_kernel<2><<<grid, 256, 0, s>>>(a); break;
        case 3: conv_igemm_kernel<3><<<grid, 256, 0, s>>>(a); break;
        case 4: conv_igemm_kernel<4><<<grid, 256, 0, s>>>(a); break;
        default: conv_igemm_kernel<5><<<grid, 256, 0, s>>>(a); break;
    }
    RF_LAUNCH_OK("conv_igemm_kernel");
    return 0;
}

extern "C" int rf_linear_fwd(const float* x, const float* wt, const float* bias, float* y, int M, int K, int N, int act,
                             float slope, void* stream) {
    return rf_conv3d_fwd(x, nullptr, 0, wt, bias, nullptr, nullptr, nullptr, nullptr, nullptr, y, M, K, 1, 1, 1, N, 1, 1, 0,
                         act, slope, stream);
}

extern "C" int rf_groupnorm_stats(const float* x, const float* x2, int C2, const float* gamma, float* gn_mu, float* gn_a,
                                  int N, int C, int D, int H, int W, int groups, float eps, void* stream) {
    RF_CHECK_ARG(gamma && gn_mu && gn_a && (x || C2 == C), "rf_groupnorm_stats: null pointer");
    RF_CHECK_ARG(N > 0 && C > 0 && groups > 0 && C % groups == 0, "rf_groupnorm_stats: C=%d not divisible by groups=%d", C, groups);
    RF_CHECK_ARG(C2 >= 0 && C2 <= C && (C2 == 0 || x2 != nullptr), "rf_groupnorm_stats: bad concat split");
    RF_CHECK_ARG(C2 == 0 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "rf_groupnorm_stats: upsampled input needs even extents");
    gn_stats_kernel<<<N * groups, 256, 0, (cudaStream_t)stream>>>(x, x2, C - C2, C2, gamma, gn_mu, gn_a, C, D, H, W, groups, eps);
    RF_LAUNCH_OK("gn_stats_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// First layers: single input channel (the TSDF / occupancy volume itself).
// With Cin = 1 a conv is 27 or 125 MACs per output channel - nothing for a GEMM
// to chew on - so this is a direct convolution: one thread per output voxel,
// the filter bank transposed in shared memory ([tap][CO], broadcast reads),
// CO accumulators in registers, GroupNorm(1 group) applied to in-bounds taps on
// the fly, bias + activation fused, fp32 channels-last output (what the
// tensor-core layers consume).
// ---------------------------------------------------------------------------
namespace {
template <int CO>
__global__ void __launch_bounds__(256) conv_cin1_direct_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, const float* __restrict__ gn_mu,
                                                               const float* __restrict__ gn_a, const float* __restrict__ gn_beta,
                                                               float* __restrict__ y, int Di, int Hi, int Wi, int Do, int Ho,
                                                               int Wo, int KS, int stride, int pad, int Cout, int act,
                                                               float slope, long M) {
    extern __shared__ float wsm[];  // [taps][CO]
    const int taps = KS * KS * KS;
    for (int i = threadIdx.x; i < taps * CO; i += blockDim.x) {
        const int t = i / CO, co = i % CO;
        wsm[i] = co < Cout ? __ldg(w + (long)co * taps + t) : 0.f;
    }
    __syncthreads();
    const float beta = gn_mu ? __ldg(gn_beta) : 0.f;
    for (long m = blockIdx.x * (long)blockDim.x + threadIdx.x; m < M; m += (long)gridDim.x * blockDim.x) {
        long t = m;
        const int ow = (int)(t % Wo); t /= Wo;
        const int oh = (int)(t % Ho); t /= Ho;
        const int od = (int)(t % Do); t /= Do;
        const long n = t;
        const float mu = gn_mu ? __ldg(gn_mu + n) : 0.f, ga = gn_mu ? __ldg(gn_a + n) : 1.f;
        const float* xn = x + n * (long)Di * Hi * Wi;
        float acc[CO];
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[c] = 0.f;
        const int d0 = od * stride - pad, h0 = oh * stride - pad, w0 = ow * stride - pad;
        int tap = 0;
        for (int kd = 0; kd < KS; ++kd) {
            const int id = d0 + kd;
            for (int kh = 0; kh < KS; ++kh) {
                const int ih = h0 + kh;
                const bool row_ok = id >= 0 && id < Di && ih >= 0 && ih < Hi;
                const float* xr = xn + ((long)id * Hi + ih) * Wi;
                for (int kw = 0; kw < KS; ++kw, ++tap) {
                    const int iw = w0 + kw;
                    float v = 0.f;
                    if (row_ok && iw >= 0 && iw < Wi) {
                        v = __ldg(xr + iw);
                        if (gn_mu) v = fmaf(v - mu, ga, beta);
                    }
                    const float4* wr = reinterpret_cast<const float4*>(wsm + tap * CO);
#pragma unroll
                    for (int c4 = 0; c4 < CO / 4; ++c4) {
                        const float4 w4 = wr[c4];
                        acc[4 * c4] = fmaf(v, w4.x, acc[4 * c4]);
                        acc[4 * c4 + 1] = fmaf(v, w4.y, acc[4 * c4 + 1]);
                        acc[4 * c4 + 2] = fmaf(v, w4.z, acc[4 * c4 + 2]);
                        acc[4 * c4 + 3] = fmaf(v, w4.w, acc[4 * c4 + 3]);
                    }
                }
            }
        }
        float* out = y + m * Cout;
        if (bias) {
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[c] += c < Cout ? __ldg(bias + c) : 0.f;
        }
        rf_act_vec(acc, act, slope);
        if (Cout == CO && (CO & 3) == 0) {  // one voxel = CO contiguous floats (16-byte aligned: m * CO * 4)
#pragma unroll
            for (int c = 0; c < CO; c += 4) *reinterpret_cast<float4*>(out + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        } else {
#pragma unroll
            for (int c = 0; c < CO; ++c)
                if (c < Cout) out[c] = acc[c];
        }
    }
}

// Register-tiled variant for stride 1: one thread computes WT consecutive outputs along w for all CO channels.  Per
// (kd, kh) it loads one row segment of WT + KS - 1 inputs and then needs only the two float4 weight loads per tap for
// WT x CO FMAs (the one-voxel kernel above issues three loads per eight FMAs and is load-issue bound: 13 TFLOP/s on the
// 5^3 first layer of Patch32).  Same tap order per output, hence bit-identical results.
template <int CO, int KS, int WT>
__global__ void __launch_bounds__(128) conv_cin1_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, const float* __restrict__ gn_mu,
                                                              const float* __restrict__ gn_a, const float* __restrict__ gn_beta,
                                                              float* __restrict__ y, int Di, int Hi, int Wi, int Do, int Ho,
                                                              int Wo, int pad, int Cout, int act, float slope, long M4) {
    __shared__ float wsm[KS * KS * KS * CO];  // [taps][CO]
    constexpr int taps = KS * KS * KS;
    for (int i = threadIdx.x; i < taps * CO; i += blockDim.x) {
        const int t = i / CO, co = i % CO;
        wsm[i] = co < Cout ? __ldg(w + (long)co * taps + t) : 0.f;
    }
    __syncthreads();
    const float beta = gn_mu ? __ldg(gn_beta) : 0.f;
    const int Wo4 = (Wo + WT - 1) / WT;
    for (long m4 = blockIdx.x * (long)blockDim.x + threadIdx.x; m4 < M4; m4 += (long)gridDim.x * blockDim.x) {
        long t = m4;
        const int ow0 = (int)(t % Wo4) * WT; t /= Wo4;
        const int oh = (int)(t % Ho); t /= Ho;
        const int od = (int)(t % Do); t /= Do;
        const long n = t;
        const float mu = gn_mu ? __ldg(gn_mu + n) : 0.f, ga = gn_mu ? __ldg(gn_a + n) : 1.f;
        const float* xn = x + n * (long)Di * Hi * Wi;
        float acc[WT][CO];
#pragma unroll
        for (int j = 0; j < WT; ++j)
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;
        const int d0 = od - pad, h0 = oh - pad, w0 = ow0 - pad;
#pragma unroll 1
        for (int kd = 0; kd < KS; ++kd) {
            const int id = d0 + kd;
#pragma unroll 1
            for (int kh = 0; kh < KS; ++kh) {
                const int ih = h0 + kh;
                const bool row_ok = id >= 0 && id < Di && ih >= 0 && ih < Hi;
                const float* xr = xn + ((long)id * Hi + ih) * Wi;
                float xv[WT + KS - 1];
#pragma unroll
                for (int i = 0; i < WT + KS - 1; ++i) {
                    const int iw = w0 + i;
                    float v = 0.f;
                    if (row_ok && iw >= 0 && iw < Wi) {
                        v = __ldg(xr + iw);
                        if (gn_mu) v = fmaf(v - mu, ga, beta);
                    }
                    xv[i] = v;
                }
                const float* wrow = wsm + (kd * KS + kh) * KS * CO;
#pragma unroll
                for (int kw = 0; kw < KS; ++kw) {
                    const float4* wr = reinterpret_cast<const float4*>(wrow + kw * CO);
#pragma unroll
                    for (int c4 = 0; c4 < CO / 4; ++c4) {
                        const float4 w4 = wr[c4];
#pragma unroll
                        for (int j = 0; j < WT; ++j) {
                            acc[j][4 * c4] = fmaf(xv[j + kw], w4.x, acc[j][4 * c4]);
                            acc[j][4 * c4 + 1] = fmaf(xv[j + kw], w4.y, acc[j][4 * c4 + 1]);
                            acc[j][4 * c4 + 2] = fmaf(xv[j + kw], w4.z, acc[j][4 * c4 + 2]);
                            acc[j][4 * c4 + 3] = fmaf(xv[j + kw], w4.w, acc[j][4 * c4 + 3]);
                        }
                    }
                }
            }
        }
        const long m0 = ((n * Do + od) * Ho + oh) * (long)Wo + ow0;
#pragma unroll
        for (int j = 0; j < WT; ++j) {
            if (ow0 + j < Wo) {
                float* out = y + (m0 + j) * Cout;
                if (bias) {
#pragma unroll
                    for (int c = 0; c < CO; ++c) acc[j][c] += c < Cout ? __ldg(bias + c) : 0.f;
                }
                rf_act_vec(acc[j], act, slope);
                if (Cout == CO) {
#pragma unroll
                    for (int c = 0; c < CO; c += 4) *reinterpret_cast<float4*>(out + c) = make_float4(acc[j][c], acc[j][c + 1], acc[j][c + 2], acc[j][c + 3]);
                } else {
#pragma unroll
                    for (int c = 0; c < CO; ++c)
                        if (c < Cout) out[c] = acc[j][c];
                }
            }
        }
    }
}
}  // namespace

extern "C" int rf_conv3d_cin1_cl_fwd(const float* x, const float* w, const float* bias, const float* gn_mu, const float* gn_a,
                                     const float* gn_beta, float* y, int N, int Di, int Hi, int Wi, int Cout, int KS,
                                     int stride, int pad, int act, float slope, void* stream) {
    RF_CHECK_ARG(x && w && y, "rf_conv3d_cin1_cl_fwd: null pointer");
    RF_CHECK_ARG(N > 0 && Di > 0 && Hi > 0 && Wi > 0 && Cout >= 1 && Cout <= 32 && KS >= 1 && KS <= 5 && stride >= 1 && pad >= 0,
                 "rf_conv3d_cin1_cl_fwd: unsupported shape (Cout <= 32, kernel <= 5)");
    RF_CHECK_ARG((gn_mu == nullptr) == (gn_a == nullptr) && (gn_mu == nullptr) == (gn_beta == nullptr), "rf_conv3d_cin1_cl_fwd: partial GroupNorm arguments");
    const int Do = (Di + 2 * pad - KS) / stride + 1, Ho = (Hi + 2 * pad - KS) / stride + 1, Wo = (Wi + 2 * pad - KS) / stride + 1;
    RF_CHECK_ARG(Do > 0 && Ho > 0 && Wo > 0, "rf_conv3d_cin1_cl_fwd: empty output");
    const long M = (long)N * Do * Ho * Wo;
    const int taps = KS * KS * KS;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = rf_grid_1d(M, 256, 148 * 32);
    if (stride == 1 && (KS == 3 || KS == 5) && Cout <= 16) {
        constexpr int WT = 4;
        const long M4 = (long)N * Do * Ho * ((Wo + WT - 1) / WT);
        const int g4 = rf_grid_1d(M4, 128, 148 * 64);
        if (Cout <= 8 && KS == 3) conv_cin1_tiled_kernel<8, 3, WT><<<g4, 128, 0, s>>>(x, w, bias, gn_mu, gn_a, gn_beta, y, Di, Hi, Wi, Do, Ho, Wo, pad, Cout, act, slope, M4);
        else if (Cout <= 8) conv_cin1_tiled_kernel<8, 5, WT><<<g4, 128, 0, s>>>(x, w, bias, gn_mu, gn_a, gn_beta, y, Di, Hi, Wi, Do, Ho, Wo, pad, Cout, act, slope, M4);
        else if (KS == 3) conv_cin1_tiled_kernel<16, 3, WT><<<g4, 128, 0, s>>>(x, w, bias, gn_mu, gn_a, gn_beta, y, Di, Hi, Wi, Do, Ho, Wo, pad, Cout, act, slope, M4);
        else conv_cin1_tiled_kernel<16, 5, WT><<<g4, 128, 0, s>>>(x, w, bias, gn_mu, gn_a, gn_beta, y, Di, Hi, Wi, Do, Ho, Wo, pad, Cout, act, slope, M4);
        RF_LAUNCH_OK("conv_cin1_tiled_kernel");
        return 0;
    }
    if (Cout <= 8)
        conv_cin1_direct_kernel<8><<<grid, 256, taps * 8 * sizeof(float), s>>>(x, w, bias, gn_mu, gn_a, gn_beta, y, Di, Hi, Wi, Do, Ho, Wo, KS, stride, pad, Cout, act, slope, M);
    else if (Cout <= 16)
        conv_cin1_direct_kernel<16><<<grid, 256, taps * 16 * sizeof(float), s>>>(x, w, bias, gn_mu, gn_a, gn_beta, y, Di, Hi, Wi, Do, Ho, Wo, KS, stride, pad, Cout, act, slope, M);
    else
        conv_cin1_direct_kernel<32><<<grid, 256, taps * 32 * sizeof(float), s>>>(x, w, bias, gn_mu, gn_a, gn_beta, y, Di, Hi, Wi, Do, Ho, Wo, KS, stride, pad, Cout, act, slope, M);
    RF_LAUNCH_OK("conv_cin1_direct_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// F.normalize(x, dim=1): one warp per row.
// ---------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) l2norm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, long M, int D,
                                                          float eps) {
    const long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* xr = x + row * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) { const float v = xr[i]; s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float nrm = fmaxf(sqrtf(s), eps);
    for (int i = lane; i < D; i += 32) y[row * D + i] = xr[i] / nrm;
}
}  // namespace

extern "C" int rf_l2_normalize_rows(const float* x, float* y, long M, int D, float eps, void* stream) {
    RF_CHECK_ARG(x && y && M > 0 && D > 0, "rf_l2_normalize_rows: bad arguments");
    const long threads = M * 32;
    l2norm_rows_kernel<<<(unsigned)rf_cdivl(threads, 256), 256, 0, (cudaStream_t)stream>>>(x, y, M, D, eps);
    RF_LAUNCH_OK("l2norm_rows_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// ReLU MLP encoder (Patch04 / Patch05 / Patch04V2) + row normalisation.
// ---------------------------------------------------------------------------
extern "C" size_t rf_mlp_encode_workspace_bytes(long M, const int* widths_host, int n_layers) {
    if (M <= 0 || !widths_host || n_layers < 1 || n_layers > 8) return 0;
    size_t wmax = 0;
    for (int l = 1; l <= n_layers; ++l) if ((size_t)widths_host[l] > wmax) wmax = widths_host[l];
    const size_t one = ((size_t)M * wmax * sizeof(float) + 255) / 256 * 256;
    return 2 * one;
}

extern "C" int rf_mlp_encode_fwd(const float* x, const float* const* wt_host, const float* const* bias_host,
                                 const int* widths_host, int n_layers, int l2_normalize, float* out, long M,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    RF_CHECK_ARG(x && wt_host && bias_host && widths_host && out && workspace, "rf_mlp_encode_fwd: null pointer");
    RF_CHECK_ARG(n_layers >= 1 && n_layers <= 8 && M > 0 && M < (1L << 31), "rf_mlp_encode_fwd: bad sizes");
    const size_t need = rf_mlp_encode_workspace_bytes(M, widths_host, n_layers);
    RF_CHECK_ARG(workspace_bytes >= need && ((uintptr_t)workspace & 255) == 0, "rf_mlp_encode_fwd: workspace too small or misaligned (%zu < %zu)", workspace_bytes, need);
    float* buf[2] = {(float*)workspace, (float*)((char*)workspace + need / 2)};
    const float* cur = x;
    for (int l = 0; l < n_layers; ++l) {
        const bool last = l == n_layers - 1;
        float* dst = (last && !l2_normalize) ? out : buf[l & 1];
        const int rc = rf_linear_fwd(cur, wt_host[l], bias_host[l], dst, (int)M, widths_host[l], widths_host[l + 1],
                                     last ? RF_ACT_NONE : RF_ACT_RELU, 0.f, stream);
        if (rc) return rc;
        cur = dst;
    }
    if (l2_normalize) return rf_l2_normalize_rows(cur, out, M, widths_host[n_layers], 1e-12f, stream);
    return 0;
}

// ---------------------------------------------------------------------------
// Pointwise head: Conv3d(C, 1, 1) + bias + activation on a channels-last volume (model/refinement.py:55-57, the
// final 1x1x1 conv + Tanh of the decoder).  Memory-bound: one thread per voxel reads its C contiguous floats as
// float4s and writes one float; fp32 FMA in channel order.
// ---------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) cl_pointwise_head_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ y, long total,
                                                                int C, int act, float slope) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const float* src = x + i * C;
        float acc = 0.f;
        if ((C & 3) == 0) {
            for (int c = 0; c < C; c += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
                acc = fmaf(v.x, __ldg(w + c), acc);
                acc = fmaf(v.y, __ldg(w + c + 1), acc);
                acc = fmaf(v.z, __ldg(w + c + 2), acc);
                acc = fmaf(v.w, __ldg(w + c + 3), acc);
            }
        } else {
            for (int c = 0; c < C; ++c) acc = fmaf(__ldg(src + c), __ldg(w + c), acc);
        }
        y[i] = rf_act(acc + (bias ? __ldg(bias) : 0.f), act, slope);
    }
}
}  // namespace

extern "C" int rf_cl_pointwise_head(const float* x, const float* w, const float* bias, float* y, long n_voxels, int C, int act,
                                    float slope, void* stream) {
    RF_CHECK_ARG(x && w && y && n_voxels > 0 && C > 0, "rf_cl_pointwise_head: bad arguments");
    RF_CHECK_ARG(((uintptr_t)x & 15) == 0, "rf_cl_pointwise_head: x must be 16-byte aligned");
    cl_pointwise_head_kernel<<<rf_grid_1d(n_voxels, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, n_voxels, C, act, slope);
    RF_LAUNCH_OK("cl_pointwise_head_kernel");
    return 0;
}
