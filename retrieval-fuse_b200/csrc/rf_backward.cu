// Backward passes of the drop-in modules (trainer/train_refinement.py:74-89 training_step_full back-
// propagates through the U-Nets, the patch attention and the decoder).  The forward of a module that is being
// differentiated runs on the fp32 NCDHW kernels (rf_conv.cu); the kernels here are their adjoints, fp32 with fp64
// reductions where a sum runs over a whole volume:
//
//   conv (k^3, stride 1)   d/dx  = the forward kernel itself on the flipped, transposed filter (host-side re-layout)
//                          d/dW  = rf_conv3d_wgrad: GEMM over all output positions, the im2col operand gathered on the
//                                  fly with the input GroupNorm applied (same gather as the forward)
//   activation             rf_act_bwd (ReLU / LeakyReLU / tanh expressed through the saved OUTPUT)
//   GroupNorm              rf_gn_bwd_reduce (sum g, sum g x^ per sample and channel) -> rf_gn_bwd_apply
//                          (dx = rstd (gamma g - mean(gamma g) - x^ mean(gamma g x^)); the nearest-upsampled half of a
//                          decoder join sums its 8 children)
//   MaxPool3d(2)           rf_maxpool3d_2_bwd (first maximum in scan order, as torch)
//   Linear                 d/dx = forward kernel on W, d/dW = rf_conv3d_wgrad with k = 1, d/db = rf_channel_sum
//   attention epilogue     rf_attention_epilogue_bwd (normalise, scores, ReLU-max switch, softmax(1024 s) or hard
//                          Gumbel with the straight-through soft gradient, weighted sum, blend), one warp per row
#include <float.h>

#include "rf_common.cuh"

namespace {

// --------------------------------------------------------------------------------------------- conv weight gradient
constexpr int WK = 64, WN = 64, WM = 16;  // tile: 64 filter taps (ci,kd,kh,kw) x 64 output channels, 16 positions per step

struct WgradArgs {
    const float *x, *x2, *gn_mu, *gn_a, *gn_beta, *dz;
    float* dw;
    int N, Cin, C1, C2, Di, Hi, Wi, Cout, Do, Ho, Wo, KS, stride, pad;
    long M;
    int Kg, m_per_slice;
};

// dW[co, k] += sum_m A[m, k] dz[m, co];  A[m, k] = normalised, zero-padded input at (position m, tap k)
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradArgs a) {
    __shared__ float As[WM][WK + 1];
    __shared__ float Zs[WM][WN + 1];
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * WK, c0 = blockIdx.y * WN;
    const long m_begin = (long)blockIdx.z * a.m_per_slice;
    long m_end = m_begin + a.m_per_slice;
    if (m_end > a.M) m_end = a.M;
    const int tk = tid % 16, tc = tid / 16;  // thread owns taps tk + 16 i, channels tc + 16 j
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int D2 = a.Di >> 1, H2 = a.Hi >> 1, W2 = a.Wi >> 1;
    const long spatial = (long)a.Do * a.Ho * a.Wo;
    const int ks3 = a.KS * a.KS * a.KS;
    for (long mb = m_begin; mb < m_end; mb += WM) {
        // A tile: 16 positions x 64 taps, thread -> (position tid % 16, taps tid / 16 + 16 i)
        {
            const int mm = tid % WM;
            const long m = mb + mm;
            const bool vm = m < m_end;
            int n = 0, id0 = 0, ih0 = 0, iw0 = 0;
            if (vm) {
                long t = m;
                const int ow = (int)(t % a.Wo); t /= a.Wo;
                const int oh = (int)(t % a.Ho); t /= a.Ho;
                const int od = (int)(t % a.Do); t /= a.Do;
                n = (int)t;
                id0 = od * a.stride - a.pad; ih0 = oh * a.stride - a.pad; iw0 = ow * a.stride - a.pad;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kl = tid / WM + 16 * i;
                const int k = k0 + kl;
                float v = 0.f;
                if (vm && k < a.Kg) {
                    const int ci = k / ks3, r = k % ks3;
                    const int kd = r / (a.KS * a.KS), kh = (r / a.KS) % a.KS, kw = r % a.KS;
                    const int id = id0 + kd, ih = ih0 + kh, iw = iw0 + kw;
                    if (id >= 0 && id < a.Di && ih >= 0 && ih < a.Hi && iw >= 0 && iw < a.Wi) {
                        if (ci < a.C1) v = __ldg(a.x + ((((long)n * a.C1 + ci) * a.Di + id) * a.Hi + ih) * (long)a.Wi + iw);
                        else v = __ldg(a.x2 + ((((long)n * a.C2 + (ci - a.C1)) * D2 + (id >> 1)) * H2 + (ih >> 1)) * (long)W2 + (iw >> 1));
                        if (a.gn_mu) {
                            const long gi = (long)n * a.Cin + ci;
                            v = fmaf(v - __ldg(a.gn_mu + gi), __ldg(a.gn_a + gi), __ldg(a.gn_beta + ci));
                        }
                    }
                }
                As[mm][kl] = v;
            }
            // dz tile: 16 positions x 64 channels
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cl = tid / WM + 16 * j;
                const int co = c0 + cl;
                float v = 0.f;
                if (vm && co < a.Cout) {
                    const long nn = m / spatial, sp = m % spatial;
                    v = __ldg(a.dz + (nn * a.Cout + co) * spatial + sp);
                }
                Zs[mm][cl] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < WM; ++mm) {
            float ar[4], zr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ar[i] = As[mm][tk + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) zr[j] = Zs[mm][tc + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], zr[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + tk + 16 * i;
        if (k >= a.Kg) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = c0 + tc + 16 * j;
            if (co < a.Cout) atomicAdd(a.dw + (long)co * a.Kg + k, acc[i][j]);
        }
    }
}

// --------------------------------------------------------------------------------------------- activation backward
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz,
                                                      long n, int act, float slope) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float g = dy[i], o = y[i];
        float r = g;
        if (act == RF_ACT_RELU) r = o > 0.f ? g : 0.f;
        else if (act == RF_ACT_LEAKY) r = o > 0.f ? g : g * slope;
        else if (act == RF_ACT_TANH) r = g * (1.f - o * o);
        dz[i] = r;
    }
}

// --------------------------------------------------------------------------------------------- channel sums
// x [N, C, V] -> out[c] = sum_{n, v} x[n, c, v]  (bias gradients; V = 1 for Linear layers)
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ x, int N, int C, long V, float* __restrict__ out) {
    const int c = blockIdx.x;
    double s = 0.0;
    const long total = (long)N * V;
    for (long e = blockIdx.y * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.y * blockDim.x) {
        const long n = e / V, v = e % V;
        s += (double)__ldg(x + (n * C + c) * V + v);
    }
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(out + c, (float)t);
    }
}

// --------------------------------------------------------------------------------------------- GroupNorm backward
// virtual input = concat(x [N,C1,D,H,W], up2(x2 [N,C2,D/2,H/2,W/2])); g = dL/d(normalised input) [N,C,D,H,W];
// mu, rstd [N,C] (per channel copies of the group's statistics).  One CTA per (n, c):
//   s1[n,c] = sum_v g,  s2[n,c] = sum_v g x^   (x^ = (x - mu) rstd)
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ x2, int C1, int C2,
                                                            const float* __restrict__ g, const float* __restrict__ mu,
                                                            const float* __restrict__ rstd, int D, int H, int W,
                                                            double* __restrict__ s12) {
    const int C = C1 + C2;
    const int n = blockIdx.x / C, c = blockIdx.x % C;
    const long V = (long)D * H * W;
    const float m = mu[(long)n * C + c], r = rstd[(long)n * C + c];
    const float* gp = g + ((long)n * C + c) * V;
    const int H2 = H >> 1, W2 = W >> 1, D2 = D >> 1;
    double a1 = 0.0, a2 = 0.0;
    for (long v = threadIdx.x; v < V; v += blockDim.x) {
        float xv;
        if (c < C1) xv = __ldg(x + ((long)n * C1 + c) * V + v);
        else {
            const int w = (int)(v % W), h = (int)((v / W) % H), d = (int)(v / ((long)W * H));
            xv = __ldg(x2 + ((((long)n * C2 + (c - C1)) * D2 + (d >> 1)) * H2 + (h >> 1)) * (long)W2 + (w >> 1));
        }
        const float gv = __ldg(gp + v);
        a1 += (double)gv;
        a2 += (double)gv * (double)((xv - m) * r);
    }
    __shared__ double red[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a1; red[1][threadIdx.x >> 5] = a2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
        s12[((long)n * C + c) * 2] = t1;
        s12[((long)n * C + c) * 2 + 1] = t2;
    }
}

// per (n, c): dx = k1 g + k2 x^ + k3 with  k1 = rstd gamma,  k2 = -rstd B / m,  k3 = -rstd A / m,
//   A = sum_{c' in group} gamma_c' s1[n,c'],  B = sum gamma_c' s2[n,c'],  m = channels per group x voxels;
// also dgamma[c] += s2[n,c], dbeta[c] += s1[n,c]
__global__ void __launch_bounds__(128) gn_bwd_coeff_kernel(const double* __restrict__ s12, const float* __restrict__ gamma,
                                                           const float* __restrict__ rstd, int N, int C, int G, double voxels,
                                                           float* __restrict__ coef, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C) return;
    const int n = i / C, c = i % C, cpg = C / G, g0 = (c / cpg) * cpg;
    double A = 0.0, B = 0.0;
    for (int j = 0; j < cpg; ++j) {
        const double gm = (double)gamma[g0 + j];
        A += gm * s12[((long)n * C + g0 + j) * 2];
        B += gm * s12[((long)n * C + g0 + j) * 2 + 1];
    }
    const double m = voxels * cpg, r = (double)rstd[i];
    coef[(long)i * 3] = (float)(r * (double)gamma[c]);
    coef[(long)i * 3 + 1] = (float)(-r * B / m);
    coef[(long)i * 3 + 2] = (float)(-r * A / m);
    if (dgamma) atomicAdd(dgamma + c, (float)s12[(long)i * 2 + 1]);
    if (dbeta) atomicAdd(dbeta + c, (float)s12[(long)i * 2]);
}

// dx [N,C1,D,H,W] (fine part) and dx2 [N,C2,D/2,H/2,W/2] (the nearest-upsampled part: sum over the 8 children)
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ x2, int C1, int C2,
                                                           const float* __restrict__ g, const float* __restrict__ mu,
                                                           const float* __restrict__ rstd, const float* __restrict__ coef,
                                                           int N, int D, int H, int W, float* __restrict__ dx,
                                                           float* __restrict__ dx2) {
    const int C = C1 + C2;
    const long V = (long)D * H * W;
    const long n1 = (long)N * C1 * V;
    const int D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    const long V2 = (long)D2 * H2 * W2;
    const long n2 = (long)N * C2 * V2;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n1 + n2; i += (long)gridDim.x * blockDim.x) {
        if (i < n1) {
            const long v = i % V;
            const long nc = i / V;
            const int n = (int)(nc / C1), c = (int)(nc % C1);
            const long si = (long)n * C + c;
            const float xh = (x[i] - mu[si]) * rstd[si];
            dx[i] = fmaf(coef[si * 3], g[si * V + v], fmaf(coef[si * 3 + 1], xh, coef[si * 3 + 2]));
        } else {
            const long j = i - n1;
            const long v2 = j % V2;
            const long nc = j / V2;
            const int n = (int)(nc / C2), c = (int)(nc % C2);
            const long si = (long)n * C + C1 + c;
            const int w2 = (int)(v2 % W2), h2 = (int)((v2 / W2) % H2), d2 = (int)(v2 / ((long)W2 * H2));
            const float xh = (x2[j] - mu[si]) * rstd[si];
            float gs = 0.f;
#pragma unroll
            for (int dz = 0; dz < 2; ++dz)
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                    for (int dxx = 0; dxx < 2; ++dxx)
                        gs += g[si * V + ((long)(2 * d2 + dz) * H + 2 * h2 + dy) * W + 2 * w2 + dxx];
            dx2[j] = fmaf(coef[si * 3], gs, 8.f * fmaf(coef[si * 3 + 1], xh, coef[si * 3 + 2]));
        }
    }
}

// nearest x2 upsampling backward alone (inputs without GroupNorm): dx2 = sum of the 8 children of g's channels [C1, C)
__global__ void __launch_bounds__(256) upsample2_bwd_kernel(const float* __restrict__ g, int N, int C, int C1, int D, int H, int W,
                                                            float* __restrict__ dx2) {
    const int C2 = C - C1, D2 = D >> 1, H2 = H >> 1, W2 = W >> 1;
    const long V = (long)D * H * W, V2 = (long)D2 * H2 * W2, total = (long)N * C2 * V2;
    for (long j = blockIdx.x * (long)blockDim.x + threadIdx.x; j < total; j += (long)gridDim.x * blockDim.x) {
        const long v2 = j % V2, nc = j / V2;
        const int n = (int)(nc / C2), c = (int)(nc % C2);
        const int w2 = (int)(v2 % W2), h2 = (int)((v2 / W2) % H2), d2 = (int)(v2 / ((long)W2 * H2));
        const float* gp = g + ((long)n * C + C1 + c) * V;
        float s = 0.f;
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dxx = 0; dxx < 2; ++dxx) s += gp[((long)(2 * d2 + dz) * H + 2 * h2 + dy) * W + 2 * w2 + dxx];
        dx2[j] = s;
    }
}

// --------------------------------------------------------------------------------------------- MaxPool3d(2) backward
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                                           long NC, int D, int H, int W) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    const long total = NC * Do * Ho * Wo;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long t = i;
        const int w = (int)(t % Wo); t /= Wo;
        const int h = (int)(t % Ho); t /= Ho;
        const int d = (int)(t % Do); t /= Do;
        const float* xp = x + t * (long)D * H * W;
        float* dp = dx + t * (long)D * H * W;
        float best = -FLT_MAX;
        int arg = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {  // scan order (d, h, w): the first maximum wins, NaN propagates like torch
            const long o = ((long)(2 * d + (q >> 2)) * H + 2 * h + ((q >> 1) & 1)) * W + 2 * w + (q & 1);
            const float v = xp[o];
            if (v > best || v != v) { best = v; arg = q; }
        }
        const float g = dy[i];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const long o = ((long)(2 * d + (q >> 2)) * H + 2 * h + ((q >> 1) & 1)) * W + 2 * w + (q & 1);
            dp[o] = q == arg ? g : 0.f;
        }
    }
}

// --------------------------------------------------------------------------------------------- attention epilogue
constexpr int FEAT = 32;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// One warp per row (b, r).  Recomputes the forward quantities of model/attention.py:84-113 from the saved inputs and
// propagates d(out) to the row's own vector, its K candidate vectors and the 32-d theta / phi features.
__global__ void __launch_bounds__(256) attention_epilogue_bwd_kernel(const float* __restrict__ xf, const float* __restrict__ pf,
                                                                     const float* __restrict__ xu, const float* __restrict__ pu,
                                                                     const float* __restrict__ noise, const float* __restrict__ dout,
                                                                     float* __restrict__ dxf, float* __restrict__ dpf,
                                                                     float* __restrict__ dxu, float* __restrict__ dpu, long R, int rp3,
                                                                     int K, int V, int normalize, int mode, int blend, float sharp) {
    const long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const long b = row / rp3, rr = row % rp3;
    const long prow0 = b * K * rp3 + rr;
    // ---- forward recompute: normalised features, scores (lane k holds s_k), switch, weights
    const float xraw = xf[row * FEAT + lane];
    const float xn = normalize ? fmaxf(sqrtf(wsum(xraw * xraw)), 1e-12f) : 1.f;
    const float xt = xraw / xn;
    float my_s = -FLT_MAX, my_pn = 1.f;
    for (int k = 0; k < K; ++k) {
        const float praw = pf[(prow0 + (long)k * rp3) * FEAT + lane];
        const float pn = normalize ? fmaxf(sqrtf(wsum(praw * praw)), 1e-12f) : 1.f;
        const float s = wsum(xt * (praw / pn));
        if (lane == k) { my_s = s; my_pn = pn; }
    }
    const float smax = wmax(my_s);
    const float sw = fmaxf(smax, 0.f);
    const unsigned mm = __ballot_sync(0xffffffffu, lane < K && my_s == smax);
    const int arg_s = __ffs(mm) - 1;  // MaxPool1d: first maximum
    float w = 0.f, ysoft = 0.f;
    const float scale = mode == 0 ? sharp : 25.f;
    {
        const float z = lane < K ? (scale * my_s + (mode == 0 ? 0.f : noise[row * K + lane])) : -FLT_MAX;
        const float zmax = wmax(z);
        const float e = lane < K ? expf(z - zmax) : 0.f;
        ysoft = e / wsum(e);
        if (mode == 0) w = ysoft;
        else {
            const float ymax = wmax(lane < K ? ysoft : -FLT_MAX);
            const unsigned m2 = __ballot_sync(0xffffffffu, lane < K && ysoft == ymax);
            w = lane < K ? (((lane == __ffs(m2) - 1) ? 1.f : 0.f) - ysoft) + ysoft : 0.f;
        }
    }
    // ---- pass over the vectors: a = sum_k w_k P_k, dswitch = <dout, a - x> (blend) or <dout, a>, dw_k = sw <dout, P_k>
    float dsw_part = 0.f, my_dw = 0.f;
    for (int v0 = 0; v0 < V; v0 += 32) {
        const int v = v0 + lane;
        const float go = v < V ? dout[row * V + v] : 0.f;
        const float xv = v < V ? xu[row * V + v] : 0.f;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            const float pv = v < V ? pu[(prow0 + (long)k * rp3) * V + v] : 0.f;
            const float wk = __shfl_sync(0xffffffffu, w, k);
            acc = fmaf(wk, pv, acc);
            const float dwk = wsum(go * pv) * sw;
            if (lane == k) my_dw += dwk;
            if (v < V) dpu[(prow0 + (long)k * rp3) * V + v] = wk * sw * go;
        }
        dsw_part += go * (blend ? acc - xv : acc);
        if (v < V) dxu[row * V + v] = blend ? go * (1.f - sw) : go;
    }
    const float dsw = wsum(dsw_part);
    // ---- scores: softmax (through the soft weights in Gumbel mode: straight-through), then the ReLU-max switch
    const float wd = wsum(lane < K ? ysoft * my_dw : 0.f);
    float ds = lane < K ? scale * ysoft * (my_dw - wd) : 0.f;
    if (smax > 0.f && lane == arg_s) ds += dsw;
    // ---- features: s_k = <x~, p~_k>
    float dxt = 0.f;
    for (int k = 0; k < K; ++k) {
        const float dsk = __shfl_sync(0xffffffffu, ds, k);
        const float pnk = __shfl_sync(0xffffffffu, my_pn, k);
        const float praw = pf[(prow0 + (long)k * rp3) * FEAT + lane];
        const float pt = praw / pnk;
        dxt = fmaf(dsk, pt, dxt);
        const float dpt = dsk * xt;  // d / d p~_k
        float dp = dpt;
        if (normalize) dp = (sqrtf(wsum(praw * praw)) > 1e-12f) ? (dpt - pt * wsum(pt * dpt)) / pnk : dpt / pnk;
        dpf[(prow0 + (long)k * rp3) * FEAT + lane] = dp;
    }
    float dx = dxt;
    if (normalize) dx = (sqrtf(wsum(xraw * xraw)) > 1e-12f) ? (dxt - xt * wsum(xt * dxt)) / xn : dxt / xn;
    dxf[row * FEAT + lane] = dx;
}

// forward epilogue on precomputed features (the differentiable path computes theta / phi layer by layer so that
// autograd keeps the activations): same arithmetic as attention_epilogue_kernel in rf_attention.cu
__global__ void __launch_bounds__(256) attention_epilogue_fwd_kernel(const float* __restrict__ xf, const float* __restrict__ pf,
                                                                     const float* __restrict__ xu, const float* __restrict__ pu,
                                                                     const float* __restrict__ noise, float* __restrict__ orows, long R,
                                                                     int rp3, int K, int V, int normalize, int mode, int blend,
                                                                     float sharp) {
    const long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const long b = row / rp3, rr = row % rp3;
    const long prow0 = b * K * rp3 + rr;
    float xv = xf[row * FEAT + lane];
    if (normalize) xv = xv / fmaxf(sqrtf(wsum(xv * xv)), 1e-12f);
    float my_s = -FLT_MAX;
    for (int k = 0; k < K; ++k) {
        float pv = pf[(prow0 + (long)k * rp3) * FEAT + lane];
        if (normalize) pv = pv / fmaxf(sqrtf(wsum(pv * pv)), 1e-12f);
        const float s = wsum(xv * pv);
        if (lane == k) my_s = s;
    }
    const float sw = fmaxf(wmax(my_s), 0.f);
    float w;
    {
        const float z = lane < K ? ((mode == 0 ? sharp : 25.f) * my_s + (mode == 0 ? 0.f : noise[row * K + lane])) : -FLT_MAX;
        const float zmax = wmax(z);
        const float e = lane < K ? expf(z - zmax) : 0.f;
        const float y = e / wsum(e);
        if (mode == 0) w = y;
        else {
            const float ymax = wmax(lane < K ? y : -FLT_MAX);
            const unsigned m2 = __ballot_sync(0xffffffffu, lane < K && y == ymax);
            w = lane < K ? (((lane == __ffs(m2) - 1) ? 1.f : 0.f) - y) + y : 0.f;
        }
    }
    for (int v = lane; v < V; v += 32) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(__shfl_sync(0xffffffffu, w, k), pu[(prow0 + (long)k * rp3) * V + v], acc);
        const float x = xu[row * V + v];
        orows[row * V + v] = blend ? (x * (1.f - sw) + acc * sw) : (x + acc * sw);
    }
}

}  // namespace

extern "C" int rf_conv3d_wgrad(const float* x, const float* x2, int C2, const float* gn_mu, const float* gn_a, const float* gn_beta,
                               const float* dz, float* dw, int N, int Cin, int Di, int Hi, int Wi, int Cout, int KS, int stride,
                               int pad, void* stream) {
    RF_CHECK_ARG(dz && dw && (x || C2 == Cin), "rf_conv3d_wgrad: null pointer");
    RF_CHECK_ARG(N > 0 && Cin > 0 && Cout > 0 && Di > 0 && Hi > 0 && Wi > 0 && KS >= 1 && KS <= 5 && stride >= 1 && pad >= 0,
                 "rf_conv3d_wgrad: bad shape");
    RF_CHECK_ARG(C2 >= 0 && C2 <= Cin && (C2 == 0 || x2), "rf_conv3d_wgrad: bad concat split");
    RF_CHECK_ARG((gn_mu == nullptr) == (gn_a == nullptr) && (gn_mu == nullptr) == (gn_beta == nullptr), "rf_conv3d_wgrad: partial GroupNorm arguments");
    WgradArgs a;
    a.x = x; a.x2 = x2; a.gn_mu = gn_mu; a.gn_a = gn_a; a.gn_beta = gn_beta; a.dz = dz; a.dw = dw;
    a.N = N; a.Cin = Cin; a.C2 = C2; a.C1 = Cin - C2; a.Di = Di; a.Hi = Hi; a.Wi = Wi; a.Cout = Cout; a.KS = KS; a.stride = stride; a.pad = pad;
    a.Do = (Di + 2 * pad - KS) / stride + 1; a.Ho = (Hi + 2 * pad - KS) / stride + 1; a.Wo = (Wi + 2 * pad - KS) / stride + 1;
    RF_CHECK_ARG(a.Do > 0 && a.Ho > 0 && a.Wo > 0, "rf_conv3d_wgrad: empty output");
    a.M = (long)N * a.Do * a.Ho * a.Wo;
    a.Kg = Cin * KS * KS * KS;
    const int gx = rf_cdiv(a.Kg, WK), gy = rf_cdiv(Cout, WN);
    long slices = (148L * 4 + (long)gx * gy - 1) / ((long)gx * gy);  // fill the chip, >= 256 positions per slice
    const long max_slices = (a.M + 255) / 256;
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    a.m_per_slice = (int)(((a.M + slices - 1) / slices + WM - 1) / WM * WM);
    const long gz = (a.M + a.m_per_slice - 1) / a.m_per_slice;
    RF_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)Cout * a.Kg * sizeof(float), (cudaStream_t)stream));
    conv_wgrad_kernel<<<dim3(gx, gy, (unsigned)gz), 256, 0, (cudaStream_t)stream>>>(a);
    RF_LAUNCH_OK("conv_wgrad_kernel");
    return 0;
}

extern "C" int rf_act_bwd(const float* dy, const float* y, float* dz, long n, int act, float slope, void* stream) {
    RF_CHECK_ARG(dy && y && dz && n > 0 && act >= RF_ACT_NONE && act <= RF_ACT_TANH, "rf_act_bwd: bad arguments");
    act_bwd_kernel<<<rf_grid_1d(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, dz, n, act, slope);
    RF_LAUNCH_OK("act_bwd_kernel");
    return 0;
}

extern "C" int rf_channel_sum(const float* x, int N, int C, long V, float* out, void* stream) {
    RF_CHECK_ARG(x && out && N > 0 && C > 0 && V > 0, "rf_channel_sum: bad arguments");
    RF_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)C * sizeof(float), (cudaStream_t)stream));
    long gy = ((long)N * V + 4095) / 4096;
    if (gy > 64) gy = 64;
    channel_sum_kernel<<<dim3(C, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(x, N, C, V, out);
    RF_LAUNCH_OK("channel_sum_kernel");
    return 0;
}

extern "C" size_t rf_gn_bwd_workspace_bytes(int N, int C) { return (size_t)N * C * (2 * sizeof(double) + 3 * sizeof(float)) + 256; }

extern "C" int rf_gn_bwd(const float* x, const float* x2, int C2, const float* g, const float* gn_mu, const float* gn_rstd,
                         const float* gamma, int N, int C, int D, int H, int W, int groups, float* dx, float* dx2, float* dgamma,
                         float* dbeta, void* workspace, void* stream) {
    RF_CHECK_ARG(g && gn_mu && gn_rstd && gamma && workspace && (x || C2 == C), "rf_gn_bwd: null pointer");
    RF_CHECK_ARG(N > 0 && C > 0 && groups > 0 && C % groups == 0 && C2 >= 0 && C2 <= C && (C2 == 0 || (x2 && dx2)) && (C2 == C || dx),
                 "rf_gn_bwd: bad arguments");
    RF_CHECK_ARG(C2 == 0 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "rf_gn_bwd: upsampled input needs even extents");
    cudaStream_t s = (cudaStream_t)stream;
    double* s12 = (double*)workspace;
    float* coef = (float*)(s12 + (size_t)N * C * 2);
    const int C1 = C - C2;
    gn_bwd_reduce_kernel<<<N * C, 256, 0, s>>>(x, x2, C1, C2, g, gn_mu, gn_rstd, D, H, W, s12);
    RF_LAUNCH_OK("gn_bwd_reduce_kernel");
    if (dgamma) RF_CUDA_OK(cudaMemsetAsync(dgamma, 0, (size_t)C * sizeof(float), s));
    if (dbeta) RF_CUDA_OK(cudaMemsetAsync(dbeta, 0, (size_t)C * sizeof(float), s));
    gn_bwd_coeff_kernel<<<rf_cdiv((long)N * C, 128), 128, 0, s>>>(s12, gamma, gn_rstd, N, C, groups, (double)D * H * W, coef, dgamma, dbeta);
    RF_LAUNCH_OK("gn_bwd_coeff_kernel");
    const long total = (long)N * C1 * D * H * W + (long)N * C2 * (D / 2) * (H / 2) * (W / 2);
    gn_bwd_apply_kernel<<<rf_grid_1d(total, 256), 256, 0, s>>>(x, x2, C1, C2, g, gn_mu, gn_rstd, coef, N, D, H, W, dx, dx2);
    RF_LAUNCH_OK("gn_bwd_apply_kernel");
    return 0;
}

extern "C" int rf_upsample2_bwd(const float* g, int N, int C, int C1, int D, int H, int W, float* dx2, void* stream) {
    RF_CHECK_ARG(g && dx2 && N > 0 && C > C1 && C1 >= 0 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "rf_upsample2_bwd: bad arguments");
    const long total = (long)N * (C - C1) * (D / 2) * (H / 2) * (W / 2);
    upsample2_bwd_kernel<<<rf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(g, N, C, C1, D, H, W, dx2);
    RF_LAUNCH_OK("upsample2_bwd_kernel");
    return 0;
}

extern "C" int rf_maxpool3d_2_bwd(const float* x, const float* dy, float* dx, int N, int C, int D, int H, int W, void* stream) {
    RF_CHECK_ARG(x && dy && dx && N > 0 && C > 0 && D >= 2 && H >= 2 && W >= 2 && D % 2 == 0 && H % 2 == 0 && W % 2 == 0,
                 "rf_maxpool3d_2_bwd: bad arguments (even extents only)");
    const long total = (long)N * C * (D / 2) * (H / 2) * (W / 2);
    maxpool2_bwd_kernel<<<rf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, (long)N * C, D, H, W);
    RF_LAUNCH_OK("maxpool2_bwd_kernel");
    return 0;
}

extern "C" int rf_attention_epilogue_fwd(const float* xf, const float* pf, const float* xu, const float* pu, const float* noise,
                                         float* orows, long R, int rp3, int K, int V, int normalize, int mode, int blend,
                                         float sharp, void* stream) {
    RF_CHECK_ARG(xf && pf && xu && pu && orows && R > 0 && rp3 > 0 && K >= 1 && K <= 32 && V > 0 && (mode == 0 || noise),
                 "rf_attention_epilogue_fwd: bad arguments");
    attention_epilogue_fwd_kernel<<<(unsigned)rf_cdivl(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(xf, pf, xu, pu, noise, orows, R, rp3, K,
                                                                                                    V, normalize, mode, blend, sharp);
    RF_LAUNCH_OK("attention_epilogue_fwd_kernel");
    return 0;
}

extern "C" int rf_attention_epilogue_bwd(const float* xf, const float* pf, const float* xu, const float* pu, const float* noise,
                                         const float* dout, float* dxf, float* dpf, float* dxu, float* dpu, long R, int rp3, int K,
                                         int V, int normalize, int mode, int blend, float sharp, void* stream) {
    RF_CHECK_ARG(xf && pf && xu && pu && dout && dxf && dpf && dxu && dpu && R > 0 && rp3 > 0 && K >= 1 && K <= 32 && V > 0 &&
                     (mode == 0 || noise), "rf_attention_epilogue_bwd: bad arguments");
    attention_epilogue_bwd_kernel<<<(unsigned)rf_cdivl(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        xf, pf, xu, pu, noise, dout, dxf, dpf, dxu, dpu, R, rp3, K, V, normalize, mode, blend, sharp);
    RF_LAUNCH_OK("attention_epilogue_bwd_kernel");
    return 0;
}
