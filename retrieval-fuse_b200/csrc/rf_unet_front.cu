// Fused front of the retrieval U-Net's first DoubleConv on 16^3 patches (model/unet.py:79-144, layer order 'gcr'):
//
//     GroupNorm(1 group, 1 channel) -> Conv3d(1, 8, 3, padding=1, bias=False) -> ReLU       (SingleConv1)
//     -> GroupNorm(G groups, 8 channels) of that output                                      (front of SingleConv2)
//     -> x16, fp16 hi / lo split -> operand planes of the shifted-window convolution (rf_tc_conv_halo.cu)
//
// As separate launches (statistics of the input, fp32 FMA convolution, statistics of its output, normalise + split)
// this front moved the 8-channel fp32 activations of every patch through HBM three times (2.1 GB each at 16 384
// patches: written by the convolution, read by the statistics, read by the split): 3.1 ms of a 39 ms step.  Here one
// persistent CTA owns a whole 16^3 sample: the normalised input sits in shared memory with its zero halo, every thread
// keeps its 8 x 8 outputs in registers while the CTA reduces the per-channel sums (fp64), and the normalised, split
// outputs go straight to the operand planes - the fp32 activations never exist in HBM.
//
// Numerics: same operations as the separate kernels (fmaf((x - mu), a, beta) normalisation, fp32 FMA convolution with
// taps in (kd, kh, kw) order, mean / var / rstd as in cl_gn_finalize_kernel from fp64 sums; a thread's 8 outputs per
// channel are pre-summed in fp32).
#include "rf_tc_common.cuh"

namespace {
using namespace rf_tc;

constexpr int S = 16, SP = S + 2, VOX = S * S * S, NT = 512, CO = 8;

struct FrontArgs {
    const float* x;        // [N,16,16,16] (channels-last with C = 1)
    const float* gamma2;   // [8] GroupNorm weight / bias of SingleConv2
    const float* beta2;
    uint4 *hi, *lo;        // operand planes: [n][d][h][w] slots, or W-pair planes [w parity][n][d][h][w / 2]
    float gamma1, beta1, eps1, eps2, scale;
    int N, wp, cpg;        // cpg: channels per group of the second GroupNorm (1 or 8)
    float w[27][CO];       // Conv3d weight, tap-major: the FMAs read it as constant-bank operands
};

// sum over the warp of 8 doubles per lane: recursive halving (9 shuffled values instead of 40), lanes 4 i .. 4 i + 3
// end up with the total of value i
__device__ __forceinline__ double warp_reduce8(double (&v)[8], int lane) {
#pragma unroll
    for (int half = 4, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const double send = up ? v[k] : v[k + half], keep = up ? v[k + half] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    double t = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 2);
    return t + __shfl_xor_sync(0xffffffffu, t, 1);
}

__global__ void __launch_bounds__(NT, 1) unet_front16_kernel(const __grid_constant__ FrontArgs a) {
    __shared__ float xs[SP * SP * SP];  // normalised input with its zero halo
    __shared__ double red[NT / 32][16];
    __shared__ float st1[2];
    __shared__ __align__(16) float st2[3][CO];  // mu, a, beta of the second GroupNorm for this sample
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = tid & 15, h = (tid >> 4) & 15, d0 = (tid >> 8) * 8;  // thread <-> the 8 voxels (d0 .. d0 + 7, h, w)
    for (int i = tid; i < SP * SP * SP; i += NT) xs[i] = 0.f;
    if (tid < CO) st2[2][tid] = __ldg(a.beta2 + tid);
    int n = blockIdx.x;
    float4 v0 = make_float4(0, 0, 0, 0), v1 = v0;
    if (n < a.N) {
        const float4* xg = reinterpret_cast<const float4*>(a.x + (long)n * VOX);
        v0 = __ldg(xg + tid);
        v1 = __ldg(xg + tid + NT);
    }
    __syncthreads();
    for (; n < a.N; n += gridDim.x) {
        // ---- statistics of the input sample (GroupNorm over its single channel)
        {
            const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            double s = 0.0, q = 0.0;
#pragma unroll
            for (int e = 0; e < 8; ++e) { s += (double)f[e]; q += (double)f[e] * (double)f[e]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
            if (lane == 0) { red[warp][0] = s; red[warp][1] = q; }
        }
        __syncthreads();
        if (warp == 0) {
            double s = lane < NT / 32 ? red[lane][0] : 0.0, q = lane < NT / 32 ? red[lane][1] : 0.0;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
            if (lane == 0) {
                const double mean = s / (double)VOX;
                double var = q / (double)VOX - mean * mean;
                if (var < 0.0) var = 0.0;
                st1[0] = (float)mean;
                st1[1] = (float)(1.0 / sqrt(var + (double)a.eps1)) * a.gamma1;
            }
        }
        __syncthreads();
        {
            const float mu = st1[0], ga = st1[1];
            const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = (e < 4 ? 4 * tid : 4 * (tid + NT)) + (e & 3);  // voxel (d, h, w) = (i >> 8, (i >> 4) & 15, i & 15)
                xs[(((i >> 8) + 1) * SP + ((i >> 4) & 15) + 1) * SP + (i & 15) + 1] = fmaf(f[e] - mu, ga, a.beta1);
            }
        }
        // the next sample's loads fly during the convolution
        if (n + (int)gridDim.x < a.N) {
            const float4* xg = reinterpret_cast<const float4*>(a.x + (long)(n + gridDim.x) * VOX);
            v0 = __ldg(xg + tid);
            v1 = __ldg(xg + tid + NT);
        }
        __syncthreads();
        // ---- Conv3d(1, 8, 3, padding 1) + ReLU: 8 voxels along d x 8 channels per thread
        float acc[8][CO];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float* col = xs + ((d0 + kd) * SP + h + kh) * SP + w + kw;
                    float xv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) xv[j] = col[j * SP * SP];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int c = 0; c < CO; ++c) acc[j][c] = fmaf(xv[j], a.w[(kd * 3 + kh) * 3 + kw][c], acc[j][c]);
                }
        // ---- statistics of the ReLU outputs per channel (fp64 sums), CTA-wide
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[j][c] = fmaxf(acc[j][c], 0.f);
        // (per thread the 8 values of a channel are summed in fp32 - 2^-24 relative per partial sum, averaged over 512
        // threads -, everything across threads in fp64: 64 live accumulators leave no room for fp64 partials)
#pragma unroll
        for (int r = 0; r < 2; ++r) {  // round 0: the sums of the 8 channels -> red[warp][0..7]; round 1: the sums of squares -> [8..15]
            float p[CO];
#pragma unroll
            for (int c = 0; c < CO; ++c) p[c] = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int c = 0; c < CO; ++c) p[c] = r ? fmaf(acc[j][c], acc[j][c], p[c]) : p[c] + acc[j][c];
            double v[8];
#pragma unroll
            for (int c = 0; c < CO; ++c) v[c] = (double)p[c];
            const double tot = warp_reduce8(v, lane);
            if (!(lane & 3)) red[warp][r * 8 + (lane >> 2)] = tot;
        }
        __syncthreads();
        if (warp == 0) {
            // lane i < 16: total of value i over the 16 warps; then the group sums
            double t = 0.0;
            if (lane < 16) {
#pragma unroll
                for (int k = 0; k < NT / 32; ++k) t += red[k][lane];
            }
            double s = t, q = __shfl_sync(0xffffffffu, t, (lane & 7) + 8);
            if (a.cpg == CO) {  // one group over the 8 channels
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
            }
            if (lane < CO) {
                const double cnt = (double)VOX * a.cpg;
                const double mean = s / cnt;
                double var = q / cnt - mean * mean;
                if (var < 0.0) var = 0.0;
                st2[0][lane] = (float)mean;
                st2[1][lane] = (float)(1.0 / sqrt(var + (double)a.eps2)) * __ldg(a.gamma2 + lane);
            }
        }
        __syncthreads();
        // ---- normalise, x scale, split into fp16 hi / lo, store one slot (8 channels) per voxel and plane
        {
            const float(&m)[CO] = st2[0];  // (broadcast reads from shared memory: 24 more live registers would spill)
            const float(&sa)[CO] = st2[1];
            const float(&sb)[CO] = st2[2];
            const long base = a.wp ? ((long)(w & 1) * a.N + n) * (VOX / 2) + (w >> 1) : (long)n * VOX + w;
            const int ws = a.wp ? S / 2 : S;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t hh[4], ll[4];
#pragma unroll
                for (int c = 0; c < CO; c += 2) {
                    const float f0 = fmaf(acc[j][c] - m[c], sa[c], sb[c]) * a.scale;
                    const float f1 = fmaf(acc[j][c + 1] - m[c + 1], sa[c + 1], sb[c + 1]) * a.scale;
                    split_f16x2(f0, f1, hh[c >> 1], ll[c >> 1]);
                }
                const long i = base + (long)((d0 + j) * S + h) * ws;
                a.hi[i] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                a.lo[i] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
            }
        }
        // (the next iteration's first barrier orders these reads of st2 / red before they are rewritten)
    }
}
}  // namespace

/* Fused front of the first DoubleConv of a 'gcr' U-Net encoder on 16^3 single-channel samples (model/unet.py:79-144):
 * GroupNorm(1,1) -> Conv3d(1,8,3,p=1) -> ReLU -> GroupNorm(groups2, 8) -> x scale -> fp16 hi / lo operand planes of
 * rf_tc_conv3d_halo_fwd (wp = 0: rf_cl_norm_split_halo's layout) or rf_tc_conv3d_halo_wp_fwd (wp = 1).
 * x [N,16,16,16] and gn2_w / gn2_b [8] on the device; the 218 scalars the FMAs read as kernel-parameter constants come
 * from the HOST: conv_w_host [8][27] (= Conv3d weight [8,1,3,3,3]), gn1_w / gn1_b by value.  No stream synchronisation,
 * so the call can be captured into a CUDA graph (the graph then holds these values: re-capture after a weight update).
 * hi / lo: rf_halo_act_bytes(N,16,16,16,8,0,1) bytes each. */
extern "C" int rf_unet_front16_fwd(const float* x, float gn1_w, float gn1_b, float eps1, const float* conv_w_host, const float* gn2_w,
                                   const float* gn2_b, int groups2, float eps2, float scale, void* hi, void* lo, int N, int wp,
                                   void* stream) {
    RF_CHECK_ARG(x && conv_w_host && gn2_w && gn2_b && hi && lo && N > 0, "rf_unet_front16_fwd: bad arguments");
    RF_CHECK_ARG(groups2 == 1 || groups2 == CO, "rf_unet_front16_fwd: the second GroupNorm must have 1 or 8 groups");
    RF_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0, "rf_unet_front16_fwd: pointers must be 16-byte aligned");
    FrontArgs a;
    a.x = x; a.gamma2 = gn2_w; a.beta2 = gn2_b; a.hi = (uint4*)hi; a.lo = (uint4*)lo;
    a.eps1 = eps1; a.eps2 = eps2; a.scale = scale; a.N = N; a.wp = wp ? 1 : 0; a.cpg = CO / groups2;
    for (int t = 0; t < 27; ++t)
        for (int c = 0; c < CO; ++c) a.w[t][c] = conv_w_host[c * 27 + t];
    a.gamma1 = gn1_w; a.beta1 = gn1_b;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unet_front16_kernel<<<N < sms ? N : sms, NT, 0, (cudaStream_t)stream>>>(a);
    RF_LAUNCH_OK("unet_front16_kernel");
    return 0;
}
