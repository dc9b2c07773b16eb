// Callers either side of the hot path: the training-side contrastive loss and target normals,
// and the evaluation metrics.  All HBM / latency-bound small kernels:
//   * sobel_normals_kernel      dataset/patched_scene_dataset.py:139-146 compute_normals (pad with trunc, three 3x3x3
//                               Sobel cross-correlations :194-196, x / sqrt(|x|^2 + 1e-5))
//   * occupancy_counts_kernel   util/metrics.py:15-16,66,83 the integer sums behind IoU / Precision / Recall
//   * chamfer_nn_kernel         external/ChamferDistancePytorch chamfer3D (NmDistanceKernel): per point the squared
//                               distance to and index of its nearest neighbour in the other cloud, brute force
//   * ntxent_rows_kernel        model/loss.py:48-69 NTXentLoss.forward: cosine / dot similarities of the 2N stacked
//                               representations, temperature (or IoU-dependent temperature) scaling, cross-entropy
//                               against the positive pair - one warp per row, the 2N x 2N matrix is never materialised
#include <float.h>

#include "rf_common.cuh"

namespace {

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

// ---------------------------------------------------------------------------------------------------- normals
// One thread per voxel; the 27 neighbours come through L1 (each value is read by 27 threads of the same CTA
// neighbourhood).  Out-of-volume neighbours are the padding constant.
__global__ void __launch_bounds__(256) sobel_normals_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int D,
                                                            int H, int W, float pad_val) {
    const long vol = (long)D * H * W, total = (long)B * vol;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long t = i;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H); t /= H;
        const int d = (int)(t % D); t /= D;
        const float* xb = x + t * vol;
        float p[3][3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int dd = d + a - 1, hh = h + b - 1, ww = w + c - 1;
                    p[a][b][c] = (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xb + ((long)dd * H + hh) * W + ww) : pad_val;
                }
        const float s3[3] = {1.f, 2.f, 1.f};
        float dx = 0.f, dy = 0.f, dz = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // sobel_3d_x[a][b][c] = g[a] s[b] s[c], sobel_3d_y = s[a] g[b] s[c] with g = (+1, 0, -1);
                    // sobel_3d_z = s[a] s[b] (-g[c])   (patched_scene_dataset.py:194-196)
                    const float ga = (float)(1 - a), gb = (float)(1 - b), gc = (float)(c - 1);
                    dx = fmaf(ga * s3[b] * s3[c], p[a][b][c], dx);
                    dy = fmaf(s3[a] * gb * s3[c], p[a][b][c], dy);
                    dz = fmaf(s3[a] * s3[b] * gc, p[a][b][c], dz);
                }
        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz + 1e-5f);
        float* o = out + t * 3 * vol + ((long)d * H + h) * W + w;
        o[0] = dx / nrm;
        o[vol] = dy / nrm;
        o[2 * vol] = dz / nrm;
    }
}

// ---------------------------------------------------------------------------------------------------- occupancy counts
// counts[b] = {sum(p & t), sum(p | t), sum(p), sum(t)} over one sample's voxels (bool tensors = one byte per voxel).
__global__ void __launch_bounds__(256) occupancy_counts_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ tgt,
                                                               long vol, unsigned long long* __restrict__ counts) {
    const int b = blockIdx.y;
    const uint8_t* p = pred + (long)b * vol;
    const uint8_t* t = tgt + (long)b * vol;
    unsigned c_and = 0, c_or = 0, c_p = 0, c_t = 0;
    const bool vec = ((((uintptr_t)p) | ((uintptr_t)t)) & 15) == 0;
    const long nvec = vec ? vol / 16 : 0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p) + i), c = __ldg(reinterpret_cast<const uint4*>(t) + i);
        const unsigned av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // byte != 0 -> one bit per byte
            unsigned pa = av[j] | (av[j] >> 4); pa |= pa >> 2; pa |= pa >> 1; pa &= 0x01010101u;
            unsigned pc = cv[j] | (cv[j] >> 4); pc |= pc >> 2; pc |= pc >> 1; pc &= 0x01010101u;
            c_and += __popc(pa & pc); c_or += __popc(pa | pc); c_p += __popc(pa); c_t += __popc(pc);
        }
    }
    for (long i = nvec * 16 + blockIdx.x * (long)blockDim.x + threadIdx.x; i < vol; i += (long)gridDim.x * blockDim.x) {
        const unsigned pa = p[i] != 0, pc = t[i] != 0;
        c_and += pa & pc; c_or += pa | pc; c_p += pa; c_t += pc;
    }
    unsigned v[4] = {c_and, c_or, c_p, c_t};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
        if ((threadIdx.x & 31) == 0 && v[j]) atomicAdd(counts + 4 * b + j, (unsigned long long)v[j]);
    }
}

// ---------------------------------------------------------------------------------------------------- chamfer
// d(a, b) = fma(dz, dz, fma(dy, dy, dx * dx)) in fp32 (the contraction nvcc applies to the reference kernel's
// (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1)); the first minimal index wins.
constexpr int CH_TILE = 2048;
__global__ void __launch_bounds__(256) chamfer_nn_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb,
                                                         float* __restrict__ dist, int* __restrict__ idx) {
    __shared__ float sb[CH_TILE * 3];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (i < na) { ax = a[3 * i]; ay = a[3 * i + 1]; az = a[3 * i + 2]; }
    float best = FLT_MAX;
    int bi = 0;
    for (int j0 = 0; j0 < nb; j0 += CH_TILE) {
        const int n = min(CH_TILE, nb - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) sb[t] = b[3 * (long)j0 + t];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float dx = sb[3 * j] - ax, dy = sb[3 * j + 1] - ay, dz = sb[3 * j + 2] - az;
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            if (d < best) { best = d; bi = j0 + j; }
        }
    }
    if (i < na) { dist[i] = best; idx[i] = bi; }
}

// ---------------------------------------------------------------------------------------------------- NT-Xent
// rep = cat(zjs, zis) [2N, C].  Row i: logits over all j != i of sim(i, j) * inv_temp(i, j); the positive is
// j = (i + N) mod 2N (model/loss.py:53-56).  loss_i = logsumexp_j(logit) - logit_pos.  One warp per row, lanes over
// j with an online (max, sum) pair, the row's vector in registers (C <= 128).
__global__ void __launch_bounds__(128) ntxent_norms_kernel(const float* __restrict__ zis, const float* __restrict__ zjs, int N, int C,
                                                           float* __restrict__ norms) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= 2 * N) return;
    const float* r = row < N ? zjs + (long)row * C : zis + (long)(row - N) * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(r[c], r[c], s);
    s = wsum(s);
    if (lane == 0) norms[row] = sqrtf(s);
}

__global__ void __launch_bounds__(128) ntxent_rows_kernel(const float* __restrict__ zis, const float* __restrict__ zjs, int N, int C,
                                                          const float* __restrict__ norms, const float* __restrict__ iou,
                                                          float temperature, float sig_scale, float sig_shift, int cosine,
                                                          float* __restrict__ row_loss) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int M = 2 * N;
    if (row >= M) return;
    const float* ri = row < N ? zjs + (long)row * C : zis + (long)(row - N) * C;
    float xi[4];  // C <= 128: element c = lane + 32 t
#pragma unroll
    for (int t = 0; t < 4; ++t) xi[t] = (lane + 32 * t < C) ? ri[lane + 32 * t] : 0.f;
    const float ni = cosine ? norms[row] : 1.f;
    const int pos = row < N ? row + N : row - N;
    float m = -FLT_MAX, s = 0.f, lpos = 0.f;
    for (int j = 0; j < M; ++j) {
        if (j == row) continue;  // warp-uniform
        const float* rj = j < N ? zjs + (long)j * C : zis + (long)(j - N) * C;
        float dot = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (lane + 32 * t < C) dot = fmaf(xi[t], __ldg(rj + lane + 32 * t), dot);
        dot = wsum(dot);
        float sim = dot;
        if (cosine) sim = dot / fmaxf(ni * norms[j], 1e-8f);  // torch.nn.CosineSimilarity(eps = 1e-8)
        float logit;
        if (j == pos || !iou) {
            logit = sim / temperature;
        } else {  // model/loss.py:63-64
            const float z = iou[(long)row * M + j] * sig_scale + sig_shift;
            const float sg = 1.f / (1.f + expf(-z));
            logit = sim / (temperature + (1.f - temperature) * sg);
        }
        if (j == pos) lpos = logit;
        if (logit > m) { s = s * expf(m - logit) + 1.f; m = logit; }
        else s += expf(logit - m);
    }
    if (lane == 0) row_loss[row] = (m + logf(s)) - lpos;
}

// deterministic final sum (fixed order, fp64), loss = sum / 2N
__global__ void ntxent_sum_kernel(const float* __restrict__ row_loss, int M, float* __restrict__ out) {
    __shared__ double part[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) acc += (double)row_loss[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
        out[0] = (float)(t / (double)M);
    }
}

}  // namespace

extern "C" int rf_sobel_normals(const float* x, float* out, int B, int D, int H, int W, float pad_val, void* stream) {
    RF_CHECK_ARG(x && out && B > 0 && D > 0 && H > 0 && W > 0, "rf_sobel_normals: bad arguments");
    const long total = (long)B * D * H * W;
    sobel_normals_kernel<<<rf_grid_1d(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(x, out, B, D, H, W, pad_val);
    RF_LAUNCH_OK("sobel_normals_kernel");
    return 0;
}

extern "C" int rf_occupancy_counts(const uint8_t* pred, const uint8_t* target, int B, long voxels_per_sample, unsigned long long* counts,
                                   void* stream) {
    RF_CHECK_ARG(pred && target && counts && B > 0 && B <= 65535 && voxels_per_sample > 0, "rf_occupancy_counts: bad arguments");
    RF_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 4 * B, (cudaStream_t)stream));
    long gx = rf_cdivl(voxels_per_sample / 16 + 1, 256);  // CTAs per sample: ~148 x 8 in total
    const long cap = 148L * 8 / B > 1 ? 148L * 8 / B : 1;
    if (gx > cap) gx = cap;
    occupancy_counts_kernel<<<dim3((unsigned)gx, B), 256, 0, (cudaStream_t)stream>>>(pred, target, voxels_per_sample, counts);
    RF_LAUNCH_OK("occupancy_counts_kernel");
    return 0;
}

extern "C" int rf_chamfer_nn(const float* a, int na, const float* b, int nb, float* dist, int* idx, void* stream) {
    RF_CHECK_ARG(a && b && dist && idx && na > 0 && nb > 0, "rf_chamfer_nn: bad arguments (empty clouds are the caller's case)");
    chamfer_nn_kernel<<<(unsigned)rf_cdivl(na, 256), 256, 0, (cudaStream_t)stream>>>(a, na, b, nb, dist, idx);
    RF_LAUNCH_OK("chamfer_nn_kernel");
    return 0;
}

extern "C" size_t rf_ntxent_workspace_bytes(int N) { return N > 0 ? (size_t)4 * N * sizeof(float) : 0; }

extern "C" int rf_ntxent_fwd(const float* zis, const float* zjs, int N, int C, const float* iou_matrix, float temperature,
                             float sig_scale, float sig_shift, int cosine, float* loss, void* workspace, size_t workspace_bytes,
                             void* stream) {
    RF_CHECK_ARG(zis && zjs && loss && workspace && N > 0, "rf_ntxent_fwd: bad arguments");
    RF_CHECK_ARG(C > 0 && C <= 128, "rf_ntxent_fwd: feature width %d unsupported (1..128)", C);
    RF_CHECK_ARG(workspace_bytes >= rf_ntxent_workspace_bytes(N), "rf_ntxent_fwd: workspace too small");
    RF_CHECK_ARG(temperature > 0.f, "rf_ntxent_fwd: temperature must be positive");
    float* norms = (float*)workspace;
    float* row_loss = norms + 2 * (size_t)N;
    const unsigned grid = (unsigned)rf_cdivl(2L * N * 32, 128);
    if (cosine) {
        ntxent_norms_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(zis, zjs, N, C, norms);
        RF_LAUNCH_OK("ntxent_norms_kernel");
    }
    ntxent_rows_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(zis, zjs, N, C, norms, iou_matrix, temperature, sig_scale, sig_shift,
                                                              cosine, row_loss);
    RF_LAUNCH_OK("ntxent_rows_kernel");
    ntxent_sum_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(row_loss, 2 * N, loss);
    RF_LAUNCH_OK("ntxent_sum_kernel");
    return 0;
}
