// Patch attention (model/attention.py:49-157) on the GPU.
//
// Rows: one row per (chunk b, sub-patch r) with r in [0, (S/E)^3); a row's
// vector is the nf*E^3 values of that sub-patch in (c, ex, ey, ez) order -
// exactly Unfold3D(E, nf)'s row layout.  The K retrieved candidates of a row
// are kept in (b, k, r) order (the order Unfold3D produces on x_retr), the
// reference's permute to (b, r, k) is only an indexing change.
//
// Stage plan: unfold -> theta / phi MLPs (one fused tcgen05 chain each, rf_tc_mlp.cu; fp32 FMA layers as the fallback)
// -> one warp-per-row score / blend stage (normalise, scores, ReLU-max switch, softmax(1024 s) or hard Gumbel
// arg-max, weighted sum of the raw candidate vectors, blend) -> fold.
// rf_attention_fuse_patched_fwd (round 2) drops the re-indexing passes around it: the candidates may arrive as the
// retrieval U-Net's un-folded patches (rows in (patch, local) order, mapped by index arithmetic), the result may be
// stored as the channels-last volume the decoder reads, and the g / o 1x1x1 convolutions of
// attn_no_output_mapping=False are one composed channel-mixing pass over the weighted-sum rows.
#include <float.h>

#include "rf_common.cuh"

namespace {

constexpr int FEAT = 32;    // cf_feat, model/attention.py:54
constexpr int HIDDEN = 128; // model/attention.py:35-41

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Row geometry of the epilogue.  P > 1: the candidates' rows (pf, pu) are in (b, k, patch, local sub-patch) order - the
// retrieval U-Net's P^3 patches per volume were unfolded without folding them into the volume first (Fold3D then
// Unfold3D is a permutation of rows).  out_cl: the blended rows are stored as the channels-last volume [B,S,S,S,nf]
// (what the decoder's first convolution reads) instead of rows [R,V] that a Fold3D would have to re-index.
struct AttnGeo {
    int P, out_cl, nf, E, Rp;
    FastDiv d_nf, d_E, d_Rp, d_ps;  // ps = Rp / P sub-patches per patch side
    int pow2, s_nf, s_E, s_Rp, s_ps, s_P;  // all five are powers of two (nf = 16, E = 2, 16^3 sub-patches in 4^3 patches): shifts
    float* side;  // != NULL (g / o output mapping follows): store the un-blended weighted sum and (switch, sum of weights) per row
};

// Unfold3D(2, C) of patches [NP, C, 8,8,8] -> rows [NP * 64, C * 8]: one CTA per patch.  Both sides of a patch are ONE
// contiguous run of C * 512 floats, the (c, x, y, z) -> (lx, ly, lz, c, ex, ey, ez) transposition happens in shared
// memory (x stride 66, c stride 532 floats: the float2 reads of a half-warp hit 16 different bank pairs).
__global__ void __launch_bounds__(256) unfold_e2_patch8_kernel(const float* __restrict__ in, float* __restrict__ out, int C) {
    extern __shared__ float sm_patch[];
    const float4* pin = reinterpret_cast<const float4*>(in + (long)blockIdx.x * C * 512);
    float4* pout = reinterpret_cast<float4*>(out + (long)blockIdx.x * C * 512);
    for (int i = threadIdx.x; i < C * 128; i += blockDim.x) {
        const float4 v = __ldg(pin + i);
        const int c = i >> 7, rem = i & 127, x = rem >> 4, yz = (rem & 15) << 2;
        float* d = sm_patch + c * 532 + x * 66 + yz;
        *reinterpret_cast<float2*>(d) = make_float2(v.x, v.y);
        *reinterpret_cast<float2*>(d + 2) = make_float2(v.z, v.w);
    }
    __syncthreads();
    const int C2 = 2 * C;
    for (int o = threadIdx.x; o < 64 * C2; o += blockDim.x) {
        const int row = o / C2, j = o - row * C2;
        const int lx = row >> 4, ly = (row >> 2) & 3, lz = row & 3, c = j >> 1, ex = j & 1;
        const float* s0 = sm_patch + c * 532 + (2 * lx + ex) * 66 + 16 * ly + 2 * lz;
        const float2 a = *reinterpret_cast<const float2*>(s0);
        const float2 bq = *reinterpret_cast<const float2*>(s0 + 8);
        pout[o] = make_float4(a.x, a.y, bq.x, bq.y);
    }
}

// xf [R,32], pf [(b,k,r),32], xu [R,V], pu [(b,k,r),V] -> orows [R,V]
// One warp per row.  KT = compile-time bound on K: the K feature rows and (for KT <= 8) the first 128 values of the
// row's own and of its K candidate vectors are requested BEFORE the score / softmax chain, so that one warp has all
// of its ~(K + 1) * 5 loads in flight at once (with a run-time K loop the loads were serialised behind the warp
// reductions: 1.35 TB/s on the 113 MB of this stage).
// kExact: K == KT and V == 128 (nf = 16, E = 2: the shipped configurations) - every k < K / v < V guard and the row
// pitch fold at compile time.  The general instantiation executed 1 086 instructions per row, half of them integer
// address arithmetic and guards, and ran at 77 % issue utilisation with DRAM at 36 % (profiles/r02s4_attention_64chunks).
template <int KT, bool kExact>
__global__ void __launch_bounds__(256) attention_epilogue_kernel(const float* __restrict__ xf, const float* __restrict__ pf,
                                                                 const float* __restrict__ xu, const float* __restrict__ pu,
                                                                 const float* __restrict__ noise, float* __restrict__ orows,
                                                                 long R, int rp3, int K_, int V_, int normalize, int mode,
                                                                 int blend, float sharp, const AttnGeo g) {
    const int K = kExact ? KT : K_, V = kExact ? 128 : V_;
    const long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const long b = row / rp3;
    unsigned rr = (unsigned)(row % rp3);
    unsigned px = 0, py = 0, pz = 0;
    // (the index arithmetic of a row is of the order of its ~300 other instructions: shifts where every extent is a
    // power of two, multiply-high division otherwise)
    if (g.P > 1 || g.out_cl) {
        if (g.pow2) {
            pz = rr & (g.Rp - 1);
            py = (rr >> g.s_Rp) & (g.Rp - 1);
            px = rr >> (2 * g.s_Rp);
        } else {
            unsigned t = rr;
            pz = fd_divmod(t, g.d_Rp);
            py = fd_divmod(t, g.d_Rp);
            px = t;
        }
    }
    if (g.P > 1) {
        // candidate rows come from Unfold3D run on the P^3 un-folded patches of each volume: (patch, local sub-patch) order
        const unsigned ps = g.d_ps.d;
        if (g.pow2) {
            const unsigned m = ps - 1;
            rr = ((((px >> g.s_ps) << g.s_P | (py >> g.s_ps)) << g.s_P | (pz >> g.s_ps)) << (3 * g.s_ps)) |
                 ((((px & m) << g.s_ps) | (py & m)) << g.s_ps) | (pz & m);
        } else {
            const unsigned qx = fd_div(px, g.d_ps), qy = fd_div(py, g.d_ps), qz = fd_div(pz, g.d_ps);
            rr = ((qx * g.P + qy) * g.P + qz) * (ps * ps * ps) + ((px - qx * ps) * ps + (py - qy * ps)) * ps + (pz - qz * ps);
        }
    }
    const long prow0 = b * K * rp3 + rr;  // candidate k lives at prow0 + k * rp3
    // channels-last output: the warp's i-th value is element (e, c) = (i / nf, i % nf) of the sub-patch, i.e. feature
    // c * E^3 + e of the row; its 2 * nf (E = 2) consecutive values are one contiguous piece of out [B,S,S,S,nf]
    const int E3 = g.E * g.E * g.E;
    const long S = (long)g.Rp * g.E;
    auto feat_of = [&](int i) -> int {
        if (!g.out_cl) return i;
        if (g.pow2) return ((i & (g.nf - 1)) << (3 * g.s_E)) | (i >> g.s_nf);
        const unsigned e = fd_div((unsigned)i, g.d_nf);
        return (int)(((unsigned)i - e * g.nf) * E3 + e);
    };
    const long out_base = g.out_cl ? (((b * S + px * g.E) * S + py * g.E) * S + pz * g.E) * g.nf : row * (long)V;
    auto out_addr = [&](int i) -> long {
        if (!g.out_cl) return out_base + i;
        unsigned e, c, ex, ey, ez;
        if (g.pow2) {
            e = (unsigned)i >> g.s_nf; c = (unsigned)i & (g.nf - 1);
            ez = e & (g.E - 1); ey = (e >> g.s_E) & (g.E - 1); ex = e >> (2 * g.s_E);
        } else {
            e = fd_div((unsigned)i, g.d_nf);
            c = (unsigned)i - e * g.nf;
            ez = fd_divmod(e, g.d_E); ey = fd_divmod(e, g.d_E); ex = e;
        }
        return out_base + ((ex * S + ey) * S + ez) * g.nf + c;
    };
    constexpr bool kPrefetch = KT <= 8;
    constexpr int PF = kPrefetch ? KT : 1;
    // channels-last store of a row that fits the prefetch registers: the loads stay in feature order (one 512-byte run
    // per vector; reading them in (e, c) order touched every sector of the run four times) and the blended row is
    // transposed through a per-warp shared-memory line (xor swizzle: conflict-free writes, two-way reads at nf = 16)
    const bool via_smem = kPrefetch && g.out_cl && V <= 128;
    __shared__ float tr_line[8][128];
    float* tr = tr_line[threadIdx.x >> 5];
    auto swz = [](int v) { return v ^ ((v >> 5) & 7); };

    float xv = __ldg(xf + row * FEAT + lane);
    float pvk[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) pvk[k] = k < K ? __ldg(pf + (prow0 + (long)k * rp3) * FEAT + lane) : 0.f;
    const float* xrow = xu + row * V;
    float xr[4], pr[PF][4];
    if (kPrefetch) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = lane + 32 * j;
            const int v = i < V ? (via_smem ? i : feat_of(i)) : 0;
            xr[j] = i < V ? __ldg(xrow + v) : 0.f;
#pragma unroll
            for (int k = 0; k < PF; ++k) pr[k][j] = (k < K && i < V) ? __ldg(pu + (prow0 + (long)k * rp3) * V + v) : 0.f;
        }
    }

    if (normalize) {
        const float n = fmaxf(sqrtf(warp_sum(xv * xv)), 1e-12f);  // F.normalize eps
        xv = xv / n;
    }
    float my_s = -FLT_MAX;  // lane k holds score k
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        if (k < K) {
            float pv = pvk[k];
            if (normalize) {
                const float n = fmaxf(sqrtf(warp_sum(pv * pv)), 1e-12f);
                pv = pv / n;
            }
            const float s = warp_sum(xv * pv);
            if (lane == k) my_s = s;
        }
    }
    const float smax = warp_max(my_s);
    const float sw = fmaxf(smax, 0.f);  // relu(max_k s), model/attention.py:99
    float w = 0.f;                      // lane k holds weight k
    if (mode == 0) {
        const float z = lane < K ? sharp * my_s : -FLT_MAX;
        const float zmax = warp_max(z);
        const float e = lane < K ? expf(z - zmax) : 0.f;
        const float den = warp_sum(e);
        w = e / den;
    } else {
        // gumbel_softmax(25 s, tau=1, hard=True): y_hard - y_soft + y_soft
        const float z = lane < K ? (25.f * my_s + noise[row * K + lane]) : -FLT_MAX;
        const float zmax = warp_max(z);
        const float e = lane < K ? expf(z - zmax) : 0.f;
        const float den = warp_sum(e);
        const float y = e / den;
        // arg-max of y_soft, first index on ties
        const float ymax = warp_max(lane < K ? y : -FLT_MAX);
        const unsigned mm = __ballot_sync(0xffffffffu, lane < K && y == ymax);
        const int arg = __ffs(mm) - 1;
        const float hard = lane == arg ? 1.f : 0.f;
        w = lane < K ? (hard - y) + y : 0.f;
    }
    float wk[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) wk[k] = __shfl_sync(0xffffffffu, w, k);
    const bool raw = g.side != nullptr;
    if (raw) {
        const float sumw = warp_sum(w);
        if (lane == 0) *reinterpret_cast<float2*>(g.side + 2 * row) = make_float2(sw, sumw);
    }
    int v0 = 0;
    if (kPrefetch) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = lane + 32 * j;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < PF; ++k)
                if (k < K) acc = fmaf(wk[k], pr[k][j], acc);
            const float o = raw ? acc : blend ? (xr[j] * (1.f - sw) + acc * sw) : (xr[j] + acc * sw);
            if (via_smem) tr[swz(v)] = o;
            else if (v < V) orows[out_addr(v)] = o;
        }
        if (via_smem) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = lane + 32 * j;
                if (i < V) orows[out_addr(i)] = tr[swz(feat_of(i))];
            }
        }
        v0 = 128;
    }
    for (int i = v0 + lane; i < V; i += 32) {
        const int v = feat_of(i);
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K) acc = fmaf(wk[k], __ldg(pu + (prow0 + (long)k * rp3) * V + v), acc);
        const float x = xrow[v];
        orows[out_addr(i)] = raw ? acc : blend ? (x * (1.f - sw) + acc * sw) : (x + acc * sw);
    }
}

// model/attention.py:56-57,95,108-109 with attn_no_output_mapping = False: g and o are 1x1x1 convolutions (channel mixing
// + bias) around the weighted sum.  Both are linear and act per voxel, so o(sum_k w_k g(p_k)) =
// (Wo Wg) sum_k w_k p_k + (Wo bg) sum_k w_k + bo: the epilogue above leaves the plain weighted sum in orows and
// (switch, sum_k w_k) in side; this kernel mixes the channels of each row (one warp per row, the row staged in shared
// memory) and blends.  mix = Wo Wg [nf, nf] row-major, mix_bg = Wo bg [nf], bo [nf].
__global__ void __launch_bounds__(256) attention_output_mapping_kernel(float* __restrict__ orows, const float* __restrict__ xu,
                                                                       const float* __restrict__ side, const float* __restrict__ mix,
                                                                       const float* __restrict__ mix_bg, const float* __restrict__ bo,
                                                                       long R, int V, int nf, int E3, int blend) {
    extern __shared__ float map_lines[];
    const long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    float* line = map_lines + (threadIdx.x >> 5) * V;
    for (int v = lane; v < V; v += 32) line[v] = orows[row * V + v];
    __syncwarp();
    const float2 ss = *reinterpret_cast<const float2*>(side + 2 * row);
    const float sw = ss.x, sumw = ss.y;
    for (int v = lane; v < V; v += 32) {
        const int co = v / E3, e = v - co * E3;
        float acc = fmaf(__ldg(mix_bg + co), sumw, __ldg(bo + co));
        for (int c = 0; c < nf; ++c) acc = fmaf(__ldg(mix + co * nf + c), line[c * E3 + e], acc);
        const float x = __ldg(xu + row * V + v);
        orows[row * V + v] = blend ? (x * (1.f - sw) + acc * sw) : (x + acc * sw);
    }
}

__global__ void __launch_bounds__(256) occ_any_kernel(const uint8_t* __restrict__ occ, uint8_t* __restrict__ out, int B, int S,
                                                      int E) {
    const int Rp = S / E;
    const long total = (long)B * Rp * Rp * Rp;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long t = i;
        const int pz = (int)(t % Rp); t /= Rp;
        const int py = (int)(t % Rp); t /= Rp;
        const int px = (int)(t % Rp); t /= Rp;
        uint8_t any = 0;
        for (int ex = 0; ex < E; ++ex)
            for (int ey = 0; ey < E; ++ey)
                for (int ez = 0; ez < E; ++ez)
                    any |= occ[((t * S + px * E + ex) * S + py * E + ey) * (long)S + pz * E + ez] != 0;
        out[i] = any;
    }
}

struct AttnWs {
    float *xu, *pu, *ha, *hb, *xf, *pf, *orows;
};

size_t attn_ws_layout(int B, int nf, int S, int E, int K, AttnWs* ws, char* base) {
    const size_t Rp = S / E, R = (size_t)B * Rp * Rp * Rp, V = (size_t)nf * E * E * E;
    const size_t Kc = K > 1 ? K : 1;
    size_t off = 0;
    auto take = [&](size_t n_floats) {
        float* p = (float*)(base + off);
        off += (n_floats * sizeof(float) + 255) / 256 * 256;
        return p;
    };
    float* xu = take(R * V);
    float* pu = take(R * Kc * V);
    float* ha = take(R * Kc * HIDDEN);
    float* hb = take(R * Kc * HIDDEN);
    float* xf = take(R * FEAT);
    float* pf = take(R * Kc * FEAT);
    float* orows = take(R * V);
    if (ws) { ws->xu = xu; ws->pu = pu; ws->ha = ha; ws->hb = hb; ws->xf = xf; ws->pf = pf; ws->orows = orows; }
    return off;
}

// theta / phi: Linear(V,128) LeakyReLU Linear(128,128) LeakyReLU Linear(128,128) LeakyReLU Linear(128,32).
// img != NULL: the fused tcgen05 chain (rf_tc_mlp_fwd) on rf_tc_mlp_weight_image operand images; else fp32 FMA kernels.
int run_mlp(const float* in, long rows, int V, const float* const* wt, const float* const* b, const void* const* img,
            float* ha, float* hb, float* out, void* stream) {
    RF_CHECK_ARG(rows < (1L << 31), "attention: too many rows");
    const float slope = 0.01f;  // nn.LeakyReLU() default
    int rc;
    if (img) {  // all four layers in one launch, hidden activations on chip (rf_tc_mlp.cu)
        const int widths[5] = {V, HIDDEN, HIDDEN, HIDDEN, FEAT};
        return rf_tc_mlp_fwd(in, V, img, b, widths, 4, RF_ACT_LEAKY, slope, 0, 0.f, out, FEAT, rows, stream);
    }
    if ((rc = rf_linear_fwd(in, wt[0], b[0], ha, (int)rows, V, HIDDEN, RF_ACT_LEAKY, slope, stream))) return rc;
    if ((rc = rf_linear_fwd(ha, wt[1], b[1], hb, (int)rows, HIDDEN, HIDDEN, RF_ACT_LEAKY, slope, stream))) return rc;
    if ((rc = rf_linear_fwd(hb, wt[2], b[2], ha, (int)rows, HIDDEN, HIDDEN, RF_ACT_LEAKY, slope, stream))) return rc;
    return rf_linear_fwd(ha, wt[3], b[3], out, (int)rows, HIDDEN, FEAT, RF_ACT_NONE, 0.f, stream);
}

}  // namespace

extern "C" size_t rf_attention_workspace_bytes(int B, int nf, int S, int E, int K) {
    if (B <= 0 || nf <= 0 || S <= 0 || E <= 0 || K <= 0 || S % E) return 0;
    return attn_ws_layout(B, nf, S, E, K, nullptr, nullptr);
}

extern "C" int rf_attention_fuse_patched_fwd(const float* x_back, const float* x_retr, const float* const* theta_wt_host,
                                             const float* const* theta_b_host, const float* const* phi_wt_host,
                                             const float* const* phi_b_host, const void* const* theta_img_host,
                                             const void* const* phi_img_host, const float* gumbel_noise, float* out, int B,
                                             int nf, int S, int E, int K, int normalize, int mode, int blend, int patch_grid,
                                             int out_channels_last, const float* const* output_mapping_host, void* workspace,
                                             size_t workspace_bytes, void* stream) {
    RF_CHECK_ARG(x_back && x_retr && out && theta_wt_host && theta_b_host && phi_wt_host && phi_b_host && workspace,
                 "rf_attention_fuse_fwd: null pointer");
    RF_CHECK_ARG(B > 0 && nf > 0 && S > 0 && E > 0 && S % E == 0, "rf_attention_fuse_fwd: bad shape");
    RF_CHECK_ARG(K >= 1 && K <= 32, "rf_attention_fuse_fwd: K=%d unsupported (1..32)", K);
    RF_CHECK_ARG(mode == 0 || (mode == 1 && gumbel_noise), "rf_attention_fuse_fwd: retrieval mode needs the Gumbel noise");
    RF_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "rf_attention_fuse_fwd: workspace must be 256-byte aligned");
    const int P = patch_grid < 1 ? 1 : patch_grid;
    const bool mapped = output_mapping_host != nullptr;  // {Wo Wg [nf,nf], Wo bg [nf], bo [nf]}
    RF_CHECK_ARG(!mapped || (output_mapping_host[0] && output_mapping_host[1] && output_mapping_host[2] && !out_channels_last),
                 "rf_attention_fuse_fwd: output mapping needs its three tensors and the NCDHW result layout");
    RF_CHECK_ARG(!mapped || (size_t)8 * nf * E * E * E * sizeof(float) <= 48 * 1024, "rf_attention_fuse_fwd: rows too long for the output mapping");
    RF_CHECK_ARG(S % P == 0 && (S / P) % E == 0, "rf_attention_fuse_fwd: patch grid %d does not tile S=%d into multiples of E=%d", P, S, E);
    AttnWs ws;
    const size_t need = attn_ws_layout(B, nf, S, E, K, &ws, (char*)workspace);
    RF_CHECK_ARG(workspace_bytes >= need, "rf_attention_fuse_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    const int Rp = S / E, rp3 = Rp * Rp * Rp, V = nf * E * E * E;
    const long R = (long)B * rp3;
    RF_CHECK_ARG(R * K * (long)V < (1L << 40), "rf_attention_fuse_fwd: too many rows");
    int rc;
    if ((rc = rf_unfold3d(x_back, ws.xu, B, nf, S, E, stream))) return rc;
    if (P > 1) {
        // x_retr = the un-folded patches [B*K*P^3, nf, S/P, S/P, S/P]: their rows in (patch, local) order, no Fold3D
        const long NP = (long)B * K * P * P * P;
        RF_CHECK_ARG(NP < (1L << 31), "rf_attention_fuse_fwd: too many patches");
        if (S / P == 8 && E == 2 && (((uintptr_t)x_retr | (uintptr_t)ws.pu) & 15) == 0 && (size_t)nf * 532 * sizeof(float) <= 48 * 1024) {
            unfold_e2_patch8_kernel<<<(unsigned)NP, 256, (size_t)nf * 532 * sizeof(float), (cudaStream_t)stream>>>(x_retr, ws.pu, nf);
            RF_LAUNCH_OK("unfold_e2_patch8_kernel");
        } else if ((rc = rf_unfold3d(x_retr, ws.pu, (int)NP, nf, S / P, E, stream))) {
            return rc;
        }
    } else if ((rc = rf_unfold3d(x_retr, ws.pu, B * K, nf, S, E, stream))) {
        return rc;
    }
    if ((rc = run_mlp(ws.xu, R, V, theta_wt_host, theta_b_host, theta_img_host, ws.ha, ws.hb, ws.xf, stream))) return rc;
    if ((rc = run_mlp(ws.pu, R * K, V, phi_wt_host, phi_b_host, phi_img_host, ws.ha, ws.hb, ws.pf, stream))) return rc;
    const float sharp = (float)(FEAT * E * E * E * 4);  // model/attention.py:105
    const unsigned egrid = (unsigned)rf_cdivl(R * 32, 256);
    AttnGeo g;
    g.P = P; g.out_cl = out_channels_last ? 1 : 0; g.nf = nf; g.E = E; g.Rp = Rp;
    g.d_nf = make_fastdiv(nf); g.d_E = make_fastdiv(E); g.d_Rp = make_fastdiv(Rp); g.d_ps = make_fastdiv(Rp / P);
    auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
    g.s_nf = lg(nf); g.s_E = lg(E); g.s_Rp = lg(Rp); g.s_ps = lg(Rp / P); g.s_P = lg(P);
    g.pow2 = (g.s_nf >= 0 && g.s_E >= 0 && g.s_Rp >= 0 && g.s_ps >= 0 && g.s_P >= 0) ? 1 : 0;
    g.side = mapped ? ws.ha : nullptr;  // (the hidden-activation buffers are free once both MLPs have run; >= R * 128 floats)
    float* erows = g.out_cl ? out : ws.orows;  // channels-last: the epilogue stores the volume itself
#define RF_ATTN_EPI(KT, EX)                                                                                         \
    attention_epilogue_kernel<KT, EX><<<egrid, 256, 0, (cudaStream_t)stream>>>(ws.xf, ws.pf, ws.xu, ws.pu, gumbel_noise, \
                                                                                erows, R, rp3, K, V, normalize, mode, blend, sharp, g)
    if (K == 4 && V == 128) RF_ATTN_EPI(4, true);
    else if (K == 8 && V == 128) RF_ATTN_EPI(8, true);
    else if (K <= 4) RF_ATTN_EPI(4, false);
    else if (K <= 8) RF_ATTN_EPI(8, false);
    else if (K <= 16) RF_ATTN_EPI(16, false);
    else RF_ATTN_EPI(32, false);
#undef RF_ATTN_EPI
    RF_LAUNCH_OK("attention_epilogue_kernel");
    if (mapped) {
        attention_output_mapping_kernel<<<egrid, 256, (size_t)8 * V * sizeof(float), (cudaStream_t)stream>>>(
            ws.orows, ws.xu, ws.ha, output_mapping_host[0], output_mapping_host[1], output_mapping_host[2], R, V, nf, E * E * E, blend);
        RF_LAUNCH_OK("attention_output_mapping_kernel");
    }
    if (g.out_cl) return 0;
    return rf_fold3d(ws.orows, out, B, nf, Rp, E, stream);
}

extern "C" int rf_attention_fuse_fwd(const float* x_back, const float* x_retr, const float* const* theta_wt_host,
                                     const float* const* theta_b_host, const float* const* phi_wt_host,
                                     const float* const* phi_b_host, const void* const* theta_img_host,
                                     const void* const* phi_img_host, const float* gumbel_noise, float* out, int B, int nf,
                                     int S, int E, int K, int normalize, int mode, int blend, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    return rf_attention_fuse_patched_fwd(x_back, x_retr, theta_wt_host, theta_b_host, phi_wt_host, phi_b_host, theta_img_host,
                                         phi_img_host, gumbel_noise, out, B, nf, S, E, K, normalize, mode, blend, 1, 0, nullptr,
                                         workspace, workspace_bytes, stream);
}

extern "C" int rf_attention_features(const float* x, const float* t, const uint8_t* occ, const float* const* theta_wt_host,
                                     const float* const* theta_b_host, const float* const* phi_wt_host,
                                     const float* const* phi_b_host, const void* const* theta_img_host,
                                     const void* const* phi_img_host, float* x_feat, float* p_feat, uint8_t* occ_any, int B,
                                     int nf, int S, int E, int normalize, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    RF_CHECK_ARG(x && t && x_feat && p_feat && theta_wt_host && theta_b_host && phi_wt_host && phi_b_host && workspace,
                 "rf_attention_features: null pointer");
    RF_CHECK_ARG(B > 0 && nf > 0 && S > 0 && E > 0 && S % E == 0, "rf_attention_features: bad shape");
    RF_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "rf_attention_features: workspace must be 256-byte aligned");
    AttnWs ws;
    const size_t need = attn_ws_layout(B, nf, S, E, 1, &ws, (char*)workspace);
    RF_CHECK_ARG(workspace_bytes >= need, "rf_attention_features: workspace too small (%zu < %zu)", workspace_bytes, need);
    const int Rp = S / E, V = nf * E * E * E;
    const long R = (long)B * Rp * Rp * Rp;
    int rc;
    if ((rc = rf_unfold3d(x, ws.xu, B, nf, S, E, stream))) return rc;
    if ((rc = rf_unfold3d(t, ws.pu, B, nf, S, E, stream))) return rc;
    if ((rc = run_mlp(ws.xu, R, V, theta_wt_host, theta_b_host, theta_img_host, ws.ha, ws.hb, normalize ? ws.xf : x_feat, stream))) return rc;
    if ((rc = run_mlp(ws.pu, R, V, phi_wt_host, phi_b_host, phi_img_host, ws.ha, ws.hb, normalize ? ws.pf : p_feat, stream))) return rc;
    if (normalize) {
        if ((rc = rf_l2_normalize_rows(ws.xf, x_feat, R, FEAT, 1e-12f, stream))) return rc;
        if ((rc = rf_l2_normalize_rows(ws.pf, p_feat, R, FEAT, 1e-12f, stream))) return rc;
    }
    if (occ && occ_any) {
        occ_any_kernel<<<rf_grid_1d(R, 256), 256, 0, (cudaStream_t)stream>>>(occ, occ_any, B, S, E);
        RF_LAUNCH_OK("occ_any_kernel");
    }
    return 0;
}
