// "Shifted window" tensor-core convolution for the U-Net 'gcr' blocks
// (model/unet.py:19-100: GroupNorm -> Conv3d k3 p1 -> ReLU), 3x3x3, stride 1,
// zero padding 1.
//
// rf_tc_conv3d_fwd (rf_tc_conv.cu) gathers every tap's operand rows from L2 again:
// 27x the activation bytes per layer, which bounds that kernel far below the MMA
// rate.  Here the activation block is staged in shared memory ONCE and the tensor
// core reads all 27 shifted windows of it in place:
//
//   * rf_cl_norm_split_halo writes the normalised activations (virtual concat of x and
//     the nearest-upsampled x2 included) as COMPACT fp16 hi / lo "slot planes"
//         [channel chunk][sample][D][H][W] x 16 B   (slot = 8 channels of one voxel).
//     One TMA tile copy (cp.async.bulk.tensor.4d) per plane moves a patch / slab WITH its
//     halo to shared memory: the box starts one voxel outside the volume and the TMA unit
//     zero-fills what lies out of bounds, so the zero padding of the convolution never
//     exists in HBM (the first version kept haloed planes in HBM: 1.4-2x the bytes on both
//     sides, and interior lines that start mid-sector).
//   * UMMA operand descriptors in the NO-swizzle K-major layout address core matrices
//     of 8 rows x 16 B at arbitrary 16-byte granularity: 8 consecutive slots are the 8
//     rows of a core matrix, the next 8-row group sits SBO bytes further, the second
//     K chunk (channels 8..15 of the K = 16 step) LBO bytes further (the other chunk
//     plane).  Tap (kd,kh,kw) of the convolution is therefore the SAME staged block
//     with the start address advanced by ((kd*Hs + kh)*Wp + kw) slots: no im2col, no
//     copies, one descriptor add per MMA.
//   * GEMM rows are positions of the HALOED block ("linear" mode: 128 consecutive
//     slots per M tile; "lines" mode when W % 8 == 0: 16 lines x 8 slots, SBO = one
//     line), rows that fall on halo positions are computed and dropped (64-94 % of
//     the rows are real outputs, depending on the extent).
//   * fp16 hi/lo split with three products per step (hi*hi + hi*lo + lo*hi), fp32
//     accumulators for the whole item in TMEM (n_tiles x Npad columns <= 512).
//   * Single-chunk inputs (C <= 8) pair two taps into one K = 16 step: LBO = 16 B makes
//     the second K chunk the neighbouring slot, i.e. tap kw+1.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "rf_tc_common.cuh"

namespace {
using namespace rf_tc;

// TMA tile copy global -> shared of one 4-D box (coordinates in elements of the tensor map, innermost first; parts of
// the box outside the tensor are zero-filled and still counted in the barrier's transaction bytes)
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

constexpr int TM = 128, NTHREADS = 512, NB_MAX = 8, MAX_ABUF = 2;  // NB_MAX: most slots of the weight ring (4 or 8 are used)
constexpr int SMEM_LIMIT = 232448;  // 227 KiB opt-in maximum per CTA on sm_100
// same for the 5-D form (slot words, W, H, D, plane) used when a haloed line exceeds 256 8-byte words (W > 126)
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
        : "memory");
}

// ------------------------------------------------------------------ activations -> haloed slot planes
struct SplitArgs {
    const float *x, *x2, *mu, *a, *beta;
    uint4 *hi, *lo;
    int N, D, H, W, C1, C2, CC1, CC;
    float scale;
    FastDiv fCC, fW, fH, fD;  // exact 32-bit divisions of the flat index (total < 2^32 is checked on the host)
    FastDiv fCG;              // octet mapping: chunk groups of 4
    unsigned n_vox;
    long total_oct;
    int wp;         // W-pair planes: [chunk][w parity][n][d][h][w / 2] (rf_cl_norm_split_halo_wp)
    long half_vol;  // D * H * W / 2
};

template <int OCT>
__global__ void __launch_bounds__(256) cl_norm_split_halo_kernel(const SplitArgs s) {
    const int c_tot = s.C1 + s.C2;
    // OCT = 0: thread <-> (voxel, channel chunk) with the CHUNK fastest: consecutive threads read consecutive 32-byte
    // pieces of one voxel's channels, and the 32 / CC voxels a warp covers per chunk land in consecutive slots of that
    // chunk's plane.  With many chunks (joins: 12 or 24) that leaves 43-byte runs per plane and store instruction.
    // OCT = 1 (CC >= 5): a warp covers 8 consecutive voxels x 4 consecutive chunks, lane = chunk * 8 + voxel: every
    // store instruction writes four aligned 128-byte runs, every load instruction reads whole 32-byte sectors of 128
    // contiguous bytes per voxel.
    // (the flat index is decomposed with multiply-high divisions: five 64-bit div / mod pairs per 48 bytes of traffic
    // made this kernel issue-bound at a third of the HBM rate)
    const long total = OCT ? s.total_oct : (long)s.CC * s.n_vox;
    for (long j = blockIdx.x * (long)blockDim.x + threadIdx.x; j < total; j += (long)gridDim.x * blockDim.x) {
        int cc;
        unsigned v;  // voxel in (n, d, h, w) order
        if (OCT) {
            unsigned t = (unsigned)(j >> 5);
            cc = (int)fd_divmod(t, s.fCG) * 4 + (int)((j >> 3) & 3);
            v = t * 8u + (unsigned)(j & 7);
            if (cc >= s.CC || v >= s.n_vox) continue;
        } else {
            unsigned t = (unsigned)j;
            cc = (int)fd_divmod(t, s.fCC);
            v = t;
        }
        long i = (long)cc * s.n_vox + v;  // slot index in the compact planes [chunk][n][d][h][w]
        unsigned t = v;
        const int w = (int)fd_divmod(t, s.fW);
        const int hq = (int)fd_divmod(t, s.fH);
        const int d = (int)fd_divmod(t, s.fD);
        const int n = (int)t;
        // W-pair planes: the even and the odd voxels of every line form two planes of half-lines
        if (s.wp) i = ((long)(cc * 2 + (w & 1)) * s.N + n) * s.half_vol + ((long)d * s.H + hq) * (s.W >> 1) + (w >> 1);
        uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
        const float* src;
        int C, c0, goff;
        if (cc < s.CC1) {
            C = s.C1; c0 = cc * 8; goff = 0;
            src = s.x + (long)v * C;
        } else {
            C = s.C2; c0 = (cc - s.CC1) * 8; goff = s.C1;
            src = s.x2 + ((((long)n * (s.D >> 1) + (d >> 1)) * (s.H >> 1) + (hq >> 1)) * (s.W >> 1) + (w >> 1)) * C;
        }
        float f[8];
        const bool full = (C & 3) == 0 && c0 + 8 <= C;
        if (full) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + c0));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + c0 + 4));
            f[0] = v0.x; f[1] = v0.y; f[2] = v0.z; f[3] = v0.w; f[4] = v1.x; f[5] = v1.y; f[6] = v1.z; f[7] = v1.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = c0 + e < C ? __ldg(src + c0 + e) : 0.f;
        }
        if (s.mu) {
            const long si = (long)n * c_tot + goff + c0;
            if (full && ((c_tot | goff) & 3) == 0) {  // whole chunk, 16-byte aligned statistics: six vector loads
                float m[8], sa[8], sb[8];
                *reinterpret_cast<float4*>(m) = __ldg(reinterpret_cast<const float4*>(s.mu + si));
                *reinterpret_cast<float4*>(m + 4) = __ldg(reinterpret_cast<const float4*>(s.mu + si + 4));
                *reinterpret_cast<float4*>(sa) = __ldg(reinterpret_cast<const float4*>(s.a + si));
                *reinterpret_cast<float4*>(sa + 4) = __ldg(reinterpret_cast<const float4*>(s.a + si + 4));
                *reinterpret_cast<float4*>(sb) = __ldg(reinterpret_cast<const float4*>(s.beta + goff + c0));
                *reinterpret_cast<float4*>(sb + 4) = __ldg(reinterpret_cast<const float4*>(s.beta + goff + c0 + 4));
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e] - m[e], sa[e], sb[e]);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (c0 + e < C) f[e] = fmaf(f[e] - __ldg(s.mu + si + e), __ldg(s.a + si + e), __ldg(s.beta + goff + c0 + e));
            }
        }
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const float v0 = c0 + e < C ? f[e] * s.scale : 0.f, v1 = c0 + e + 1 < C ? f[e + 1] * s.scale : 0.f;
            split_f16x2(v0, v1, h[e >> 1], l[e >> 1]);
        }
        s.hi[i] = make_uint4(h[0], h[1], h[2], h[3]);
        s.lo[i] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// Parity planes for the stride-2 layers: [chunk][parity (d,h,w)][n][ceil(D/2)][ceil(H/2)][ceil(W/2)] x 16 B, i.e. the 8
// sub-lattices of every sample as dense volumes, so that the convolution stages an item's 8 parity sub-blocks with plain
// TMA boxes (element-strided boxes over the compact planes fetched every sub-block's bounding box: the 32 -> 64 layer of
// PCPatch48 at 42^3 ran slower than the gathering kernel).  Slots past an odd extent are written as zeros (a real row
// multiplies them by the zero weights of a padded tap).  No GroupNorm: the conv patch encoders have none.
struct SplitP8Args {
    const float* x;
    uint4 *hi, *lo;
    int N, D, H, W, C, CC, D2, H2, W2;
    float scale;
    FastDiv fCC, fW2, fH2, fD2, fN;
    long total;
};

__global__ void __launch_bounds__(256) cl_split_parity_planes_kernel(const SplitP8Args s) {
    for (long j = blockIdx.x * (long)blockDim.x + threadIdx.x; j < s.total; j += (long)gridDim.x * blockDim.x) {
        unsigned t = (unsigned)j;  // (parity, n, d2, h2, w2, chunk), chunk fastest: a voxel's channels are read as one run
        const int cc = (int)fd_divmod(t, s.fCC);
        const int w2 = (int)fd_divmod(t, s.fW2);
        const int h2 = (int)fd_divmod(t, s.fH2);
        const int d2 = (int)fd_divmod(t, s.fD2);
        const int n = (int)fd_divmod(t, s.fN);
        const int pq = (int)t;
        const int d = 2 * d2 + (pq >> 2), h = 2 * h2 + ((pq >> 1) & 1), w = 2 * w2 + (pq & 1);
        uint32_t hh[4] = {0, 0, 0, 0}, ll[4] = {0, 0, 0, 0};
        if (d < s.D && h < s.H && w < s.W) {
            const float* src = s.x + ((((long)n * s.D + d) * s.H + h) * s.W + w) * s.C + cc * 8;
            float f[8];
            if ((s.C & 3) == 0 && cc * 8 + 8 <= s.C) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                f[0] = v0.x; f[1] = v0.y; f[2] = v0.z; f[3] = v0.w; f[4] = v1.x; f[5] = v1.y; f[6] = v1.z; f[7] = v1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = cc * 8 + e < s.C ? __ldg(src + e) : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; e += 2) split_f16x2(f[e] * s.scale, f[e + 1] * s.scale, hh[e >> 1], ll[e >> 1]);
        }
        const long i = ((((long)(cc * 8 + pq) * s.N + n) * s.D2 + d2) * s.H2 + h2) * s.W2 + w2;
        s.hi[i] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
        s.lo[i] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    }
}

// ------------------------------------------------------------------ layer modes and the weight image
// A K = 16 step of the MMA reads two 16-byte chunks per GEMM row: the slot at the step's offset and the slot LBO
// further.  What the two chunks are depends on the layer:
//   mode 0 (>= 2 channel chunks): one tap, the two chunk planes of a channel-chunk pair ("stage"); 27 steps per stage;
//   mode 1 (one channel chunk, C <= 8): two neighbouring TAPS of the same plane (LBO = one slot): (kw 0, kw 1) and
//          (kw 2, zeros) per (kd,kh) line, 18 steps (pairing across lines, 14 steps with other LBOs, measured slower:
//          8 -> 16 @ 16^3 3.98 -> 4.43 ms);
//   mode 2 (ONE input channel, kernel edge 3 or 5): the slot of voxel w holds the 8 consecutive values x[w .. w+7] of
//          its line ("W-run", written by rf_cl_norm_split_wrun), so one chunk carries all kw taps of a (kd,kh) line and
//          a step pairs two lines: 5 steps for 3^3, 13 for 5^3 - the single-channel first layers of the U-Nets and
//          patch encoders on tensor cores instead of the fp32 FMA kernel;
//   mode 3 ("W pairs", small Cout): a GEMM row is a PAIR of output voxels (w, w+1), N = 2 Cout.  For N <= 64 the tensor
//          pipe is bound by reading the A operand (40-48 cycles per M128 K16 MMA whatever N is), so halving the rows at
//          twice the columns halves the pipe time.  The pair reads the 4 input positions u = 2r .. 2r+3 (padded
//          coordinates) of each (kd,kh) line: the planes are W-de-interleaved (rf_cl_norm_split_halo_wp), the item's
//          block is staged as two sub-blocks A = even u, B = odd u (P_sub slots apart), and a K = 16 step pairs
//          (A[r+p], B[r+p]), p = 0 / 1, with LBO = P_sub: 18 steps per channel chunk and row pair, none of them
//          padding (a Toeplitz expansion of the weights over [N,D,H,W/2,2C] views needs 27 half-empty steps).
//          One stage per channel chunk.  B rows: n = v * Cout + co for output voxel v of the pair.
// Steps are grouped kpg at a time into the weight ring's slots (n_groups groups per stage; padded steps carry zero
// weights).  Image: [stage][group][k step][K chunk][hi rows | lo rows (Npad each)][16 B].
struct Layer {
    int mode, KS, stride;  // stride 2: 'valid' 3^3 only (the patch encoders' down-sampling layers)
    int wp;                // mode 3: W pairs (two output voxels per GEMM row)
    int pool;              // W pairs + 'planes' items: the epilogue writes MaxPool3d(2) of the activated output
    int gn_out;            // whole-sample items: the epilogue applies the NEXT layer's GroupNorm and writes its operand planes
    int C1, C2, Cp1, Cp2, CC, CCe, Cout, Npad;
    int ck, n_stages, kpg, n_groups;
    int hd, hw;  // block extent beyond the output extent: D / H (KS - 1) and W (KS - 1; 0 in mode 2)
};

__global__ void __launch_bounds__(256) halo_weight_image_kernel(const float* __restrict__ w, const Layer L, float scale,
                                                                uint8_t* __restrict__ img) {
    const int Cin = L.C1 + L.C2, Npad = L.Npad, kpg = L.kpg, taps = L.KS * L.KS * L.KS;
    const long total = (long)L.n_stages * L.n_groups * kpg * 2 * Npad;
    const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (gid >= total) return;
    long t = gid;
    const int n = (int)(t % Npad); t /= Npad;
    const int kc = (int)(t % 2); t /= 2;
    const int ks = (int)(t % kpg); t /= kpg;
    const int g = (int)(t % L.n_groups);
    const int st = (int)(t / L.n_groups);
    const int idx = g * kpg + ks;
    uint32_t hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
    if (n < (L.wp ? 2 * L.Cout : L.Cout)) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float val = 0.f;
            if (L.mode == 3) {
                // step 2 l + p of line l = (kd, kh): chunk kc = position u = 2 p + kc of the pair's window; output voxel
                // v of the pair sees it as tap kw = u - v
                const int v = n >= L.Cout ? 1 : 0, co = n - v * L.Cout;
                const int kw = 2 * (idx & 1) + kc - v;
                const int cs = st * 8 + e;
                int ci = -1;
                if (cs < L.Cp1) { if (cs < L.C1) ci = cs; }
                else if (cs - L.Cp1 < L.C2) ci = L.C1 + (cs - L.Cp1);
                if (kw >= 0 && kw < 3 && ci >= 0) val = w[((long)co * Cin + ci) * 27 + (idx >> 1) * 3 + kw] * scale;
            } else if (L.mode == 2) {
                const int line = 2 * idx + kc;  // (kd, kh); element e of the chunk <-> kw
                if (line < L.KS * L.KS && e < L.KS) val = w[(long)n * taps + line * L.KS + e] * scale;
            } else {
                // mode 1: step 2 l = taps (kw 0, kw 1) of line l = (kd, kh), step 2 l + 1 = (kw 2, zeros)
                const int tap = L.mode == 1 ? ((idx & 1) && kc ? 27 : (idx >> 1) * 3 + (idx & 1) * 2 + kc) : idx;
                const int cs = (L.mode == 1 ? 0 : st * 2 + kc) * 8 + e;
                int ci = -1;
                if (cs < L.Cp1) { if (cs < L.C1) ci = cs; }
                else if (cs - L.Cp1 < L.C2) ci = L.C1 + (cs - L.Cp1);
                if (tap < 27 && ci >= 0) val = w[((long)n * Cin + ci) * 27 + tap] * scale;
            }
            uint32_t hv, lv;
            split_f16(val, hv, lv);
            hi[e >> 1] |= hv << (16 * (e & 1));
            lo[e >> 1] |= lv << (16 * (e & 1));
        }
    }
    // one k step = a K-major operand of 2*Npad rows: rows [0, Npad) the hi parts, rows [Npad, 2 Npad) the lo parts, so
    // that [W_hi; W_lo] can be ONE N = 2 Npad operand (LBO = 2 Npad * 16 B between the two K chunks)
    const long step = (long)Npad * 64;
    uint8_t* base = img + ((long)(st * L.n_groups + g) * kpg + ks) * step + (long)kc * (2L * Npad * 16) + (long)n * 16;
    *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + (long)Npad * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------ the convolution
// phase timestamps of CTA 0's first items (tuning aid, read back by rf_tc_conv3d_halo_debug_read)
__device__ long long g_halo_dbg[64];

struct HaloArgs {
    const uint8_t* wimg;
    const float* bias;
    float* y;
    int N, D, H, W, Hp, Wp;
    int pad, CC, tm5;             // conv padding (0 / 1); real channel chunks (chunk >= CC: all zero); 5-D tensor maps
    int s2, P_sub;                // stride 2: 8 parity sub-blocks per plane, P_sub slots apart
    int nb_shift;                 // log2 of the weight ring's slots
    int pool;                     // epilogue max-pools 2x2x2 (W pairs, planes mode): y is [N, D/2, H/2, W/2, Cb]
    // gn_out: the epilogue computes the GroupNorm statistics of the sample's activated output, normalises, splits and
    // writes the operand planes of the next convolution (o_hi / o_lo; o_wp: its W-pair layout) instead of y
    int gn_out, o_wp, o_cpg;
    float o_eps, o_scale;
    const float *o_gamma, *o_beta;
    uint4 *o_hi, *o_lo;
    int wp, Cb;                   // W pairs: 2 sub-blocks per plane (even / odd padded positions); channels of the bias vector
    long V;                       // slots per haloed sample volume
    int Dt, Ht, Wt, Hs, G, stacked;  // item = G stacked whole samples, or a Dt x Ht x Wt slab of one sample
    int n_dt, n_ht, n_wt, Ls;     // slabs per sample; lines per stacked sample (Dp * Hp)
    int lines, n_wblk, n_tiles, P;  // P = slots per staged plane (incl. over-read slack)
    int S_st;                     // staged slots per plane that the bulk copies fill
    int n_stages, nbuf, ck, kpg, n_groups, w0;  // w0: W coordinate of the block's first slot (-pad; 0 for W-runs)
    int Cout, Npad, act, out_ncdhw, n_iss, n_items, n_sets;
    int n_fused;  // tiles [0, n_fused) use the fused cross-product scheme (2 Npad accumulator columns), the rest two passes
    float slope, out_scale;
    uint32_t bslot_bytes, tmem_cols;
    uint16_t tile_off[32];        // first slot of M tile t within the staged block (uniform-indexed constant loads)
    uint32_t ktab[32];            // k step -> slot offset of its first chunk | (slots to its second chunk) << 16
};

// Persistent: CTA c walks over items c, c + gridDim.x, ...  Barriers, TMEM and the zeroed staging buffers are set
// up once; the producers run ahead into the next item while the epilogue warps drain the accumulators, so launch,
// allocation and first-load latency (10-30 thousand cycles per item when every item was its own CTA) are paid once
// per CTA instead of once per item.
//   warp 0: activation producer   warp 3: weight producer   warps 1,2,4-7: MMA issuers   warps 8-15: epilogue
//
// tm_hi / tm_lo: 4-D tensor maps of the compact hi / lo planes, dims (2 W 8-byte words, H, D, CC * N), box = the
// haloed block of one item (2 (W+2), Hs, Dt+2, G); coordinates that fall outside are zero-filled by the TMA unit.
// RES = CTAs per SM the register budget allows: 2 (64 registers; small items whose shared memory lets two CTAs share
// an SM) or 1 (128 registers: the epilogue keeps the next accumulator block in flight).
// EPI = 1: the instantiation that carries the special epilogues (pooling, next layer's GroupNorm), kept out of the
// standard kernels' register budget.
template <int RES, int PIPE, int EPI = 0>
__global__ void __launch_bounds__(NTHREADS, RES) tc_conv3d_halo_kernel(const HaloArgs a, const __grid_constant__ CUtensorMap tm_hi,
                                                                       const __grid_constant__ CUtensorMap tm_lo) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (base - smem_u32(smem_raw));
    const int planes = 2 * a.ck;
    const uint32_t abuf_bytes = (uint32_t)planes * (uint32_t)a.P * 16u;
    const uint32_t sA = base;
    const uint32_t sB = sA + (uint32_t)a.nbuf * abuf_bytes;
    const uint32_t NB = 1u << a.nb_shift;  // slots of the weight ring
    const uint32_t bars = sB + NB * a.bslot_bytes;
    const uint32_t bar_afull = bars, bar_aempty = bars + 8 * MAX_ABUF;
    const uint32_t bar_bfull = bars + 16 * MAX_ABUF, bar_bempty = bar_bfull + 8 * NB_MAX;
    const uint32_t bar_dfull = bar_bempty + 8 * NB_MAX, bar_dempty = bar_dfull + 16;  // one pair per accumulator set
    const uint32_t tmem_slot = bar_dempty + 16;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - base));
    // row table: GEMM row (tile, r) -> output voxel offset relative to the item's first voxel, stacked sample index in
    // the top 6 bits; -1 = halo row.  Built once per CTA: the decode costs several integer divisions per row, which a
    // lone epilogue warp per scheduler cannot hide.
    int* row_tab = reinterpret_cast<int*>(smem_al + (tmem_slot + 16 - base));

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp: uniform register
    const bool resident = a.n_stages == 1;  // one stage per item: loaded once, used by both accumulation passes

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nbuf; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, a.n_iss); }
        for (int s = 0; s < NB_MAX; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, a.n_iss); }
        for (int k = 0; k < 2; ++k) { mbar_init(bar_dfull + 8 * k, a.n_iss); mbar_init(bar_dempty + 8 * k, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Zero the staging buffers once: the tile copies never fill the over-read slack behind the item's block.  Operand
    // rows that touch it are dropped, but in pair mode a REAL row multiplies one such slot by a zero weight, and
    // 0 * NaN would poison it.
    {
        const int total = a.nbuf * planes * a.P;
        for (int i = threadIdx.x; i < total; i += NTHREADS) *reinterpret_cast<uint4*>(smem_al + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = threadIdx.x; i < a.n_tiles * TM; i += NTHREADS) {
        const int t = i >> 7, r = i & 127;
        int line, w;
        if (a.lines == 2) {  // "planes": tile = the 16 lines of ONE d plane of the slab (Ht = 16): no halo line is a GEMM row
            line = (t / a.n_wblk) * a.Hs + (r >> 3);
            w = (t % a.n_wblk) * 8 + (r & 7);
        } else if (a.lines) {
            line = (t / a.n_wblk) * 16 + (r >> 3);
            w = (t % a.n_wblk) * 8 + (r & 7);
        } else {
            line = i / a.Wp; w = i % a.Wp;
        }
        const int g = a.stacked ? line / a.Ls : 0;
        const int rem = a.stacked ? line % a.Ls : line;
        const int dd = rem / a.Hs, hh = rem % a.Hs;
        const bool valid = g < a.G && dd < a.Dt && hh < a.Ht && w < a.Wt;
        row_tab[i] = valid ? ((((g * a.D + dd) * a.H + hh) * a.W + w) | (g << 26)) : -1;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    const int per_sample = a.n_dt * a.n_ht * a.n_wt;

    if (warp == 0) {
      if (lane == 0) {
        // ---- activation producer: one stage = ck channel chunks x (hi, lo) planes of the item's haloed block (one
        // TMA tile copy per plane); two passes over the channel stages (cross products first, then hi * hi; see the
        // issuers)
        const int n_pass = a.n_fused == a.n_tiles ? 1 : 2;
        const int n_loads = resident ? 1 : n_pass * a.n_stages;
        const uint32_t plane_bytes = (uint32_t)a.S_st * 16u * (a.s2 ? 8u : a.wp ? 2u : 1u);
        uint32_t lc = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
            int n0, d0 = 0, h0 = 0, w0 = 0;
            if (a.stacked) { n0 = item * a.G; }
            else {
                n0 = item / per_sample;
                int r = item % per_sample;
                w0 = (r % a.n_wt) * a.Wt; r /= a.n_wt;
                d0 = (r / a.n_ht) * a.Dt; h0 = (r % a.n_ht) * a.Ht;
            }
            for (int l = 0; l < n_loads; ++l, ++lc) {
                const int s = l >= a.n_stages ? l - a.n_stages : l;
                const uint32_t b = lc % (uint32_t)a.nbuf;
                mbar_wait_relaxed(bar_aempty + 8 * b, ((lc / (uint32_t)a.nbuf) & 1u) ^ 1u);
                mbar_arrive_expect_tx(bar_afull + 8 * b, plane_bytes * planes);
                for (int pl = 0; pl < planes; ++pl) {
                    const int hl = pl / a.ck, c = pl % a.ck;
                    const int cc = s * a.ck + c;
                    // a padding chunk (cc >= CC) or the samples past N of a ragged last item lie outside the tensor: zeros
                    const int plane = cc < a.CC ? cc * a.N + n0 : a.CC * a.N;
                    const uint32_t dst = sA + b * abuf_bytes + (uint32_t)pl * (uint32_t)a.P * 16u;
                    if (a.s2) {
                        // stride 2: the 8 parity sub-blocks of the item, dense boxes of the parity planes
                        // [chunk][parity][n][D/2][H/2][W/2] (rf_cl_split_parity_planes); sub-block coordinates = output coordinates
                        for (int pq = 0; pq < 8; ++pq)
                            tma_load_5d(dst + (uint32_t)(pq * a.P_sub) * 16u, hl ? &tm_lo : &tm_hi, 0, w0, h0, d0,
                                        cc < a.CC ? (cc * 8 + pq) * a.N + n0 : 8 * a.CC * a.N, bar_afull + 8 * b);
                    } else if (a.wp) {
                        // W pairs: sub-block A = the padded positions u = 2 r (voxels of parity 'pad', from half-line index
                        // w0 - pad), sub-block B = u = 2 r + 1 (the other parity, from w0); planes [chunk][parity][n]
                        for (int sb = 0; sb < 2; ++sb) {
                            const int par = sb ? 1 - a.pad : a.pad, ws = w0 - (sb ? 0 : a.pad);
                            const int pq = cc < a.CC ? (cc * 2 + par) * a.N + n0 : 2 * a.CC * a.N;
                            const uint32_t dq = dst + (uint32_t)(sb * a.P_sub) * 16u;
                            if (a.tm5) tma_load_5d(dq, hl ? &tm_lo : &tm_hi, 0, ws, h0 - a.pad, d0 - a.pad, pq, bar_afull + 8 * b);
                            else tma_load_4d(dq, hl ? &tm_lo : &tm_hi, 2 * ws, h0 - a.pad, d0 - a.pad, pq, bar_afull + 8 * b);
                        }
                    } else if (a.tm5) tma_load_5d(dst, hl ? &tm_lo : &tm_hi, 0, a.w0 + w0, h0 - a.pad, d0 - a.pad, plane, bar_afull + 8 * b);
                    else tma_load_4d(dst, hl ? &tm_lo : &tm_hi, 2 * (a.w0 + w0), h0 - a.pad, d0 - a.pad, plane, bar_afull + 8 * b);
                }
            }
        }
      }
    } else if (warp == 3) {
      if (lane == 0) {
        // ---- weight producer: ring of (kd,kh) groups; the second pass of every item streams the same groups again
        const int per_pass = a.n_stages * a.n_groups, n_pass = a.n_fused == a.n_tiles ? 1 : 2;
        uint32_t gc = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x)
            for (int gi = 0; gi < n_pass * per_pass; ++gi, ++gc) {
                const uint32_t sl = gc & (NB - 1u);
                const int img = gi >= per_pass ? gi - per_pass : gi;
                mbar_wait_relaxed(bar_bempty + 8 * sl, ((gc >> a.nb_shift) & 1u) ^ 1u);
                mbar_arrive_expect_tx(bar_bfull + 8 * sl, a.bslot_bytes);
                bulk_g2s(sB + sl * a.bslot_bytes, a.wimg + (long)img * a.bslot_bytes, a.bslot_bytes, bar_bfull + 8 * sl);
            }
      }
    } else if (warp < 8) {
        // ---- MMA issuers: warps 1, 2, 4, 5, 6, 7 -> issuer 0..5; issuer i owns the M tiles t = i (mod n_iss).
        // One warp cannot feed the tensor pipe here: every tcgen05.mma needs its descriptors moved from vector to
        // uniform registers (R2UR) under an elected lane, ~150 cycles per MMA measured, while the pipe needs 39-48
        // cycles for an M128 K16 step with N <= 64.  Tiles are independent accumulators, so several warps issue
        // concurrently (each warp's own MMAs stay ordered, which is all one accumulator needs); every issuer
        // commits to the stage / ring / accumulator barriers, which count n_iss arrivals.
        const int iss = warp <= 2 ? warp - 1 : warp - 2;
        if (iss < a.n_iss) {
            const uint32_t leader = elect_one();
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t idesc = idesc_f16(a.Npad), idesc2 = idesc_f16(2 * a.Npad);
            // descriptor = {lo: start >> 4 | (LBO >> 4) << 16, hi: SBO >> 4 | version 1 << 14}; addresses advance in
            // 16-byte units = slots, so "+ slots" on the low word moves the window.
            const uint32_t a_hi32 = (a.lines ? (uint32_t)a.Wp : 8u) | (1u << 14);
            const uint32_t b_hi32 = 8u | (1u << 14);
            const uint32_t b_lbo = (uint32_t)(2 * a.Npad) << 16;  // K chunks of a k step are 2 Npad rows apart
            const uint32_t step_u = (uint32_t)a.Npad * 4u;        // one k step of weights, in 16-byte units
            const uint32_t lo_off = (uint32_t)(a.ck * a.P);     // hi -> lo plane, in slots
            const uint32_t npad = (uint32_t)a.Npad, nf = (uint32_t)a.n_fused;
            const uint32_t set_cols = (nf + (uint32_t)a.n_tiles) * npad;  // fused tiles take 2 Npad columns
            const int n_vs = (a.n_fused == a.n_tiles ? 1 : 2) * a.n_stages;
            // The tensor core truncates when it aligns the 16 products of a K step with the fp32 accumulator, a
            // bias that grows with the number of accumulations at full magnitude (measured: error linear in the
            // MMA count).  So the two small cross products (hi*lo, lo*hi: 2^-11 of the result) of ALL taps and
            // stages are accumulated first, while the accumulator is tiny, and the hi*hi products in a second pass
            // over the stages: a third of the accumulations happen at full magnitude.
            uint32_t lc = 0, gc = 0, it = 0;
            for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
                const bool dbg = blockIdx.x == 0 && iss == 0 && leader && it < 6;
                if (dbg) g_halo_dbg[it * 8 + 0] = clock64();
                // accumulator set of this item (two sets when they fit TMEM: the epilogue of item i overlaps the MMAs of
                // item i+1); wait until the epilogue has drained the set's previous item
                const uint32_t set = a.n_sets == 2 ? (it & 1u) : 0u, use = a.n_sets == 2 ? (it >> 1) : it;
                mbar_wait_warp_backoff(bar_dempty + 8 * set, (use & 1u) ^ 1u, 200);
                tc_fence_after();
                if (dbg) g_halo_dbg[it * 8 + 1] = clock64();
                uint32_t b = 0;
                for (int vs = 0; vs < n_vs; ++vs) {
                    const int pass = vs >= a.n_stages ? 1 : 0;
                    if (!resident || vs == 0) {
                        b = lc % (uint32_t)a.nbuf;
                        mbar_wait_warp_backoff(bar_afull + 8 * b, (lc / (uint32_t)a.nbuf) & 1u, 100);
                        tc_fence_after();
                        if (dbg && vs == 0) g_halo_dbg[it * 8 + 2] = clock64();
                    }
                    const uint32_t abase = ((sA + b * abuf_bytes) & 0x3FFFFu) >> 4;
                    // Issue order inside a weight group: tile-major.  The per-tile terms (accumulator address, first slot,
                    // scheme) are derived once per (group, tile) and the k step table entries (slot offset | LBO << 16, added
                    // to the descriptor's low word) of up to four steps are fetched before the wait, so an MMA costs a handful
                    // of uniform-datapath instructions: with one tile per issuer the k-step-major loop spent ~350 cycles
                    // per step and warp on dependent constant loads and address arithmetic, which bounded items of few tiles
                    // (3 tiles: 118 cycles per step and tile against 88 of pipe time).
                    for (int g = 0; g < a.n_groups; ++g, ++gc) {
                        const uint32_t sl = gc & (NB - 1u);
                        const int k0 = g * a.kpg;
                        for (int kc0 = 0; kc0 < a.kpg; kc0 += 4) {
                            uint32_t ktg[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) ktg[j] = a.ktab[(k0 + kc0 + j) & 31];
                            if (kc0 == 0) {
                                mbar_wait_warp_backoff(bar_bfull + 8 * sl, (gc >> a.nb_shift) & 1u, 40);
                                tc_fence_after();
                            }
                            const uint32_t bbase = ((((sB + sl * a.bslot_bytes) & 0x3FFFFu) >> 4) | b_lbo) + (uint32_t)kc0 * step_u;
                            const int nk = a.kpg - kc0 < 4 ? a.kpg - kc0 : 4;
                            for (int t = iss; t < a.n_tiles; t += a.n_iss) {
                                const bool tf = (uint32_t)t < nf;
                                if (tf && pass) continue;  // (mixed items: the second pass serves the two-pass tiles only)
                                const uint32_t d = tmem_u + set * set_cols + ((uint32_t)t + (tf ? (uint32_t)t : nf)) * npad;
                                const uint32_t da0 = abase + a.tile_off[t];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if (j >= nk) break;
                                    const uint32_t b_hi = bbase + (uint32_t)j * step_u, b_lo = b_hi + npad;  // lo rows follow the hi rows
                                    const uint32_t acc = (vs | g | kc0 | j) ? 1u : 0u;  // the very first MMA of a tile overwrites its accumulator
                                    const uint32_t da = da0 + ktg[j];
                                    if (tf) {
                                        // A_hi x [W_hi; W_lo] as ONE N = 2 Npad MMA -> columns [main | cross]; A_lo x W_hi into
                                        // the cross block.  For N <= 64 the pipe time is set by reading the A operand, so
                                        // the doubled N is almost free: two MMAs instead of three, one pass over the stages,
                                        // and the main block still sees only the hi*hi accumulations.
                                        tc_mma2(d, da, a_hi32, b_hi, b_hi32, idesc2, acc, leader);                 // hi * [hi | lo]
                                        tc_mma2(d + npad, da + lo_off, a_hi32, b_hi, b_hi32, idesc, 1u, leader);   // lo * hi -> cross
                                    } else if (pass == 0) {
                                        tc_mma2(d, da, a_hi32, b_lo, b_hi32, idesc, acc, leader);          // hi * lo
                                        tc_mma2(d, da + lo_off, a_hi32, b_hi, b_hi32, idesc, 1u, leader);  // lo * hi
                                    } else {
                                        tc_mma2(d, da, a_hi32, b_hi, b_hi32, idesc, 1u, leader);           // hi * hi
                                    }
                                }
                            }
                        }
                        if (leader) tc_commit(bar_bempty + 8 * sl);
                    }
                    if (!resident || vs == n_vs - 1) {
                        if (leader) tc_commit(bar_aempty + 8 * b);
                        ++lc;
                    }
                }
                if (leader) tc_commit(bar_dfull + 8 * set);
                if (dbg) g_halo_dbg[it * 8 + 3] = clock64();
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue warps 8..15: warp % 4 = TMEM lane quadrant, (warp - 8) / 4 = tile parity; one thread <-> one
        // GEMM row of its tiles
        const int q = warp & 3, half = (warp - 8) >> 2;
        const long So = (long)a.D * a.H * a.W;
        const bool vec4 = !a.out_ncdhw && (a.Cout & 3) == 0;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
            int n0, gact = 1, d0 = 0, h0 = 0, w0 = 0;
            if (a.stacked) { n0 = item * a.G; gact = min(a.G, a.N - n0); }
            else {
                n0 = item / per_sample;
                int r = item % per_sample;
                w0 = (r % a.n_wt) * a.Wt; r /= a.n_wt;
                d0 = (r / a.n_ht) * a.Dt; h0 = (r % a.n_ht) * a.Ht;
            }
            const long vox0 = (((long)n0 * a.D + d0) * a.H + h0) * a.W + w0;
            const bool dbg = blockIdx.x == 0 && threadIdx.x == 256 && it < 6;
            if (dbg) g_halo_dbg[it * 8 + 4] = clock64();
            const uint32_t set = a.n_sets == 2 ? (it & 1u) : 0u, use = a.n_sets == 2 ? (it >> 1) : it;
            const uint32_t npad = (uint32_t)a.Npad, nf = (uint32_t)a.n_fused;
            const uint32_t tm_set = tmem_base + set * (nf + (uint32_t)a.n_tiles) * npad;
            mbar_wait_warp_sleepy(bar_dfull + 8 * set, use & 1u);
            tc_fence_after();
            if (dbg) g_halo_dbg[it * 8 + 5] = clock64();
            if constexpr (EPI == 1) {
              if (a.gn_out) {
                // ---- GroupNorm of the NEXT layer in the epilogue (item = one whole sample): phase A reduces the per-channel
                // sums of the activated outputs (warp: transposing shuffle reduction in fp32 over its 32 rows, then fp64
                // per lane; CTA: shared memory), phase B re-reads the accumulators from TMEM, normalises, splits into fp16
                // hi / lo and writes the slots of the next convolution's operand planes.  The fp32 activations, the
                // statistics pass and the split pass over them never touch HBM.
                double* red = reinterpret_cast<double*>(smem_al + (((tmem_slot + 16 - base) + (uint32_t)a.n_tiles * 512u + 15u) & ~15u));
                float* stt = reinterpret_cast<float*>(red + 8 * 2 * 64);  // [3][64]: mean, rstd * gamma, beta per channel
                const int nblk = a.Npad >> 4, ew = warp - 8, etid = threadIdx.x - 256;
                // (tile, 16-column block) -> activated outputs of the thread's row (zeros on rows that are no output voxel)
                auto load_act = [&](int t, int blk, bool valid, float* w16) {
                    const bool tf = (uint32_t)t < nf;
                    const uint32_t ta = tm_set + ((uint32_t)(q * 32) << 16) + ((uint32_t)t + (tf ? (uint32_t)t : nf)) * npad + (uint32_t)(blk << 4);
                    float v[16], u[16];
                    tc_ld16_issue(ta, v);
                    if (tf) tc_ld16_issue(ta + npad, u);
                    tc_ld_wait();
                    tc_ld_fence16(v);
                    if (tf) {
                        tc_ld_fence16(u);
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] += u[e];
                    }
                    float z[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int c = (blk << 4) + e;
                        z[e] = fmaf(v[e], a.out_scale, (a.bias && c < a.Cout) ? __ldg(a.bias + c) : 0.f);
                    }
                    rf_act_vec(z, a.act, a.slope);
#pragma unroll
                    for (int e = 0; e < 16; ++e) w16[e] = valid ? z[e] : 0.f;
                };
                // 16 values per lane -> lanes 2 i, 2 i + 1 hold the warp total of value i
                auto reduce16 = [&](float* v) {
#pragma unroll
                    for (int hf = 8, bit = 16; hf >= 1; hf >>= 1, bit >>= 1) {
                        const bool up = (lane & bit) != 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if (k < hf) {
                                const float send = up ? v[k] : v[k + hf], keep = up ? v[k + hf] : v[k];
                                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                            }
                        }
                    }
                    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
                };
                // ---- phase A, block-major: a thread first adds up its rows of all its tiles, then ONE shuffle reduction per
                // 16 channels (fp32 over <= 32 x tiles rows, fp64 from there on).  (Keeping the loads of two tiles in flight
                // needs more registers than the kernel has: the spills made the epilogue 2.3x slower.)
                for (int blk = 0; blk < nblk; ++blk) {
                    float s16[16], q16[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) { s16[e] = 0.f; q16[e] = 0.f; }
                    for (int t = half; t < a.n_tiles; t += 2) {
                        const int rt = row_tab[t * TM + q * 32 + lane];
                        float w16[16];
                        load_act(t, blk, rt >= 0 && (rt >> 26) < gact, w16);
#pragma unroll
                        for (int e = 0; e < 16; ++e) { s16[e] += w16[e]; q16[e] = fmaf(w16[e], w16[e], q16[e]); }
                    }
                    const double ts = (double)reduce16(s16), tq = (double)reduce16(q16);
                    if (!(lane & 1)) { red[(ew * 2 + 0) * 64 + blk * 16 + (lane >> 1)] = ts; red[(ew * 2 + 1) * 64 + blk * 16 + (lane >> 1)] = tq; }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (etid < a.Cout) {  // channel etid: sums of its group over the 8 warps
                    const int g0 = etid / a.o_cpg * a.o_cpg;
                    double s1 = 0.0, s2 = 0.0;
                    for (int c = g0; c < g0 + a.o_cpg; ++c)
#pragma unroll
                        for (int k = 0; k < 8; ++k) { s1 += red[(k * 2 + 0) * 64 + c]; s2 += red[(k * 2 + 1) * 64 + c]; }
                    const double cnt = (double)a.D * a.H * a.W * a.o_cpg;
                    const double mean = s1 / cnt;
                    double var = s2 / cnt - mean * mean;
                    if (var < 0.0) var = 0.0;
                    stt[etid] = (float)mean;
                    stt[64 + etid] = (float)(1.0 / sqrt(var + (double)a.o_eps)) * __ldg(a.o_gamma + etid);
                    stt[128 + etid] = __ldg(a.o_beta + etid);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                // ---- phase B: accumulators again -> normalise -> split -> the next convolution's operand planes
                const long Vs = (long)a.D * a.H * a.W;
                for (int t = half; t < a.n_tiles; t += 2) {
                    const int rt = row_tab[t * TM + q * 32 + lane];
                    const bool valid = rt >= 0 && (rt >> 26) < gact;
                    const int vox = rt & 0x3FFFFFF;  // (d * H + h) * W + w within the sample
                    long slot = (long)n0 * Vs + vox, cstride = (long)a.N * Vs;
                    if (a.o_wp) {  // W-pair planes [chunk][w parity][n][d][h][w / 2]
                        const int line = vox / a.W, wq = vox - line * a.W;
                        slot = ((long)(wq & 1) * a.N + n0) * (Vs >> 1) + (long)line * (a.W >> 1) + (wq >> 1);
                        cstride = 2 * (long)a.N * (Vs >> 1);
                    }
                    for (int blk = 0; blk < nblk; ++blk) {
                        float w16[16];
                        load_act(t, blk, valid, w16);
                        if (!valid) continue;
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            const int c0 = (blk << 4) + ch * 8;
                            if (c0 >= a.Cout) break;
                            uint32_t hh[4], ll[4];
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                const int c = c0 + e;
                                const float f0 = c < a.Cout ? fmaf(w16[ch * 8 + e] - stt[c], stt[64 + c], stt[128 + c]) * a.o_scale : 0.f;
                                const float f1 = c + 1 < a.Cout ? fmaf(w16[ch * 8 + e + 1] - stt[c + 1], stt[64 + c + 1], stt[128 + c + 1]) * a.o_scale : 0.f;
                                split_f16x2(f0, f1, hh[e >> 1], ll[e >> 1]);
                            }
                            const long i = slot + (long)(c0 >> 3) * cstride;
                            a.o_hi[i] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                            a.o_lo[i] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_dempty + 8 * set);
                continue;
              }
              if (a.pool) {
                // MaxPool3d(2) of the activated output in the epilogue (W pairs, planes mode: tile t = d plane t of the slab,
                // row = (h = q * 4 + lane / 8, pair lane & 7)): the pooling window is the thread's two voxels (w), the
                // same row of lane ^ 8 (h) and the same row of tile t ^ 1 (d).  max commutes with the monotone scale /
                // bias / activation, so those run once on the pooled values.  The full-resolution output never exists.
                const int nkb = a.Cb >> 4;  // 16-channel blocks per voxel
                const int hh = q * 4 + (lane >> 3);
                for (int p = half; p < (a.n_tiles >> 1); p += 2)
                    for (int kb = 0; kb < nkb; ++kb) {
                        float m[16];
#pragma unroll
                        for (int sv = 0; sv < 4; ++sv) {  // (tile, voxel of the pair) = (2 p + sv / 2, sv % 2)
                            const uint32_t t = 2u * (uint32_t)p + (uint32_t)(sv >> 1);
                            const bool tf = t < nf;
                            const uint32_t ta = tm_set + ((uint32_t)(q * 32) << 16) + (t + (tf ? t : nf)) * npad + (uint32_t)(((sv & 1) * nkb + kb) << 4);
                            float v[16], u[16];
                            tc_ld16_issue(ta, v);
                            if (tf) tc_ld16_issue(ta + npad, u);
                            tc_ld_wait();
                            tc_ld_fence16(v);
                            if (tf) {
                                tc_ld_fence16(u);
#pragma unroll
                                for (int e = 0; e < 16; ++e) v[e] += u[e];
                            }
#pragma unroll
                            for (int e = 0; e < 16; ++e) m[e] = sv ? fmaxf(m[e], v[e]) : v[e];
                        }
#pragma unroll
                        for (int e = 0; e < 16; ++e) m[e] = fmaxf(m[e], __shfl_xor_sync(0xffffffffu, m[e], 8));
                        if (!(lane & 8)) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) m[e] = fmaf(m[e], a.out_scale, a.bias ? __ldg(a.bias + kb * 16 + e) : 0.f);
                            rf_act_vec(m, a.act, a.slope);
                            const long po = ((((long)n0 * (a.D >> 1) + ((d0 >> 1) + p)) * (a.H >> 1) + ((h0 + hh) >> 1)) * a.W + (w0 + (lane & 7))) * a.Cb + kb * 16;
#pragma unroll
                            for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(a.y + po + e) = make_float4(m[e], m[e + 1], m[e + 2], m[e + 3]);
                        }
                    }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_dempty + 8 * set);
                continue;
              }
            }
            // (tile, 16-column block) pairs of this warp, software-pipelined: the TMEM loads of the next pair are in
            // flight while the current one is scaled, activated and stored (one pair took ~1200 cycles of exposed
            // TMEM latency + store issue; small-Cout layers were bound by this loop, not by the MMAs)
            const int nblk = a.Npad >> 4;
            const int n_my = ((a.n_tiles - half + 1) >> 1) * nblk;
            auto issue = [&](int j, float* v, float* u) {
                const int t = half + 2 * (j / nblk), c0 = (j % nblk) << 4;
                const bool tf = (uint32_t)t < nf;
                const uint32_t ta = tm_set + ((uint32_t)(q * 32) << 16) + ((uint32_t)t + (tf ? (uint32_t)t : nf)) * npad + (uint32_t)c0;
                tc_ld16_issue(ta, v);
                if (tf) tc_ld16_issue(ta + npad, u);  // main + cross blocks
            };
            auto process = [&](int j, float* v, float* u) {
                tc_ld_fence16(v);
                const int t = half + 2 * (j / nblk), c0 = (j % nblk) << 4;
                if ((uint32_t)t < nf) {
                    tc_ld_fence16(u);
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] += u[e];
                }
                if (dbg && t == 0 && c0 == 0) g_halo_dbg[it * 8 + 7] = clock64();
                if (dbg && t == 2 && c0 == 0) g_halo_dbg[48 + it] = clock64();
                const int rt = row_tab[t * TM + q * 32 + lane];
                const bool valid = rt >= 0 && (rt >> 26) < gact;
                if (!valid) return;
                const long vox = vox0 + (rt & 0x3FFFFFF);
                if (a.bias) {
#pragma unroll
                    for (int e = 0; e < 16; ++e)  // (W pairs: columns [Cb, 2 Cb) are the second voxel's channels)
                        v[e] = fmaf(v[e], a.out_scale, c0 + e < a.Cout ? __ldg(a.bias + (c0 + e >= a.Cb ? c0 + e - a.Cb : c0 + e)) : 0.f);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] *= a.out_scale;
                }
                float w16[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) w16[e] = v[e];
                rf_act_vec(w16, a.act, a.slope);
                if (vec4) {
                    float* dst = a.y + vox * a.Cout + c0;
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        if (c0 + e < a.Cout) *reinterpret_cast<float4*>(dst + e) = make_float4(w16[e], w16[e + 1], w16[e + 2], w16[e + 3]);
                } else if (a.out_ncdhw) {
                    const long nn = vox / So, sp = vox % So;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (c0 + e < a.Cout) a.y[(nn * a.Cout + c0 + e) * So + sp] = w16[e];
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (c0 + e < a.Cout) a.y[vox * a.Cout + c0 + e] = w16[e];
                }
            };
            if constexpr (PIPE == 1) {
                float v0[16], u0[16], v1[16], u1[16];
                if (n_my > 0) issue(0, v0, u0);
                for (int j = 0; j < n_my; j += 2) {
                    tc_ld_wait();
                    if (j + 1 < n_my) issue(j + 1, v1, u1);
                    process(j, v0, u0);
                    if (j + 1 < n_my) {
                        tc_ld_wait();
                        if (j + 2 < n_my) issue(j + 2, v0, u0);
                        process(j + 1, v1, u1);
                    }
                }
            } else {
                float v0[16], u0[16];
                for (int j = 0; j < n_my; ++j) {
                    issue(j, v0, u0);
                    tc_ld_wait();
                    process(j, v0, u0);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_dempty + 8 * set);  // 8 arrivals: the set may be overwritten
            if (dbg) g_halo_dbg[it * 8 + 6] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
    }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup: the library must load (and export its
// symbols) on machines without a driver, so it cannot link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

struct Geo {
    int Wt, P_sub, nb;
    int Dt, Ht, Hs, G, stacked, lines, n_wblk, n_tiles, P, S_st, n_items, nbuf, two_resident, n_sets, fused, n_fused;
    int halo;  // 0: every sample / slab carries its own halo; 1 (stacked, 'same' padding): neighbours share it
    int hd, hw;  // extent of the item's block beyond its outputs in D / H and in W
    uint32_t tmem_cols, bslot;
    size_t smem;
    double score;
};

// Chooses the item shape: maximise (real outputs / computed GEMM rows), penalise grids that leave SMs idle.
bool choose_geometry(int N, int D, int H, int W, const Layer& L, int pad, Geo& best) {
    const int planes = 2 * L.ck, kpg = L.kpg, n_stages = L.n_stages, Npad = L.Npad;
    const uint32_t bslot = (uint32_t)kpg * 2u * (uint32_t)Npad * 32u;
    static const int nb_env = [] { const char* e = getenv("RF_HALO_NB"); return e ? atoi(e) : 0; }();  // tuning aid
    const int NB = nb_env == 8 ? 8 : 4;
    const long avail_all = (long)SMEM_LIMIT - 1024 - 256 - (long)NB * bslot;
    const bool shareable = pad == 1 && L.KS == 3 && L.mode != 2 && L.stride == 1;  // zero padding of one voxel all around
    best.score = -1.0;
    // halo = 1 ("shared halo", stacked whole samples with zero padding only): the TMA box starts one voxel before the
    // volume and ends AT its far faces, so a sample occupies (D+1)(H+1)(W+1) slots whose index-0 faces are zero: the
    // far neighbours of the last voxel of a line / plane / sample are the zero slots that open the next line / plane /
    // sample (the zeroed slack behind the block for the last sample).  51 % of the rows of stacked 4^3 patches are
    // real outputs instead of 30 %.
    const int Wfull = W;
    auto consider = [&](int stacked, int G, int Dt, int Ht, int Wt, int lines, int halo) {
        const int W = Wt;  // the item's extent along w (the full line unless the slab is tiled along w too)
        const int hd = halo ? 1 : L.hd, hw = halo ? 1 : L.hw;
        const int Dp = D + hd, Hp = H + hd, Wp = W + hw;
        const long V = (long)Dp * Hp * Wp;
        const int Hs = stacked ? Hp : Ht + L.hd;
        const long S_st = stacked ? (long)G * V : (long)(Dt + L.hd) * Hs * Wp;
        const int tr = L.stride == 2 ? 1 : L.KS - 1;  // largest tap offset per dimension inside a (sub-)block
        // furthest slot a row reads beyond its own: the last tap (second chunk of the last k step included)
        const long reach = ((long)tr * Hs + tr) * Wp + (L.mode == 2 ? 0 : L.wp ? 1 : tr) + 1;
        const long lines_needed = stacked ? (long)(G - 1) * Dp * Hp + (long)(D - 1) * Hp + H : (long)(Dt - 1) * Hs + Ht;
        long n_tiles, max_slot;
        const int n_wblk = lines ? (W + 7) / 8 : 1;
        if (L.gn_out && !(stacked && G == 1)) return;  // GroupNorm epilogue: one whole sample per item
        if (L.pool && (lines != 2 || (Dt & 1) || n_wblk != 1)) return;  // pooling epilogue: d planes in tile pairs, one 8-pair block per line
        if (lines == 2) {  // planes: one tile per (d plane, 8-voxel block) of a slab with exactly 16 lines per plane
            if (stacked || Ht != 16) return;
            n_tiles = (long)Dt * n_wblk;
            max_slot = ((long)(Dt - 1) * Hs + 15) * Wp + (n_wblk * 8 - 1) + reach;
        } else if (lines) {
            const long lb = (lines_needed + 15) / 16;
            n_tiles = lb * n_wblk;
            max_slot = (lb * 16 - 1) * Wp + (n_wblk * 8 - 1) + reach;
        } else {
            const long rows = (lines_needed - 1) * Wp + W;
            n_tiles = (rows + 127) / 128;
            max_slot = n_tiles * 128 - 1 + reach;
        }
        long P = max_slot + 1 > S_st ? max_slot + 1 : S_st;
        P = (P + 7) / 8 * 8;
        long P_sub = 0;
        if (L.stride == 2) {  // 8 parity sub-blocks, the over-read slack only behind the last one
            P_sub = (S_st + 7) / 8 * 8;
            P += 7 * P_sub;
        }
        if (L.wp) {  // two sub-blocks (even / odd padded positions), each with its own zeroed over-read slack: with a
            P_sub = P;  // shared halo the far neighbours of a sample's last voxels are read from that slack
            P += P_sub;
        }
        const long n_sub = L.stride == 2 ? 8 : L.wp ? 2 : 1;
        if (n_tiles * Npad > 512 || n_tiles > 32 || P * 16 >= (1L << 18)) return;  // (the fused scheme needs twice the columns: checked below)
        const long gn_smem = L.gn_out ? 16 + 8 * 2 * 64 * 8 + 3 * 64 * 4 : 0;  // per-warp channel sums (fp64) + the sample's mean / rstd table
        const long avail = avail_all - n_tiles * 512 - gn_smem;  // the row table: n_tiles x 128 ints
        // two staging buffers whenever they fit: the next stage (or the next item's block) loads during the MMAs
        const long smemA1 = (long)planes * P * 16;
        const int nbuf = 2 * smemA1 <= avail ? 2 : 1;
        if (nbuf == 1 && n_stages > 1) return;
        const long smemA = nbuf * smemA1;
        if (smemA > avail || S_st * n_sub * 16 * planes >= (1L << 20)) return;
        const long outputs = (stacked ? (long)G * D * H * W : (long)Dt * Ht * W) * (L.wp ? 2 : 1);
        const long n_items = stacked ? (N + G - 1) / G : (long)N * (D / Dt) * (H / Ht) * (Wfull / Wt);
        // Cost model, calibrated on B200 (tools/halo_geo_sweep.sh): an M128 K16 MMA occupies the tensor pipe for ~40
        // (N <= 32) to 48 (N = 64) cycles, one issuer warp sustains one MMA per ~150 cycles, the epilogue costs ~1200
        // cycles per (tile pair, 16 columns), and every item pays ~3000 cycles of pipeline fill.  Accumulators are
        // double-buffered (epilogue of item i under the MMAs of item i+1) when two sets fit TMEM; otherwise a second
        // resident CTA hides part of the epilogue.
        const double k_steps = (double)n_stages * L.n_groups * kpg * (double)n_tiles;  // (k step, tile) pairs of one item
        const int n_iss = n_tiles < 6 ? (int)n_tiles : 6;
        const double waves = (double)((n_items + 147) / 148);   // items every SM walks through (the tensor pipe is per SM)
        auto pipe_cycles = [](int n) { return n <= 32 ? 40.0 : n <= 64 ? 48.0 : n <= 128 ? 64.0 : 128.0; };
        const char* force = getenv("RF_HALO_FUSED");  // tuning aid: "0" / "1" forces the MMA scheme
        // fused = 1: two MMAs per step, A_hi x [W_hi; W_lo] (N = 2 Npad) and A_lo x W_hi, accumulators [main | cross]
        // of 2 Npad columns per tile, one pass over the stages; fused = 0: three N = Npad MMAs in two passes.
        // fused = 2 ("mixed"): when TMEM cannot hold 2 Npad columns for every tile, as many leading tiles as fit use the
        // fused scheme and the rest the two-pass scheme (96 -> 56 @ 8^3: 3 of 5 tiles; two passes over the stages, the
        // second serves only the two-pass tiles).  Measured on that layer: 2.25 ms against 2.12 ms for plain two-pass
        // (the second pass leaves three of five issuers idle and becomes issue-bound), so it is only reachable through
        // RF_HALO_FUSED=2.
        for (int fused = 0; fused < 3; ++fused) {
            if (force ? atoi(force) != fused : fused == 2) continue;
            if (fused && 2 * Npad > 256) continue;
            for (int n_sets = 1; n_sets <= 2; ++n_sets) {
                int n_fused = fused == 1 ? (int)n_tiles : 0;
                if (fused == 2) {  // only where the fused scheme cannot cover every tile even with one accumulator set
                    if (n_sets == 2 || 2L * Npad * n_tiles <= 512) continue;
                    n_fused = (int)((512 - n_tiles * Npad) / Npad);
                    if (n_fused < 1 || n_fused >= n_tiles) continue;
                }
                const long cols = (long)(n_tiles + n_fused) * Npad;
                if ((long)n_sets * cols > 512) break;
                uint32_t cols_needed = 32;
                while ((long)cols_needed < cols * n_sets) cols_needed <<= 1;
                const int tile_cols = (int)(cols / n_tiles);  // average, for the epilogue estimate
                const long smem_total = 1024 + smemA + (long)NB * bslot + 256 + n_tiles * 512 + gn_smem;
                static const int force_res = [] { const char* e = getenv("RF_HALO_RES"); return e ? atoi(e) : 0; }();  // tuning aid
                const bool two_resident = force_res != 1 && smem_total <= 113 * 1024 && cols_needed <= 256 && !L.pool && !L.gn_out;
                const double issue = 150.0 / (n_iss * (two_resident ? 2 : 1));
                const double step_fused = fmax(pipe_cycles(2 * Npad), issue) + fmax(pipe_cycles(Npad), issue);
                const double step_two = 3.0 * fmax(pipe_cycles(Npad), issue);
                const double t_mma = (double)n_stages * L.n_groups * kpg * (n_fused * step_fused + (n_tiles - n_fused) * step_two);
                // (per 16 OUTPUT columns: the fused scheme's cross block is a second TMEM load added to the same registers,
                // not a second block of work - counting it as one made the chooser avoid the fused scheme for W pairs)
                const double t_epi = 1000.0 + (double)((n_tiles + 1) / 2) * (Npad / 16) * (n_fused ? 1500.0 : 1200.0);
                (void)tile_cols;
                // staging: ~25 bytes per clock and SM from L2 when every SM pulls (decoder 16 -> 16 @ 64^3: 127 KB per
                // item in ~5000 cycles); hidden behind the MMAs only with a second buffer or a second resident CTA
                const double n_loads = n_stages == 1 ? 1.0 : (fused == 1 ? 1.0 : 2.0) * n_stages;
                const double t_load = (double)planes * (double)(S_st * n_sub) * 16.0 * n_loads / 25.0;
                const double t_core = n_sets == 2 ? fmax(t_mma, t_epi) : t_mma + (two_resident ? 0.3 : 1.0) * t_epi;
                const double t_item = (fused == 1 ? 2000.0 : 3000.0) + ((nbuf == 2 || two_resident) ? fmax(t_core, t_load) : t_core + t_load);
                const double score = (double)outputs * (double)n_items / (waves * t_item);
                if (score > best.score) {
                    best.Wt = Wt; best.P_sub = (int)P_sub;
                    best.Dt = stacked ? D : Dt; best.Ht = stacked ? H : Ht; best.Hs = Hs; best.G = stacked ? G : 1; best.stacked = stacked;
                    best.lines = lines; best.n_wblk = n_wblk; best.n_tiles = (int)n_tiles; best.P = (int)P; best.S_st = (int)S_st;
                    best.n_items = (int)n_items; best.nbuf = nbuf; best.bslot = bslot; best.two_resident = two_resident ? 1 : 0;
                    best.tmem_cols = cols_needed;
                    best.nb = NB; best.n_sets = n_sets; best.fused = fused; best.n_fused = n_fused; best.halo = halo; best.hd = hd; best.hw = hw;
                    best.smem = (size_t)smem_total;
                    best.score = score;
                }
            }
        }
    };
    if (const char* e = getenv("RF_HALO_GEO")) {  // tuning aid: "stacked,G,Dt,Ht,lines" forces the item shape
        int st, G, Dt, Ht, ln, hl = 0;
        if (sscanf(e, "%d,%d,%d,%d,%d,%d", &st, &G, &Dt, &Ht, &ln, &hl) >= 5) {
            int wt = W;
            if (const char* e2 = getenv("RF_HALO_WT")) wt = atoi(e2);
            consider(st, G, st ? D : Dt, st ? H : Ht, st || W % wt ? W : wt, ln, st && shareable && hl == 1 ? 1 : 0);
            return best.score > 0.0;
        }
    }
    for (int lines = 0; lines < 3; ++lines) {
        for (int G = 1; G <= 32 && G <= N && lines < 2; ++G) {  // (the row table keeps the stacked sample index in 5 bits)
            consider(1, G, D, H, W, lines, 0);
            if (shareable) consider(1, G, D, H, W, lines, 1);
        }
        for (int Dt = 1; Dt <= D; ++Dt) {
            if (D % Dt) continue;
            for (int Ht = 1; Ht <= H; ++Ht) {
                if (H % Ht) continue;
                // long lines may be tiled along w as well (halves, quarters, ...: the block's halo overhead and the
                // bytes staged per item shrink, two staging buffers fit)
                consider(0, 1, Dt, Ht, W, lines, 0);
                for (int Wt = W >> 1; Wt >= 16 && W % Wt == 0; Wt >>= 1) consider(0, 1, Dt, Ht, Wt, lines, 0);
            }
        }
    }
    return best.score > 0.0;
}

// mode 0 / 1 layers: 3x3x3 over C1 + C2 channels (x2 upsampled); mode 2: KS^3 over ONE channel (C1 = 1, KS 3 or 5)
bool make_layer(int Cout, int C1, int C2, int KS, bool wrun, Layer& L, int stride = 1, bool wp = false, bool pool = false) {
    if (Cout < 1 || Cout > 256 || C1 < 0 || C2 < 0 || C1 + C2 < 1 || (stride != 1 && stride != 2)) return false;
    if (stride == 2 && (wrun || C2 != 0)) return false;
    if (pool && (!wp || (Cout & 15))) return false;
    L.KS = KS; L.C1 = C1; L.C2 = C2; L.Cout = Cout; L.stride = stride; L.wp = wp ? 1 : 0; L.pool = pool ? 1 : 0; L.gn_out = 0;
    L.Cp1 = round_up(C1, 8); L.Cp2 = round_up(C2, 8);
    L.CC = (L.Cp1 + L.Cp2) / 8;
    L.Npad = round_up(Cout, 16);
    if (wp) {  // W pairs: N = 2 Cout (both scheme's operands [W_hi; W_lo] must stay within N = 256), one stage per chunk
        if (wrun || stride != 1 || KS != 3 || 2 * Cout > 128) return false;
        L.Npad = round_up(2 * Cout, 16);
        L.mode = 3; L.CCe = L.CC; L.ck = 1; L.n_stages = L.CC; L.kpg = 2; L.n_groups = 9;
        L.hd = 2; L.hw = 1;
        return true;
    }
    if (wrun) {
        if (C1 != 1 || C2 != 0 || (KS != 3 && KS != 5)) return false;
        const int n_ks = (KS * KS + 1) / 2;  // two (kd,kh) lines per K = 16 step
        L.mode = 2; L.CCe = 1; L.ck = 1; L.n_stages = 1;
        L.kpg = n_ks <= 6 ? n_ks : 7; L.n_groups = (n_ks + L.kpg - 1) / L.kpg;
        L.hd = KS - 1; L.hw = 0;
        return true;
    }
    if (KS != 3) return false;
    // stride 2: the item's block is split into its 8 parity sub-blocks (tap k reads parity k & 1 at offset k >> 1), each
    // one slot larger than the output extent per dimension
    L.hd = L.hw = stride == 2 ? 1 : 2;
    if (L.CC == 1) { L.mode = 1; L.CCe = 1; L.ck = 1; L.n_stages = 1; L.kpg = 2; L.n_groups = 9; }
    else { L.mode = 0; L.CCe = round_up(L.CC, 2); L.ck = 2; L.n_stages = L.CCe / 2; L.kpg = 3; L.n_groups = 9; }
    return true;
}

size_t weight_image_bytes(const Layer& L) { return (size_t)L.n_stages * L.n_groups * L.kpg * 2 * L.Npad * 32; }

int weight_image(const float* w, const Layer& L, float scale, void* image, void* stream) {
    const long threads = (long)L.n_stages * L.n_groups * L.kpg * 2 * L.Npad;
    halo_weight_image_kernel<<<(unsigned)rf_cdivl(threads, 256), 256, 0, (cudaStream_t)stream>>>(w, L, scale, (uint8_t*)image);
    RF_LAUNCH_OK("halo_weight_image_kernel");
    return 0;
}

int conv_init() {
    RF_SMEM_OPT_IN((tc_conv3d_halo_kernel<1, 1>), SMEM_LIMIT);
    RF_SMEM_OPT_IN((tc_conv3d_halo_kernel<1, 0>), SMEM_LIMIT);
    RF_SMEM_OPT_IN((tc_conv3d_halo_kernel<2, 0>), SMEM_LIMIT);
    RF_SMEM_OPT_IN((tc_conv3d_halo_kernel<1, 0, 1>), SMEM_LIMIT);
    return 0;
}

// D, H, W: OUTPUT extents.  The planes hold Din x Hin x Win slots per (chunk, sample): Din = D + hd - 2 pad, likewise H;
// Win = W + hw - 2 pad (mode 2: Win = W, the W-runs already cover the taps and the padding along W).
struct GnOut {  // the next layer's GroupNorm + operand planes (Layer::gn_out)
    const float *gamma, *beta;
    int groups, wp;
    float eps, scale;
    void *hi, *lo;
};

int launch_conv(const Layer& L, const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N, int D, int H,
                int W, int pad, int act, float slope, float out_scale, int out_ncdhw, void* stream, int Din_in = 0, int Hin_in = 0,
                int Win_in = 0, const GnOut* gn = nullptr) {
    Geo g;
    RF_CHECK_ARG(choose_geometry(N, D, H, W, L, pad, g), "rf_tc_conv3d_halo_fwd: no item shape fits (N=%d out %dx%dx%d Cout=%d C=%d+%d KS=%d)",
                 N, D, H, W, L.Cout, L.C1, L.C2, L.KS);
    HaloArgs a;
    a.wimg = (const uint8_t*)weight_image; a.bias = bias; a.y = y;
    a.N = N; a.D = D; a.H = H; a.W = W; a.Hp = H + g.hd; a.Wp = g.Wt + g.hw;
    a.pad = pad; a.CC = L.CC; a.w0 = L.mode == 2 ? 0 : -pad;
    a.s2 = L.stride == 2 ? 1 : 0; a.P_sub = g.P_sub;
    a.wp = L.wp; a.Cb = L.Cout; a.pool = L.pool; a.gn_out = 0;
    if (gn) { a.gn_out = 1; a.o_wp = gn->wp; a.o_cpg = L.Cout / gn->groups; a.o_eps = gn->eps; a.o_scale = gn->scale; a.o_gamma = gn->gamma; a.o_beta = gn->beta; a.o_hi = (uint4*)gn->hi; a.o_lo = (uint4*)gn->lo; } a.nb_shift = g.nb == 8 ? 3 : 2;
    a.V = (long)(D + g.hd) * (H + g.hd) * (W + g.hw);
    a.Dt = g.Dt; a.Ht = g.Ht; a.Wt = g.Wt; a.Hs = g.Hs; a.G = g.G; a.stacked = g.stacked;
    a.n_dt = D / g.Dt; a.n_ht = H / g.Ht; a.n_wt = W / g.Wt; a.Ls = (D + g.hd) * (H + g.hd);
    a.lines = g.lines; a.n_wblk = g.n_wblk; a.n_tiles = g.n_tiles; a.P = g.P; a.S_st = g.S_st;
    a.n_stages = L.n_stages; a.nbuf = g.nbuf; a.ck = L.ck; a.kpg = L.kpg; a.n_groups = L.n_groups;
    a.Cout = L.wp ? 2 * L.Cout : L.Cout; a.Npad = L.Npad; a.act = act; a.out_ncdhw = out_ncdhw; a.slope = slope; a.out_scale = out_scale;
    a.bslot_bytes = g.bslot; a.tmem_cols = g.tmem_cols;
    a.n_iss = g.n_tiles < 6 ? g.n_tiles : 6;
    // tile t = lb * n_wblk + wb starts at slot lb * (16 lines) + wb * 8 (linear mode: n_wblk = 1, 128 slots per tile)
    RF_CHECK_ARG(g.n_tiles <= 32, "rf_tc_conv3d_halo_fwd: internal: more than 32 M tiles");
    for (int t = 0; t < 32; ++t) {
        const long off = g.lines == 2 ? (long)(t / g.n_wblk) * g.Hs * a.Wp + (t % g.n_wblk) * 8
                       : g.lines ? (long)(t / g.n_wblk) * 16 * a.Wp + (t % g.n_wblk) * 8 : (long)t * 128;
        a.tile_off[t] = (uint16_t)(t < g.n_tiles ? off : 0);
    }
    // k step table: slot offset of the step's first chunk, distance to its second chunk
    RF_CHECK_ARG(L.kpg * L.n_groups <= 32, "rf_tc_conv3d_halo_fwd: internal: more than 32 k steps per stage");
    for (int i = 0; i < 32; ++i) {
        long off = 0, lbo = 1;
        auto tap_off = [&](int t) {
            const int kd = t / 9, kh = (t / 3) % 3, kw = t % 3;
            if (L.stride == 2)  // parity sub-block (kd & 1, kh & 1, kw & 1), offset (kd >> 1, kh >> 1, kw >> 1) inside it
                return (long)((kd & 1) * 4 + (kh & 1) * 2 + (kw & 1)) * a.P_sub + ((long)(kd >> 1) * a.Hs + (kh >> 1)) * a.Wp + (kw >> 1);
            return ((long)kd * a.Hs + kh) * a.Wp + kw;
        };
        auto line_off = [&](int l) { return ((long)(l / L.KS) * a.Hs + l % L.KS) * a.Wp; };
        if (i < L.kpg * L.n_groups) {
            if (L.mode == 3) { off = ((long)(i / 6) * a.Hs + (i / 2) % 3) * a.Wp + (i & 1); lbo = a.P_sub; }  // (A[r + p], B[r + p]) of line i / 2
            else if (L.mode == 0) { off = tap_off(i); lbo = a.P; }
            else if (L.mode == 1) {  // (kw 0, kw 1) and (kw 2, zeros) of line i / 2
                const int t0 = (i / 2) * 3 + (i % 2) * 2;
                off = tap_off(t0);
                if (L.stride == 2 && !(i % 2)) lbo = tap_off(t0 + 1) - off;  // kw 1 lives in the next parity sub-block
            }
            else { if (2 * i < L.KS * L.KS) { off = line_off(2 * i); if (2 * i + 1 < L.KS * L.KS) lbo = line_off(2 * i + 1) - off; } }
        }
        RF_CHECK_ARG(off >= 0 && off < 65536 && lbo > 0 && lbo < 16384, "rf_tc_conv3d_halo_fwd: internal: k step offsets out of range");
        a.ktab[i] = (uint32_t)off | ((uint32_t)lbo << 16);
    }
    if (int rc = conv_init()) return rc;
    a.n_items = g.n_items; a.n_sets = g.n_sets; a.n_fused = g.n_fused;
    // tensor maps of the compact planes [CC * N][Din][Hin][Win] x 16 B, seen as 8-byte words so that a whole line of
    // the item's block is the innermost box extent (<= 256 elements)
    // (stride 2: the parity planes hold ceil(extent / 2) slots per dimension)
    const int Din = Din_in ? (Din_in + 1) / 2 : D + L.hd - 2 * pad, Hin = Hin_in ? (Hin_in + 1) / 2 : H + L.hd - 2 * pad;
    const int Win = Win_in ? (Win_in + 1) / 2 : (L.mode == 2 ? W : L.wp ? W + 1 - pad : W + L.hw - 2 * pad);  // (W pairs: half-lines)
    const int bW = g.Wt + g.hw, bD = g.stacked ? D + g.hd : g.Dt + L.hd;  // the item's box (its H extent is g.Hs)
    RF_CHECK_ARG(bW <= 256 && g.Hs <= 256 && bD <= 256 && g.G <= 256, "rf_tc_conv3d_halo_fwd: item box exceeds the TMA limits");
    a.tm5 = (a.s2 || 2 * bW > 256) ? 1 : 0;  // a line longer than 256 words, or strided boxes: slots as a dimension of their own
    CUtensorMap tm[2];
    const EncodeTiledFn encode = encode_tiled_fn();
    RF_CHECK_ARG(encode != nullptr, "rf_tc_conv3d_halo_fwd: the driver does not export cuTensorMapEncodeTiled");
    for (int k = 0; k < 2; ++k) {
        const cuuint32_t bG = (cuuint32_t)(g.stacked ? g.G : 1);
        const cuuint64_t planes = (cuuint64_t)L.CC * (cuuint64_t)N * (L.wp ? 2u : a.s2 ? 8u : 1u);
        const cuuint32_t es = 1u;
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult cr;
        if (a.tm5) {
            const cuuint64_t gdim[5] = {2, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)Din, planes};
            const cuuint64_t gstr[4] = {16, 16ull * Win, 16ull * Win * Hin, 16ull * Win * Hin * Din};
            const cuuint32_t box[5] = {2, (cuuint32_t)(es * bW - (es - 1)), (cuuint32_t)(es * g.Hs - (es - 1)), (cuuint32_t)(es * bD - (es - 1)), bG};
            cr = encode(&tm[k], CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<void*>(k ? lo : hi), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            const cuuint64_t gdim[4] = {(cuuint64_t)(2 * Win), (cuuint64_t)Hin, (cuuint64_t)Din, planes};
            const cuuint64_t gstr[3] = {16ull * Win, 16ull * Win * Hin, 16ull * Win * Hin * Din};
            const cuuint32_t box[4] = {(cuuint32_t)(2 * bW), (cuuint32_t)g.Hs, (cuuint32_t)bD, bG};
            cr = encode(&tm[k], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(k ? lo : hi), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        RF_CHECK_ARG(cr == CUDA_SUCCESS, "rf_tc_conv3d_halo_fwd: cuTensorMapEncodeTiled failed (%d) for planes %dx%dx%dx%d, item box %dx%dx%dx%u", (int)cr,
                     Win, Hin, Din, L.CC * N, bW, g.Hs, bD, bG);
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int slots = sms * (g.two_resident ? 2 : 1);
    const unsigned grid = (unsigned)(g.n_items < slots ? g.n_items : slots);
    static const int pipe = [] { const char* e = getenv("RF_HALO_PIPE"); return e ? atoi(e) : 1; }();  // tuning aid
    if (a.pool || a.gn_out) tc_conv3d_halo_kernel<1, 0, 1><<<grid, NTHREADS, g.smem, (cudaStream_t)stream>>>(a, tm[0], tm[1]);
    else if (g.two_resident) tc_conv3d_halo_kernel<2, 0><<<grid, NTHREADS, g.smem, (cudaStream_t)stream>>>(a, tm[0], tm[1]);
    else if (pipe) tc_conv3d_halo_kernel<1, 1><<<grid, NTHREADS, g.smem, (cudaStream_t)stream>>>(a, tm[0], tm[1]);
    else tc_conv3d_halo_kernel<1, 0><<<grid, NTHREADS, g.smem, (cudaStream_t)stream>>>(a, tm[0], tm[1]);
    RF_LAUNCH_OK("tc_conv3d_halo_kernel");
    return 0;
}

}  // namespace

int rf_tc_conv_halo_init() { return conv_init(); }

extern "C" size_t rf_halo_act_bytes(int N, int D, int H, int W, int C1, int C2, int pad) {
    Layer L;
    if (!make_layer(16, C1, C2, 3, false, L) || N < 1 || D < 1 || H < 1 || W < 1 || pad < 0 || pad > 1) return 0;
    return (size_t)L.CC * N * (size_t)D * H * W * 16;  // compact planes: the halo is made by the TMA unit's zero fill
}

static int split_halo(const float* x, int C1, const float* x2, int C2, const float* gn_mu, const float* gn_a, const float* gn_beta,
                      void* hi, void* lo, int N, int D, int H, int W, int pad, float scale, int wp, void* stream) {
    Layer L;
    RF_CHECK_ARG(make_layer(16, C1, C2, 3, false, L), "rf_cl_norm_split_halo: bad channel counts");
    RF_CHECK_ARG(hi && lo && (C1 == 0 || x) && (C2 == 0 || x2) && N > 0 && D > 0 && H > 0 && W > 0 && (pad == 0 || pad == 1),
                 "rf_cl_norm_split_halo: bad arguments");
    RF_CHECK_ARG(C2 == 0 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "rf_cl_norm_split_halo: upsampled input needs even extents");
    RF_CHECK_ARG(!wp || W % 2 == 0, "rf_cl_norm_split_halo_wp: W-pair planes need an even W");
    RF_CHECK_ARG((gn_mu == nullptr) == (gn_a == nullptr) && (gn_mu == nullptr) == (gn_beta == nullptr), "rf_cl_norm_split_halo: partial GroupNorm arguments");
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)x2 & 15) == 0,
                 "rf_cl_norm_split_halo: pointers must be 16-byte aligned");
    SplitArgs s;
    s.x = x; s.x2 = x2; s.mu = gn_mu; s.a = gn_a; s.beta = gn_beta; s.hi = (uint4*)hi; s.lo = (uint4*)lo;
    s.N = N; s.D = D; s.H = H; s.W = W; s.C1 = C1; s.C2 = C2; s.CC1 = L.Cp1 / 8; s.CC = L.CC; s.scale = scale;
    const int CC = L.CC;
    const long n_vox = (long)N * D * H * W;
    const long total = (long)CC * n_vox;
    RF_CHECK_ARG(total < (1L << 32) - 256, "rf_cl_norm_split_halo: more than 2^32 slots");
    s.fCC = make_fastdiv(CC);
    s.fW = make_fastdiv(W);
    s.fH = make_fastdiv(H);
    s.fD = make_fastdiv(D);
    const int CG = (CC + 3) / 4;
    s.fCG = make_fastdiv(CG);
    s.n_vox = (unsigned)n_vox;
    s.wp = wp; s.half_vol = (long)D * H * W / 2;
    s.total_oct = (n_vox + 7) / 8 * CG * 32;
    if (CC >= 5)
        cl_norm_split_halo_kernel<1><<<rf_grid_1d(s.total_oct, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(s);
    else
        cl_norm_split_halo_kernel<0><<<rf_grid_1d(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(s);
    RF_LAUNCH_OK("cl_norm_split_halo_kernel");
    return 0;
}

extern "C" int rf_cl_norm_split_halo(const float* x, int C1, const float* x2, int C2, const float* gn_mu, const float* gn_a,
                                     const float* gn_beta, void* hi, void* lo, int N, int D, int H, int W, int pad, float scale,
                                     int interior_only, void* stream) {
    (void)interior_only;  // the planes have no halo any more: every call writes every slot
    return split_halo(x, C1, x2, C2, gn_mu, gn_a, gn_beta, hi, lo, N, D, H, W, pad, scale, 0, stream);
}

/* Same activations as W-PAIR planes [chunk][w parity][n][d][h][w / 2] x 16 B (same size as rf_halo_act_bytes, W even):
 * the operand layout of rf_tc_conv3d_halo_wp_fwd. */
extern "C" int rf_cl_norm_split_halo_wp(const float* x, int C1, const float* x2, int C2, const float* gn_mu, const float* gn_a,
                                        const float* gn_beta, void* hi, void* lo, int N, int D, int H, int W, int pad, float scale,
                                        void* stream) {
    return split_halo(x, C1, x2, C2, gn_mu, gn_a, gn_beta, hi, lo, N, D, H, W, pad, scale, 1, stream);
}

extern "C" size_t rf_tc_conv_halo_weight_image_bytes(int Cout, int C1, int C2) {
    Layer L;
    return make_layer(Cout, C1, C2, 3, false, L) ? weight_image_bytes(L) : 0;
}

extern "C" int rf_tc_conv_halo_weight_image(const float* w, int Cout, int C1, int C2, float scale, void* image, void* stream) {
    Layer L;
    RF_CHECK_ARG(w && image, "rf_tc_conv_halo_weight_image: null pointer");
    RF_CHECK_ARG(make_layer(Cout, C1, C2, 3, false, L), "rf_tc_conv_halo_weight_image: unsupported shape");
    RF_CHECK_ARG(((uintptr_t)image & 15) == 0, "rf_tc_conv_halo_weight_image: image must be 16-byte aligned");
    return weight_image(w, L, scale, image, stream);
}

/* 1 when rf_tc_conv3d_halo_fwd can run this layer (an item shape fits shared memory and TMEM).  D, H, W are the
 * INPUT extents; the output extents are D + 2 pad - 2 (pad 1: 'same', pad 0: 'valid'). */
extern "C" int rf_tc_conv3d_halo_supported(int N, int D, int H, int W, int Cout, int C1, int C2, int pad) {
    Layer L;
    if (!make_layer(Cout, C1, C2, 3, false, L) || N < 1 || pad < 0 || pad > 1) return 0;
    const int Do = D + 2 * pad - 2, Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
    if (Do < 1 || Ho < 1 || Wo < 1) return 0;
    if ((long)N * D * H * W * L.CC >= (1L << 32) - 256) return 0;
    Geo g;
    return choose_geometry(N, Do, Ho, Wo, L, pad, g) && g.n_tiles <= 32 ? 1 : 0;
}

extern "C" int rf_tc_conv3d_halo_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y,
                                     int N, int D, int H, int W, int pad, int Cout, int C1, int C2, int act, float slope,
                                     float out_scale, int out_ncdhw, void* stream) {
    RF_CHECK_ARG(hi && lo && weight_image && y, "rf_tc_conv3d_halo_fwd: null pointer");
    Layer L;
    RF_CHECK_ARG(make_layer(Cout, C1, C2, 3, false, L) && N > 0 && (pad == 0 || pad == 1),
                 "rf_tc_conv3d_halo_fwd: unsupported shape Cout=%d C1=%d C2=%d pad=%d", Cout, C1, C2, pad);
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)weight_image & 15) == 0 && ((uintptr_t)y & 15) == 0,
                 "rf_tc_conv3d_halo_fwd: pointers must be 16-byte aligned");
    // from here on D, H, W are the OUTPUT extents
    D += 2 * pad - 2; H += 2 * pad - 2; W += 2 * pad - 2;
    RF_CHECK_ARG(D > 0 && H > 0 && W > 0, "rf_tc_conv3d_halo_fwd: empty output");
    return launch_conv(L, hi, lo, weight_image, bias, y, N, D, H, W, pad, act, slope, out_scale, out_ncdhw, stream);
}

/* Debug / test aid: the item shape the chooser picks (returns 0 when unsupported). */
extern "C" int rf_tc_conv3d_halo_geometry(int N, int D, int H, int W, int Cout, int C1, int C2, int pad, int* out8) {
    Layer L;
    if (!make_layer(Cout, C1, C2, 3, false, L)) return 0;
    Geo g;
    if (!choose_geometry(N, D + 2 * pad - 2, H + 2 * pad - 2, W + 2 * pad - 2, L, pad, g)) return 0;
    out8[0] = g.stacked; out8[1] = g.G; out8[2] = g.Dt; out8[3] = g.Ht; out8[4] = (g.lines & 1) + 2 * (g.fused != 0) + 4 * (g.halo == 1) + 8 * (g.fused == 2) + 64 * (g.lines == 2); out8[5] = g.n_tiles;
    out8[6] = g.n_items; out8[7] = (int)g.smem;
    return 1;
}

/* Operand planes of the stride-2 layers: the 8 parity sub-lattices of every sample as dense volumes
 * [chunk][parity][n][ceil(D/2)][ceil(H/2)][ceil(W/2)] x 16 B (fp16 hi / lo of scale * x; slots past an odd extent are zero). */
extern "C" size_t rf_halo_s2_act_bytes(int N, int D, int H, int W, int C1) {
    if (N < 1 || D < 1 || H < 1 || W < 1 || C1 < 1) return 0;
    return (size_t)((C1 + 7) / 8) * 8 * N * (size_t)((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2) * 16;
}

extern "C" int rf_cl_split_parity_planes(const float* x, int C1, void* hi, void* lo, int N, int D, int H, int W, float scale, void* stream) {
    RF_CHECK_ARG(x && hi && lo && N > 0 && D > 0 && H > 0 && W > 0 && C1 > 0, "rf_cl_split_parity_planes: bad arguments");
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)x & 15) == 0, "rf_cl_split_parity_planes: pointers must be 16-byte aligned");
    SplitP8Args s;
    s.x = x; s.hi = (uint4*)hi; s.lo = (uint4*)lo; s.N = N; s.D = D; s.H = H; s.W = W; s.C = C1; s.CC = (C1 + 7) / 8;
    s.D2 = (D + 1) / 2; s.H2 = (H + 1) / 2; s.W2 = (W + 1) / 2; s.scale = scale;
    s.total = (long)s.CC * 8 * N * s.D2 * s.H2 * s.W2;
    RF_CHECK_ARG(s.total < (1L << 32) - 256, "rf_cl_split_parity_planes: more than 2^32 slots");
    s.fCC = make_fastdiv(s.CC); s.fW2 = make_fastdiv(s.W2); s.fH2 = make_fastdiv(s.H2); s.fD2 = make_fastdiv(s.D2); s.fN = make_fastdiv(N);
    cl_split_parity_planes_kernel<<<rf_grid_1d(s.total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(s);
    RF_LAUNCH_OK("cl_split_parity_planes_kernel");
    return 0;
}

/* Stride-2 'valid' 3x3x3 convolution (the down-sampling layers of the conv patch encoders, model/retrieval.py:4-28,
 * 187-275) on the same kernel: the planes are the parity planes of rf_cl_split_parity_planes; the producer stages the item's 8
 * parity sub-blocks (plain TMA boxes) and tap (kd,kh,kw) reads sub-block (kd&1, kh&1, kw&1) at offset
 * (kd>>1, kh>>1, kw>>1).  D, H, W: INPUT extents; outputs (D - 3) / 2 + 1 etc.  Weight image: rf_tc_conv_halo_weight_image. */
extern "C" int rf_tc_conv3d_halo_s2_supported(int N, int D, int H, int W, int Cout, int C1) {
    Layer L;
    if (!make_layer(Cout, C1, 0, 3, false, L, 2) || N < 1 || D < 3 || H < 3 || W < 3) return 0;
    if ((long)N * ((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2) * L.CC * 8 >= (1L << 32) - 256) return 0;
    Geo g;
    return choose_geometry(N, (D - 3) / 2 + 1, (H - 3) / 2 + 1, (W - 3) / 2 + 1, L, 0, g) && g.n_tiles <= 32 ? 1 : 0;
}

extern "C" int rf_tc_conv3d_halo_s2_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N,
                                        int D, int H, int W, int Cout, int C1, int act, float slope, float out_scale, int out_ncdhw,
                                        void* stream) {
    RF_CHECK_ARG(hi && lo && weight_image && y, "rf_tc_conv3d_halo_s2_fwd: null pointer");
    Layer L;
    RF_CHECK_ARG(make_layer(Cout, C1, 0, 3, false, L, 2) && N > 0 && D >= 3 && H >= 3 && W >= 3,
                 "rf_tc_conv3d_halo_s2_fwd: unsupported shape Cout=%d C=%d in %dx%dx%d", Cout, C1, D, H, W);
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)weight_image & 15) == 0 && ((uintptr_t)y & 15) == 0,
                 "rf_tc_conv3d_halo_s2_fwd: pointers must be 16-byte aligned");
    return launch_conv(L, hi, lo, weight_image, bias, y, N, (D - 3) / 2 + 1, (H - 3) / 2 + 1, (W - 3) / 2 + 1, 0, act, slope, out_scale,
                       out_ncdhw, stream, D, H, W);
}

/* W-pair variant of the 3x3x3 stride-1 layer for small Cout (2 Cout <= 128): a GEMM row is a PAIR of output voxels
 * (w, w + 1) with N = 2 Cout columns - half the rows at the same pipe time per MMA (the tensor pipe is bound by the A
 * read for N <= 64), 18 K steps per channel chunk and row pair.  Planes: rf_cl_norm_split_halo_wp; weight image:
 * rf_tc_conv_halo_wp_weight_image.  D, H, W: INPUT extents (output W must be even).  _supported: 0 = no, 1 = runs,
 * 2 = runs and the item cost model rates it faster than rf_tc_conv3d_halo_fwd on this shape.  Output: channels-last. */
extern "C" size_t rf_tc_conv_halo_wp_weight_image_bytes(int Cout, int C1, int C2) {
    Layer L;
    return make_layer(Cout, C1, C2, 3, false, L, 1, true) ? weight_image_bytes(L) : 0;
}

extern "C" int rf_tc_conv_halo_wp_weight_image(const float* w, int Cout, int C1, int C2, float scale, void* image, void* stream) {
    Layer L;
    RF_CHECK_ARG(w && image, "rf_tc_conv_halo_wp_weight_image: null pointer");
    RF_CHECK_ARG(make_layer(Cout, C1, C2, 3, false, L, 1, true), "rf_tc_conv_halo_wp_weight_image: unsupported shape");
    RF_CHECK_ARG(((uintptr_t)image & 15) == 0, "rf_tc_conv_halo_wp_weight_image: image must be 16-byte aligned");
    return weight_image(w, L, scale, image, stream);
}

extern "C" int rf_tc_conv3d_halo_wp_supported(int N, int D, int H, int W, int Cout, int C1, int C2, int pad) {
    Layer L, L0;
    if (!make_layer(Cout, C1, C2, 3, false, L, 1, true) || N < 1 || pad < 0 || pad > 1) return 0;
    const int Do = D + 2 * pad - 2, Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
    if (Do < 1 || Ho < 1 || Wo < 2 || (Wo & 1)) return 0;
    if ((long)N * D * H * W * L.CC >= (1L << 32) - 256) return 0;
    Geo g, g0;
    if (!choose_geometry(N, Do, Ho, Wo / 2, L, pad, g) || g.n_tiles > 32) return 0;
    if (make_layer(Cout, C1, C2, 3, false, L0) && choose_geometry(N, Do, Ho, Wo, L0, pad, g0) && g0.n_tiles <= 32 && 1.15 * g0.score >= g.score) return 1;  // (the model overrates the variant: measured 1.15-1.25x where it says 1.2-1.5x, and losses below that)
    return 2;
}

/* W-pair convolution whose epilogue writes MaxPool3d(2) of the activated output: y is [N, D/2, H/2, W/2, Cout]
 * channels-last (model/unet.py:210-253: an encoder level whose full-resolution output only feeds the next level's pooling).
 * Needs Cout % 16 == 0, 'same' padding, 16-line slabs ("planes" items); _supported says whether an item shape exists. */
extern "C" int rf_tc_conv3d_halo_wp_pool_supported(int N, int D, int H, int W, int Cout, int C1, int C2) {
    Layer L;
    if (!make_layer(Cout, C1, C2, 3, false, L, 1, true, true) || N < 1 || D < 2 || H < 2 || W < 2 || ((D | H | W) & 1)) return 0;
    if ((long)N * D * H * W * L.CC >= (1L << 32) - 256) return 0;
    Geo g;
    return choose_geometry(N, D, H, W / 2, L, 1, g) && g.n_tiles <= 32 ? 1 : 0;
}

extern "C" int rf_tc_conv3d_halo_wp_pool_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N,
                                             int D, int H, int W, int Cout, int C1, int C2, int act, float slope, float out_scale,
                                             void* stream) {
    RF_CHECK_ARG(hi && lo && weight_image && y, "rf_tc_conv3d_halo_wp_pool_fwd: null pointer");
    Layer L;
    RF_CHECK_ARG(make_layer(Cout, C1, C2, 3, false, L, 1, true, true) && N > 0 && D > 1 && H > 1 && W > 1 && !((D | H | W) & 1),
                 "rf_tc_conv3d_halo_wp_pool_fwd: unsupported shape Cout=%d C1=%d C2=%d %dx%dx%d", Cout, C1, C2, D, H, W);
    RF_CHECK_ARG(out_scale > 0.f, "rf_tc_conv3d_halo_wp_pool_fwd: the output scale must be positive (max is taken before it)");
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)weight_image & 15) == 0 && ((uintptr_t)y & 15) == 0,
                 "rf_tc_conv3d_halo_wp_pool_fwd: pointers must be 16-byte aligned");
    return launch_conv(L, hi, lo, weight_image, bias, y, N, D, H, W / 2, 1, act, slope, out_scale, 0, stream);
}

/* Convolution whose epilogue applies the NEXT SingleConv's GroupNorm (model/unet.py:79-144: the first convolution of a
 * DoubleConv, whose output only feeds the second) and writes that layer's operand planes: statistics of the sample's
 * activated output, normalisation, x scale2, fp16 hi / lo split, all from the accumulators in TMEM.  out_hi / out_lo:
 * rf_halo_act_bytes(N, Do, Ho, Wo, Cout, 0, 1) bytes each, in rf_cl_norm_split_halo's layout (out_wp = 0) or
 * rf_cl_norm_split_halo_wp's (out_wp = 1).  Needs an item shape with one whole sample per item (_supported). */
extern "C" int rf_tc_conv3d_halo_gn_supported(int N, int D, int H, int W, int pad, int Cout, int C1, int C2, int groups2) {
    Layer L;
    if (!make_layer(Cout, C1, C2, 3, false, L) || N < 1 || pad < 0 || pad > 1 || groups2 < 1 || Cout % groups2 || Cout > 64) return 0;
    const int Do = D + 2 * pad - 2, Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
    if (Do < 1 || Ho < 1 || Wo < 1) return 0;
    if ((long)N * D * H * W * L.CC >= (1L << 32) - 256) return 0;
    Geo g, g0;
    const bool free_ok = choose_geometry(N, Do, Ho, Wo, L, pad, g0) && g0.n_tiles <= 32;
    L.gn_out = 1;
    if (!choose_geometry(N, Do, Ho, Wo, L, pad, g) || g.n_tiles > 32) return 0;
    // 2: worth it - the one-sample item is (about) what the chooser would take anyway AND the layer has enough MMA work per
    // item (>= 8 channel chunks) to carry an epilogue that visits every accumulator block twice (measured, 16 384 samples
    // of 8^3: 96 -> 56 9.12 ms against 8.25 + 0.52 + 0.70 as separate launches; 16 -> 16 1.08 against 0.72 + 0.18 + 0.18);
    // 1: it runs
    return (!free_ok || g.score >= 0.9 * g0.score) && L.CC >= 8 ? 2 : 1;
}

extern "C" int rf_tc_conv3d_halo_gn_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, int N, int D, int H,
                                        int W, int pad, int Cout, int C1, int C2, int act, float slope, float out_scale,
                                        const float* gn2_w, const float* gn2_b, int groups2, float eps2, float scale2, void* out_hi,
                                        void* out_lo, int out_wp, void* stream) {
    RF_CHECK_ARG(hi && lo && weight_image && gn2_w && gn2_b && out_hi && out_lo, "rf_tc_conv3d_halo_gn_fwd: null pointer");
    Layer L;
    RF_CHECK_ARG(make_layer(Cout, C1, C2, 3, false, L) && N > 0 && (pad == 0 || pad == 1) && groups2 >= 1 && Cout % groups2 == 0 && Cout <= 64,
                 "rf_tc_conv3d_halo_gn_fwd: unsupported shape Cout=%d C1=%d C2=%d pad=%d groups=%d", Cout, C1, C2, pad, groups2);
    L.gn_out = 1;
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)weight_image & 15) == 0 && ((uintptr_t)out_hi & 15) == 0 &&
                 ((uintptr_t)out_lo & 15) == 0, "rf_tc_conv3d_halo_gn_fwd: pointers must be 16-byte aligned");
    D += 2 * pad - 2; H += 2 * pad - 2; W += 2 * pad - 2;
    RF_CHECK_ARG(D > 0 && H > 0 && W > 0 && (!out_wp || (W & 1) == 0), "rf_tc_conv3d_halo_gn_fwd: empty output (or odd W for W-pair planes)");
    GnOut gn{gn2_w, gn2_b, groups2, out_wp ? 1 : 0, eps2, scale2, out_hi, out_lo};
    return launch_conv(L, hi, lo, weight_image, bias, nullptr, N, D, H, W, pad, act, slope, out_scale, 0, stream, 0, 0, 0, &gn);
}

/* Debug / test aid: item shape and cost-model score (outputs per cycle and SM) of the W-pair variant [0] and of the plain
 * layout [1]; out16 = 2 x {stacked, G, Dt, Ht, flags, n_tiles, n_items, smem}. */
extern "C" int rf_tc_conv3d_halo_wp_geometry(int N, int D, int H, int W, int Cout, int C1, int C2, int pad, int* out16, double* scores2) {
    const int Do = D + 2 * pad - 2, Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
    int ok = 0;
    for (int k = 0; k < 2; ++k) {
        Layer L;
        Geo g;
        scores2[k] = -1.0;
        if (!make_layer(Cout, C1, C2, 3, false, L, 1, k == 0) || (k == 0 && (Wo & 1))) continue;
        if (!choose_geometry(N, Do, Ho, k == 0 ? Wo / 2 : Wo, L, pad, g)) continue;
        int* o = out16 + 8 * k;
        o[0] = g.stacked; o[1] = g.G; o[2] = g.Dt; o[3] = g.Ht; o[4] = (g.lines & 1) + 2 * (g.fused != 0) + 4 * (g.halo == 1) + 8 * (g.fused == 2) + 16 * (g.n_sets == 2) + 32 * g.two_resident + 64 * (g.lines == 2);
        o[5] = g.n_tiles; o[6] = g.n_items; o[7] = (int)g.smem;
        scores2[k] = g.score;
        ok |= 1 << k;
    }
    return ok;
}

extern "C" int rf_tc_conv3d_halo_wp_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N,
                                        int D, int H, int W, int pad, int Cout, int C1, int C2, int act, float slope, float out_scale,
                                        void* stream) {
    RF_CHECK_ARG(hi && lo && weight_image && y, "rf_tc_conv3d_halo_wp_fwd: null pointer");
    Layer L;
    RF_CHECK_ARG(make_layer(Cout, C1, C2, 3, false, L, 1, true) && N > 0 && (pad == 0 || pad == 1),
                 "rf_tc_conv3d_halo_wp_fwd: unsupported shape Cout=%d C1=%d C2=%d pad=%d", Cout, C1, C2, pad);
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)weight_image & 15) == 0 && ((uintptr_t)y & 15) == 0,
                 "rf_tc_conv3d_halo_wp_fwd: pointers must be 16-byte aligned");
    D += 2 * pad - 2; H += 2 * pad - 2; W += 2 * pad - 2;  // OUTPUT extents
    RF_CHECK_ARG(D > 0 && H > 0 && W > 0 && (W & 1) == 0, "rf_tc_conv3d_halo_wp_fwd: the output W must be even and positive");
    return launch_conv(L, hi, lo, weight_image, bias, y, N, D, H, W / 2, pad, act, slope, out_scale, 0, stream);
}

// ------------------------------------------------------------------ single-channel layers as W-runs (mode 2)
namespace {
struct WrunArgs {
    const float *x, *mu, *a, *beta;
    uint4 *hi, *lo;
    int W, Wo, pad;
    float scale;
    FastDiv fWo, fH, fD;
    unsigned total;
};

// slot (n, d, h, p), p in [0, Wo): the 8 values x[n, d, h, p - pad + e], e = 0..7 (zero outside the line), normalised
// (single channel: one GroupNorm group) and split into fp16 hi / lo
__global__ void __launch_bounds__(256) cl_norm_split_wrun_kernel(const WrunArgs s) {
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < s.total; j += gridDim.x * blockDim.x) {
        unsigned t = j;
        const int p = (int)fd_divmod(t, s.fWo);
        const unsigned line = t;  // (n * D + d) * H + h
        (void)fd_divmod(t, s.fH);
        (void)fd_divmod(t, s.fD);
        const int n = (int)t;
        const float* src = s.x + (long)line * s.W;
        float m = 0.f, sa = 1.f, sb = 0.f;
        if (s.mu) { m = __ldg(s.mu + n); sa = __ldg(s.a + n); sb = __ldg(s.beta); }
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            float v[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int w = p - s.pad + e + u;
                float val = 0.f;  // zero padding is applied AFTER the normalisation
                if (w >= 0 && w < s.W) {
                    val = __ldg(src + w);
                    if (s.mu) val = fmaf(val - m, sa, sb);
                    val *= s.scale;
                }
                v[u] = val;
            }
            split_f16x2(v[0], v[1], h[e >> 1], l[e >> 1]);
        }
        s.hi[j] = make_uint4(h[0], h[1], h[2], h[3]);
        s.lo[j] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}
}  // namespace

/* Single-channel KS^3 convolution (KS = 3 or 5, stride 1, zero padding pad <= 1) on tensor cores.  D, H, W: INPUT
 * extents; outputs Do = D + 2 pad - KS + 1 etc.  Planes: [N][D][H][Wo] slots. */
extern "C" size_t rf_wrun_act_bytes(int N, int D, int H, int W, int KS, int pad) {
    const int Wo = W + 2 * pad - KS + 1;
    if (N < 1 || D < 1 || H < 1 || Wo < 1 || (KS != 3 && KS != 5) || pad < 0 || pad > 1) return 0;
    return (size_t)N * D * H * Wo * 16;
}

extern "C" int rf_cl_norm_split_wrun(const float* x, const float* gn_mu, const float* gn_a, const float* gn_beta, void* hi, void* lo, int N,
                                     int D, int H, int W, int KS, int pad, float scale, void* stream) {
    RF_CHECK_ARG(x && hi && lo && rf_wrun_act_bytes(N, D, H, W, KS, pad) > 0, "rf_cl_norm_split_wrun: bad arguments");
    RF_CHECK_ARG((gn_mu == nullptr) == (gn_a == nullptr) && (gn_mu == nullptr) == (gn_beta == nullptr), "rf_cl_norm_split_wrun: partial GroupNorm arguments");
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0, "rf_cl_norm_split_wrun: planes must be 16-byte aligned");
    WrunArgs s;
    s.x = x; s.mu = gn_mu; s.a = gn_a; s.beta = gn_beta; s.hi = (uint4*)hi; s.lo = (uint4*)lo;
    s.W = W; s.Wo = W + 2 * pad - KS + 1; s.pad = pad; s.scale = scale;
    const long total = (long)N * D * H * s.Wo;
    RF_CHECK_ARG(total < (1L << 32) - 256, "rf_cl_norm_split_wrun: more than 2^32 slots");
    s.total = (unsigned)total;
    s.fWo = make_fastdiv(s.Wo); s.fH = make_fastdiv(H); s.fD = make_fastdiv(D);
    cl_norm_split_wrun_kernel<<<rf_grid_1d(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(s);
    RF_LAUNCH_OK("cl_norm_split_wrun_kernel");
    return 0;
}

extern "C" size_t rf_tc_conv_wrun_weight_image_bytes(int Cout, int KS) {
    Layer L;
    return make_layer(Cout, 1, 0, KS, true, L) ? weight_image_bytes(L) : 0;
}

extern "C" int rf_tc_conv_wrun_weight_image(const float* w, int Cout, int KS, float scale, void* image, void* stream) {
    Layer L;
    RF_CHECK_ARG(w && image && make_layer(Cout, 1, 0, KS, true, L), "rf_tc_conv_wrun_weight_image: unsupported shape Cout=%d KS=%d", Cout, KS);
    RF_CHECK_ARG(((uintptr_t)image & 15) == 0, "rf_tc_conv_wrun_weight_image: image must be 16-byte aligned");
    return weight_image(w, L, scale, image, stream);
}

extern "C" int rf_tc_conv3d_wrun_supported(int N, int D, int H, int W, int Cout, int KS, int pad) {
    Layer L;
    if (!make_layer(Cout, 1, 0, KS, true, L) || rf_wrun_act_bytes(N, D, H, W, KS, pad) == 0) return 0;
    const int Do = D + 2 * pad - KS + 1, Ho = H + 2 * pad - KS + 1, Wo = W + 2 * pad - KS + 1;
    if (Do < 1 || Ho < 1) return 0;
    Geo g;
    return choose_geometry(N, Do, Ho, Wo, L, pad, g) && g.n_tiles <= 32 ? 1 : 0;
}

extern "C" int rf_tc_conv3d_wrun_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N, int D,
                                     int H, int W, int KS, int pad, int Cout, int act, float slope, float out_scale, int out_ncdhw,
                                     void* stream) {
    RF_CHECK_ARG(hi && lo && weight_image && y, "rf_tc_conv3d_wrun_fwd: null pointer");
    Layer L;
    RF_CHECK_ARG(make_layer(Cout, 1, 0, KS, true, L) && rf_wrun_act_bytes(N, D, H, W, KS, pad) > 0,
                 "rf_tc_conv3d_wrun_fwd: unsupported shape Cout=%d KS=%d pad=%d", Cout, KS, pad);
    RF_CHECK_ARG(((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0 && ((uintptr_t)weight_image & 15) == 0 && ((uintptr_t)y & 15) == 0,
                 "rf_tc_conv3d_wrun_fwd: pointers must be 16-byte aligned");
    const int Do = D + 2 * pad - KS + 1, Ho = H + 2 * pad - KS + 1, Wo = W + 2 * pad - KS + 1;
    RF_CHECK_ARG(Do > 0 && Ho > 0 && Wo > 0, "rf_tc_conv3d_wrun_fwd: empty output");
    return launch_conv(L, hi, lo, weight_image, bias, y, N, Do, Ho, Wo, pad, act, slope, out_scale, out_ncdhw, stream);
}

/* Tuning aid: phase timestamps (clock64) of CTA 0's first six items of the last rf_tc_conv3d_halo_fwd launch:
 * per item [issuer: start, accumulators free, first stage landed, all MMAs issued; epilogue: start waiting,
 * accumulators complete, stores done, -]. */
extern "C" int rf_tc_conv3d_halo_debug_read(long long* out64) {
    RF_CUDA_OK(cudaDeviceSynchronize());
    RF_CUDA_OK(cudaMemcpyFromSymbol(out64, g_halo_dbg, sizeof(long long) * 64));
    return 0;
}
