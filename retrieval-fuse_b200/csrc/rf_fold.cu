// Patch fold / unfold / pad-unfold / recompose / compose / pool / upsample:
// pure re-indexing kernels, HBM-bound.  The fast paths move one float4 per
// thread with consecutive threads walking the contiguous (patch) side, so every
// warp access there is one 512-byte segment and the permuted side is touched in
// whole 32-byte sectors; flat indices are decomposed with an exact
// multiply-high division (FastDiv) or once per thread (pad-unfold); E = 2 is
// transposed through shared memory.  Grids are sized in multiples of the SM
// count (rf_grid_1d).  DESIGN.md 4.4 has the measurements behind these choices.
#include "rf_common.cuh"

// ---------------------------------------------------------------------------
// Unfold3D / Fold3D (model/attention.py:160-188).  One index space for both:
// element (b, px,py,pz, c, ex,ey,ez) of the patch tensor <-> element
// (b, c, px*E+ex, py*E+ey, pz*E+ez) of the volume.
// ---------------------------------------------------------------------------
template <bool kFold>
__global__ void __launch_bounds__(256) fold_unfold_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                                          int C, int R, int E) {
    const long total = (long)B * C * R * R * R * E * E * E;
    const int S = R * E;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        // decompose the PATCH-side linear index
        long t = i;
        const int ez = (int)(t % E); t /= E;
        const int ey = (int)(t % E); t /= E;
        const int ex = (int)(t % E); t /= E;
        const int c = (int)(t % C); t /= C;
        const int pz = (int)(t % R); t /= R;
        const int py = (int)(t % R); t /= R;
        const int px = (int)(t % R); t /= R;
        const int b = (int)t;
        const long v = ((((long)b * C + c) * S + (px * E + ex)) * S + (py * E + ey)) * S + (pz * E + ez);
        if (kFold) out[v] = in[i];
        else out[i] = in[v];
    }
}

// Vectorised variant for E in {4, 8, 16}: one thread moves one float4, consecutive threads walk the PATCH side
// linearly (every warp access there is one contiguous 512-byte segment) and touch the volume side in z-runs of
// E floats (whole 32-byte sectors); two independent float4s are in flight per thread.  The earlier one-thread-per-
// z-run mapping wrote 16 bytes into each of 32 different sectors per store instruction.
template <int E, bool kFold>
__global__ void __launch_bounds__(256) fold_unfold_vec_kernel(const float4* __restrict__ in, float4* __restrict__ out, FastDiv C,
                                                              FastDiv R, unsigned total4) {
    constexpr unsigned E4 = E / 4;
    const unsigned S = R.d * E;
    const unsigned step = gridDim.x * blockDim.x;
    auto vol_index = [&](unsigned i4) -> long {
        const unsigned j = i4 % E4;
        unsigned t = i4 / E4;
        const unsigned ey = t % E; t /= E;
        const unsigned ex = t % E; t /= E;
        const unsigned c = fd_divmod(t, C);
        const unsigned pz = fd_divmod(t, R);
        const unsigned py = fd_divmod(t, R);
        const unsigned px = fd_divmod(t, R);
        const unsigned b = t;
        return (((((long)b * C.d + c) * S + (px * E + ex)) * S + (py * E + ey)) * S + pz * E) / 4 + j;
    };
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + step < total4; i += 2 * step) {
        const long v0 = vol_index(i), v1 = vol_index(i + step);
        const float4 a = __ldg(in + (kFold ? (long)i : v0));
        const float4 b = __ldg(in + (kFold ? (long)(i + step) : v1));
        out[kFold ? v0 : (long)i] = a;
        out[kFold ? v1 : (long)(i + step)] = b;
    }
    if (i < total4) {
        const long v0 = vol_index(i);
        out[kFold ? v0 : (long)i] = __ldg(in + (kFold ? (long)i : v0));
    }
}

// E = 2 (the attention's Unfold3D(2, nf) / Fold3D(16, 2, nf)): one thread moves one (patch, channel) block of
// 2 x 2 x 2 floats: two float4s on the patch side, four float2s on the volume side.
template <bool kFold>
__global__ void __launch_bounds__(256) fold_unfold_e2_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int R,
                                                             int total_blocks) {
    const int S = R * 2;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total_blocks; r += gridDim.x * blockDim.x) {
        int t = r;
        const int c = t % C; t /= C;
        const int pz = t % R; t /= R;
        const int py = t % R; t /= R;
        const int px = t % R;
        const int b = t / R;
        const long v = ((((long)b * C + c) * S + px * 2) * S + py * 2) * S + pz * 2;
        if (kFold) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(in + (long)r * 8));
            const float4 bq = __ldg(reinterpret_cast<const float4*>(in + (long)r * 8 + 4));
            *reinterpret_cast<float2*>(out + v) = make_float2(a.x, a.y);
            *reinterpret_cast<float2*>(out + v + S) = make_float2(a.z, a.w);
            *reinterpret_cast<float2*>(out + v + (long)S * S) = make_float2(bq.x, bq.y);
            *reinterpret_cast<float2*>(out + v + (long)S * S + S) = make_float2(bq.z, bq.w);
        } else {
            const float2 p00 = __ldg(reinterpret_cast<const float2*>(in + v));
            const float2 p01 = __ldg(reinterpret_cast<const float2*>(in + v + S));
            const float2 p10 = __ldg(reinterpret_cast<const float2*>(in + v + (long)S * S));
            const float2 p11 = __ldg(reinterpret_cast<const float2*>(in + v + (long)S * S + S));
            *reinterpret_cast<float4*>(out + (long)r * 8) = make_float4(p00.x, p00.y, p01.x, p01.y);
            *reinterpret_cast<float4*>(out + (long)r * 8 + 4) = make_float4(p10.x, p10.y, p11.x, p11.y);
        }
    }
}

// E = 2 through shared memory: one CTA per (b, px, py) line of R sub-patches.  Its patch side is ONE contiguous
// run of R * C * 8 floats, its volume side C * 4 z-rows of 2R floats, so both global sides move as float4 in
// 128-byte (or longer) contiguous pieces and the (c, ex, ey, ez) <-> (pz, c, ...) transposition happens in shared
// memory.  (The register-only kernel above reads the volume as 8-byte pieces 128 KB apart across the lanes of a warp:
// 20 % of the HBM rate on the attention's Unfold3D(2, nf).)  Patch-order staging with 4 floats of padding per pz
// block: the float2 accesses of the volume side then spread over all banks.
template <bool kFold>
__global__ void __launch_bounds__(256) fold_unfold_e2_smem_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                                  int R) {
    extern __shared__ float4 e2_smem4[];
    float* sm = reinterpret_cast<float*>(e2_smem4);
    const int S = 2 * R, S4 = S / 4, blk = C * 8 + 4;
    const int line = blockIdx.x;                       // (b * R + px) * R + py
    const int py = line % R, px = (line / R) % R, b = line / (R * R);
    const int n_items = 2 * C * R;                     // float4s on either side
    const float* pin = in;
    float* pout = out;
    const long patch_base = (long)line * R * C * 8;    // floats
    const long vol_base = (((long)b * C * S + 2 * px) * S + 2 * py) * S;
    auto vol_addr = [&](int i, int& c, int& ex, int& ey, int& z4) -> long {
        z4 = i % S4;
        const int row = i / S4;
        ey = row & 1; ex = (row >> 1) & 1; c = row >> 2;
        return vol_base + (((long)c * S + ex) * S + ey) * S + 4 * z4;
    };
    if (!kFold) {
        for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
            int c, ex, ey, z4;
            const float4 v = __ldg(reinterpret_cast<const float4*>(pin + vol_addr(i, c, ex, ey, z4)));
            float* d0 = sm + (2 * z4) * blk + c * 8 + ex * 4 + ey * 2;
            *reinterpret_cast<float2*>(d0) = make_float2(v.x, v.y);
            *reinterpret_cast<float2*>(d0 + blk) = make_float2(v.z, v.w);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
            const int j = i % (2 * C), pz = i / (2 * C);
            *reinterpret_cast<float4*>(pout + patch_base + 4L * i) = *reinterpret_cast<const float4*>(sm + pz * blk + 4 * j);
        }
    } else {
        for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
            const int j = i % (2 * C), pz = i / (2 * C);
            *reinterpret_cast<float4*>(sm + pz * blk + 4 * j) = __ldg(reinterpret_cast<const float4*>(pin + patch_base + 4L * i));
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
            int c, ex, ey, z4;
            const long va = vol_addr(i, c, ex, ey, z4);
            const float* s0 = sm + (2 * z4) * blk + c * 8 + ex * 4 + ey * 2;
            const float2 a = *reinterpret_cast<const float2*>(s0);
            const float2 bq = *reinterpret_cast<const float2*>(s0 + blk);
            *reinterpret_cast<float4*>(pout + va) = make_float4(a.x, a.y, bq.x, bq.y);
        }
    }
}

template <bool kFold>
static int launch_fold_unfold(const float* in, float* out, int B, int C, int R, int E, cudaStream_t s) {
    const long total = (long)B * C * R * R * R * E * E * E;
    const bool aligned = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
    if (aligned && total < (1L << 31) && (E == 2 || E == 4 || E == 8 || E == 16)) {
        const int units = (int)(E == 2 ? total / 8 : total / 4);
        if (E == 2) {
            const size_t smem = (size_t)R * (C * 8 + 4) * sizeof(float);
            if ((R & 1) == 0 && smem <= 48 * 1024 && (long)B * R * R < (1L << 31))
                fold_unfold_e2_smem_kernel<kFold><<<(unsigned)(B * R * R), 256, smem, s>>>(in, out, C, R);
            else
                fold_unfold_e2_kernel<kFold><<<rf_grid_1d(units, 256, 148 * 64), 256, 0, s>>>(in, out, C, R, units);
        } else {
            // two float4s per thread and loop trip; 148 x 8 resident CTAs x 2 waves at most
            const int grid = rf_grid_1d((units + 1) / 2, 256, 148 * 16);
            const float4* in4 = reinterpret_cast<const float4*>(in);
            float4* out4 = reinterpret_cast<float4*>(out);
            const FastDiv fc = make_fastdiv(C), fr = make_fastdiv(R);
            switch (E) {
                case 4: fold_unfold_vec_kernel<4, kFold><<<grid, 256, 0, s>>>(in4, out4, fc, fr, (unsigned)units); break;
                case 8: fold_unfold_vec_kernel<8, kFold><<<grid, 256, 0, s>>>(in4, out4, fc, fr, (unsigned)units); break;
                default: fold_unfold_vec_kernel<16, kFold><<<grid, 256, 0, s>>>(in4, out4, fc, fr, (unsigned)units); break;
            }
        }
    } else {
        fold_unfold_kernel<kFold><<<rf_grid_1d(total, 256), 256, 0, s>>>(in, out, B, C, R, E);
    }
    return 0;
}

extern "C" int rf_unfold3d(const float* x, float* out, int B, int C, int S, int E, void* stream) {
    RF_CHECK_ARG(x && out, "rf_unfold3d: null pointer");
    RF_CHECK_ARG(B > 0 && C > 0 && S > 0 && E > 0 && S % E == 0, "rf_unfold3d: bad shape B=%d C=%d S=%d E=%d", B, C, S, E);
    launch_fold_unfold<false>(x, out, B, C, S / E, E, (cudaStream_t)stream);
    RF_LAUNCH_OK("fold_unfold_kernel<unfold>");
    return 0;
}

extern "C" int rf_fold3d(const float* x, float* out, int B, int C, int R, int E, void* stream) {
    RF_CHECK_ARG(x && out, "rf_fold3d: null pointer");
    RF_CHECK_ARG(B > 0 && C > 0 && R > 0 && E > 0, "rf_fold3d: bad shape B=%d C=%d R=%d E=%d", B, C, R, E);
    launch_fold_unfold<true>(x, out, B, C, R, E, (cudaStream_t)stream);
    RF_LAUNCH_OK("fold_unfold_kernel<fold>");
    return 0;
}

// ---------------------------------------------------------------------------
// Unfold3DPadStride / Patcher.__call__ with optional fused normalisation.
// ---------------------------------------------------------------------------
struct Int3 { int v[3]; };

__global__ void __launch_bounds__(256) pad_unfold_kernel(const float* __restrict__ x, float* __restrict__ out, int B,
                                                         int C, Int3 size, Int3 kernel, Int3 pad, Int3 stride, Int3 cnt,
                                                         float pad_val, float norm_sub, float norm_div) {
    const long total = (long)B * cnt.v[0] * cnt.v[1] * cnt.v[2] * C * kernel.v[0] * kernel.v[1] * kernel.v[2];
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long t = i;
        const int kz = (int)(t % kernel.v[2]); t /= kernel.v[2];
        const int ky = (int)(t % kernel.v[1]); t /= kernel.v[1];
        const int kx = (int)(t % kernel.v[0]); t /= kernel.v[0];
        const int c = (int)(t % C); t /= C;
        const int iz = (int)(t % cnt.v[2]); t /= cnt.v[2];
        const int iy = (int)(t % cnt.v[1]); t /= cnt.v[1];
        const int ix = (int)(t % cnt.v[0]); t /= cnt.v[0];
        const int b = (int)t;
        const int sx = ix * stride.v[0] + kx - pad.v[0];
        const int sy = iy * stride.v[1] + ky - pad.v[1];
        const int sz = iz * stride.v[2] + kz - pad.v[2];
        float v = pad_val;
        if (sx >= 0 && sx < size.v[0] && sy >= 0 && sy < size.v[1] && sz >= 0 && sz < size.v[2])
            v = x[((((long)b * C + c) * size.v[0] + sx) * size.v[1] + sy) * size.v[2] + sz];
        if (norm_div != 0.f) v = __fdiv_rn(__fsub_rn(v, norm_sub), norm_div);  // two rounded fp32 ops, as numpy does
        out[i] = v;
    }
}

// Same, one thread per float4 of ONE batch item's output, looping over batch items (kernel z extent a multiple of
// 4, 16-byte aligned output).  Consecutive threads write consecutive 16-byte pieces, so every warp store is one
// contiguous 512-byte segment (the first version - one thread per z-run - scattered each store instruction over 32
// sectors and ran at 16-36 % of the HBM rate).  The position of a float4 inside its item fixes everything but the
// batch index: the index decomposition, the bounds tests and the source offset are computed ONCE per thread, and an
// iteration is load / normalise / store plus two pointer increments.  (The second version decomposed the flat index
// per float4: ~350 instructions each, issue-bound at 1.5-2.9 TB/s.)  The source float4 is loaded in one piece when it
// lies inside the volume and is 16-byte aligned, else element by element with the pad value; the overlapping reads
// (8x for the 32^3 / stride-16 target patches) are L1 / L2 hits.
struct PadUnfoldDivs { FastDiv kz4, ky, kx, c, cz, cy; };

__global__ void __launch_bounds__(256) pad_unfold_item_kernel(const float* __restrict__ x, float4* __restrict__ out, Int3 size,
                                                              Int3 pad, Int3 stride, PadUnfoldDivs dv, float pad_val,
                                                              float norm_sub, float norm_div, float norm_rcp, unsigned item4,
                                                              long item_in, int B, int x_aligned) {
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;  // float4 index inside one item's patches
    if (w >= item4) return;
    const bool norm = norm_div != 0.f;
    const float pad_out = norm ? rf_div_rn_fixed(__fsub_rn(pad_val, norm_sub), norm_div, norm_rcp) : pad_val;
    unsigned t = w;
    const int kz = 4 * (int)fd_divmod(t, dv.kz4);
    const int ky = (int)fd_divmod(t, dv.ky);
    const int kx = (int)fd_divmod(t, dv.kx);
    const int c = (int)fd_divmod(t, dv.c);
    const int iz = (int)fd_divmod(t, dv.cz);
    const int iy = (int)fd_divmod(t, dv.cy);
    const int ix = (int)t;
    const int sx = ix * stride.v[0] + kx - pad.v[0];
    const int sy = iy * stride.v[1] + ky - pad.v[1];
    const int sz = iz * stride.v[2] + kz - pad.v[2];
    const bool row_in = sx >= 0 && sx < size.v[0] && sy >= 0 && sy < size.v[1] && sz > -4 && sz < size.v[2];
    const bool vec = row_in && x_aligned && (size.v[2] & 3) == 0 && (item_in & 3) == 0 && sz >= 0 && sz + 4 <= size.v[2] && (sz & 3) == 0;
    bool in_e[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) in_e[e] = row_in && sz + e >= 0 && sz + e < size.v[2];
    const long off = (((long)c * size.v[0] + (row_in ? sx : 0)) * size.v[1] + (row_in ? sy : 0)) * size.v[2] + sz;
    const float* src = x + (long)blockIdx.y * item_in + off;
    float4* dst = out + (long)blockIdx.y * item4 + w;
    const long src_step = (long)gridDim.y * item_in, dst_step = (long)gridDim.y * item4;
    const float4 padv = make_float4(pad_out, pad_out, pad_out, pad_out);
    if (!row_in) {
        for (int b = blockIdx.y; b < B; b += gridDim.y, dst += dst_step) *dst = padv;
    } else if (vec) {
        int b = blockIdx.y;
        for (; b + (int)gridDim.y < B; b += 2 * gridDim.y, src += 2 * src_step, dst += 2 * dst_step) {  // two loads in flight
            float4 q0 = __ldg(reinterpret_cast<const float4*>(src));
            float4 q1 = __ldg(reinterpret_cast<const float4*>(src + src_step));
            if (norm) {  // two rounded fp32 ops per element, as numpy does
                q0.x = rf_div_rn_fixed(__fsub_rn(q0.x, norm_sub), norm_div, norm_rcp); q0.y = rf_div_rn_fixed(__fsub_rn(q0.y, norm_sub), norm_div, norm_rcp);
                q0.z = rf_div_rn_fixed(__fsub_rn(q0.z, norm_sub), norm_div, norm_rcp); q0.w = rf_div_rn_fixed(__fsub_rn(q0.w, norm_sub), norm_div, norm_rcp);
                q1.x = rf_div_rn_fixed(__fsub_rn(q1.x, norm_sub), norm_div, norm_rcp); q1.y = rf_div_rn_fixed(__fsub_rn(q1.y, norm_sub), norm_div, norm_rcp);
                q1.z = rf_div_rn_fixed(__fsub_rn(q1.z, norm_sub), norm_div, norm_rcp); q1.w = rf_div_rn_fixed(__fsub_rn(q1.w, norm_sub), norm_div, norm_rcp);
            }
            dst[0] = q0;
            dst[dst_step] = q1;
        }
        if (b < B) {
            float4 q0 = __ldg(reinterpret_cast<const float4*>(src));
            if (norm) {
                q0.x = rf_div_rn_fixed(__fsub_rn(q0.x, norm_sub), norm_div, norm_rcp); q0.y = rf_div_rn_fixed(__fsub_rn(q0.y, norm_sub), norm_div, norm_rcp);
                q0.z = rf_div_rn_fixed(__fsub_rn(q0.z, norm_sub), norm_div, norm_rcp); q0.w = rf_div_rn_fixed(__fsub_rn(q0.w, norm_sub), norm_div, norm_rcp);
            }
            dst[0] = q0;
        }
    } else {
        // partly outside or not 16-byte aligned: element loads under per-thread constant predicates, two items in flight
        int b = blockIdx.y;
        for (; b + (int)gridDim.y < B; b += 2 * gridDim.y, src += 2 * src_step, dst += 2 * dst_step) {
            float f[4], g[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                f[e] = in_e[e] ? __ldg(src + e) : 0.f;
                g[e] = in_e[e] ? __ldg(src + src_step + e) : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (norm) {
                    f[e] = rf_div_rn_fixed(__fsub_rn(f[e], norm_sub), norm_div, norm_rcp);
                    g[e] = rf_div_rn_fixed(__fsub_rn(g[e], norm_sub), norm_div, norm_rcp);
                }
                if (!in_e[e]) { f[e] = pad_out; g[e] = pad_out; }
            }
            dst[0] = make_float4(f[0], f[1], f[2], f[3]);
            dst[dst_step] = make_float4(g[0], g[1], g[2], g[3]);
        }
        if (b < B) {
            float f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                f[e] = pad_out;
                if (in_e[e]) {
                    f[e] = __ldg(src + e);
                    if (norm) f[e] = rf_div_rn_fixed(__fsub_rn(f[e], norm_sub), norm_div, norm_rcp);
                }
            }
            *dst = make_float4(f[0], f[1], f[2], f[3]);
        }
    }
}

extern "C" int rf_unfold3d_pad_stride(const float* x, float* out, int B, int C, const int size[3], const int kernel[3],
                                      const int pad[3], const int stride[3], float pad_val, float norm_sub,
                                      float norm_div, void* stream) {
    RF_CHECK_ARG(x && out && size && kernel && pad && stride, "rf_unfold3d_pad_stride: null pointer");
    Int3 s, k, p, st, cnt;
    long total = (long)B * C;
    for (int a = 0; a < 3; ++a) {
        s.v[a] = size[a]; k.v[a] = kernel[a]; p.v[a] = pad[a]; st.v[a] = stride[a];
        RF_CHECK_ARG(size[a] > 0 && kernel[a] > 0 && pad[a] >= 0 && stride[a] > 0, "rf_unfold3d_pad_stride: bad axis %d", a);
        RF_CHECK_ARG(size[a] + 2 * pad[a] >= kernel[a], "rf_unfold3d_pad_stride: kernel larger than padded input");
        cnt.v[a] = (size[a] + 2 * pad[a] - kernel[a]) / stride[a] + 1;  // Tensor.unfold count
        total *= (long)cnt.v[a] * kernel[a];
    }
    RF_CHECK_ARG(B > 0 && C > 0, "rf_unfold3d_pad_stride: bad B/C");
    const long item4 = total / B / 4;  // float4s of one batch item's patches
    if (kernel[2] % 4 == 0 && ((uintptr_t)out & 15) == 0 && item4 < (1L << 31)) {
        PadUnfoldDivs dv;
        dv.kz4 = make_fastdiv(kernel[2] / 4); dv.ky = make_fastdiv(kernel[1]); dv.kx = make_fastdiv(kernel[0]);
        dv.c = make_fastdiv(C); dv.cz = make_fastdiv(cnt.v[2]); dv.cy = make_fastdiv(cnt.v[1]);
        // grid.x covers one item, grid.y strides over the batch: ~148 x 16 CTAs in total, at least 2 items per thread
        const unsigned gx = (unsigned)rf_cdivl(item4, 256);
        long gy = rf_cdivl(148L * 16, gx);
        if (gy > (B + 1) / 2) gy = (B + 1) / 2;
        if (gy < 1) gy = 1;
        if (gy > 65535) gy = 65535;
        pad_unfold_item_kernel<<<dim3(gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(
            x, reinterpret_cast<float4*>(out), s, p, st, dv, pad_val, norm_sub, norm_div, rf_host_rcp_for_div(norm_div), (unsigned)item4,
            (long)C * size[0] * size[1] * size[2], B, ((uintptr_t)x & 15) == 0);
    } else {
        pad_unfold_kernel<<<rf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(x, out, B, C, s, k, p, st, cnt, pad_val,
                                                                                   norm_sub, norm_div);
    }
    RF_LAUNCH_OK("pad_unfold_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// Patcher.recompose_patches as a gather: the reference writes patches in scan
// order (x outer, z inner) so the last writer of a voxel is the patch with the
// largest covering index along every axis.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) recompose_kernel(const float* __restrict__ patches, float* __restrict__ out,
                                                        int B, int C, Int3 size, Int3 kernel, Int3 pad, Int3 stride,
                                                        Int3 cnt, float pad_val) {
    const long total = (long)B * C * size.v[0] * size.v[1] * size.v[2];
    const long n_patches = (long)cnt.v[0] * cnt.v[1] * cnt.v[2];
    const long kvol = (long)kernel.v[0] * kernel.v[1] * kernel.v[2];
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long t = i;
        int co[3];
        co[2] = (int)(t % size.v[2]); t /= size.v[2];
        co[1] = (int)(t % size.v[1]); t /= size.v[1];
        co[0] = (int)(t % size.v[0]); t /= size.v[0];
        t /= C;  // channel: every channel receives the same patch data (broadcast assignment)
        const int b = (int)t;
        int pi[3], off[3];
        bool covered = true;
        for (int a = 0; a < 3; ++a) {
            const int pc = co[a] + pad.v[a];  // coordinate in the padded volume
            int idx = pc / stride.v[a];
            if (idx > cnt.v[a] - 1) idx = cnt.v[a] - 1;
            const int o = pc - idx * stride.v[a];
            if (o >= kernel.v[a]) covered = false;
            pi[a] = idx; off[a] = o;
        }
        float v = pad_val;
        if (covered) {
            const long p = ((long)pi[0] * cnt.v[1] + pi[1]) * cnt.v[2] + pi[2];
            v = patches[((long)b * n_patches + p) * kvol + ((long)off[0] * kernel.v[1] + off[1]) * kernel.v[2] + off[2]];
        }
        out[i] = v;
    }
}

extern "C" int rf_recompose_patches(const float* patches, float* out, int B, int C, const int size[3],
                                    const int kernel[3], const int pad[3], const int stride[3], const int count[3],
                                    float pad_val, void* stream) {
    RF_CHECK_ARG(patches && out && size && kernel && pad && stride && count, "rf_recompose_patches: null pointer");
    Int3 s, k, p, st, cnt;
    long total = (long)B * C;
    for (int a = 0; a < 3; ++a) {
        s.v[a] = size[a]; k.v[a] = kernel[a]; p.v[a] = pad[a]; st.v[a] = stride[a]; cnt.v[a] = count[a];
        RF_CHECK_ARG(size[a] > 0 && kernel[a] > 0 && pad[a] >= 0 && stride[a] > 0 && count[a] > 0,
                     "rf_recompose_patches: bad axis %d", a);
        RF_CHECK_ARG((count[a] - 1) * stride[a] + kernel[a] <= size[a] + 2 * pad[a],
                     "rf_recompose_patches: patches exceed the padded volume on axis %d", a);
        total *= size[a];
    }
    recompose_kernel<<<rf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(patches, out, B, C, s, k, p, st, cnt,
                                                                              pad_val);
    RF_LAUNCH_OK("recompose_kernel");
    return 0;
}

// ---------------------------------------------------------------------------
// compose: out[c,k,dst block] = store[scene][src block] * ratio
// (util/retrieval.py:145-164).  One CTA per (chunk, k, patch).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compose_kernel(const float* __restrict__ rows, const int* __restrict__ dst_ext,
                                                      const float* __restrict__ store, float* __restrict__ out, int P,
                                                      int K, int n_scenes, Int3 ssz, Int3 csz, float trunc, float ratio,
                                                      float norm_sub, float norm_div, float norm_rcp, int patch_major) {
    const int p = blockIdx.x, k = blockIdx.y, c = blockIdx.z;
    const float* row = rows + (((long)c * P + p) * K + k) * 8;
    const int scene = (int)row[0];
    if (scene < -1) return;  // cell without an owner (overlapping compose): the pre-filled truncation value stays
    // .astype(np.int32) on fp32 extents (util/retrieval.py:153)
    const int X0 = (int)row[1], X1 = (int)row[2], Y0 = (int)row[3], Y1 = (int)row[4], Z0 = (int)row[5], Z1 = (int)row[6];
    const int* de = dst_ext + p * 6;
    const int ex = de[1] - de[0], ey = de[3] - de[2], ez = de[5] - de[4];
    // the sentinel block is a float64 numpy array in the reference (:161), so the
    // product is rounded to fp32 only once, on assignment
    const float fill = (float)((double)trunc * (double)ratio);
    // patch_major: out is [chunk, k, patch, ex, ey, ez] - every destination block contiguous, i.e. what Unfold3D(16, 1)
    // (trainer/train_refinement.py:110-111) makes of the composed volume when equal blocks tile it in patch order
    const long cvol = (long)csz.v[0] * csz.v[1] * csz.v[2];
    if (patch_major && (long)ex * ey * ez * P != cvol) return;  // (the host checked the tiling; never write out of the slot)
    float* o = out + ((long)c * K + k) * cvol + (patch_major ? (long)p * ex * ey * ez : 0L);
    const float* s = store + (long)(scene < 0 ? 0 : scene) * ssz.v[0] * ssz.v[1] * ssz.v[2];
    const int n = ex * ey * ez;
    // fast path: the whole block lies inside the scene and every z-run is 16-byte aligned on both sides -> float4 moves
    const bool inside = scene >= 0 && scene < n_scenes && X0 >= 0 && Y0 >= 0 && Z0 >= 0 && X0 + ex <= X1 && Y0 + ey <= Y1 &&
                        Z0 + ez <= Z1 && X1 <= ssz.v[0] && Y1 <= ssz.v[1] && Z1 <= ssz.v[2];
    if (inside && (ez & 3) == 0 && (Z0 & 3) == 0 && (de[4] & 3) == 0 && (ssz.v[2] & 3) == 0 && (csz.v[2] & 3) == 0 &&
        (((uintptr_t)store | (uintptr_t)out) & 15) == 0) {
        const int ez4 = ez >> 2, n4 = ex * ey * ez4;
        const bool pow2 = (ez4 & (ez4 - 1)) == 0 && (ey & (ey - 1)) == 0;
        const int sh_z = 31 - __clz(ez4), sh_zy = sh_z + 31 - __clz(ey);
        // four independent 16-byte loads in flight per thread before the first store (a 16^3 block is 1024 float4s =
        // one trip of this loop): the copy is a dependent chain row -> scene -> data, so latency, not bandwidth, is
        // what a CTA sees
        for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
            float4 v[4];
            long dsti[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i < n4) {
                    int z, y, x;
                    if (pow2) { z = (i & (ez4 - 1)) << 2; y = (i >> sh_z) & (ey - 1); x = i >> sh_zy; }   // 16^3 blocks: no divisions
                    else { z = (i % ez4) << 2; y = (i / ez4) % ey; x = i / (ez4 * ey); }
                    v[u] = __ldg(reinterpret_cast<const float4*>(s + ((long)(X0 + x) * ssz.v[1] + (Y0 + y)) * ssz.v[2] + Z0 + z));
                    dsti[u] = patch_major ? ((long)x * ey + y) * ez + z
                                          : ((long)(de[0] + x) * csz.v[1] + (de[2] + y)) * csz.v[2] + (de[4] + z);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u * blockDim.x < n4) {
                    float4 w = v[u];
                    w.x = __fmul_rn(w.x, ratio); w.y = __fmul_rn(w.y, ratio); w.z = __fmul_rn(w.z, ratio); w.w = __fmul_rn(w.w, ratio);
                    if (norm_div != 0.f) {
                        w.x = rf_div_rn_fixed(__fsub_rn(w.x, norm_sub), norm_div, norm_rcp); w.y = rf_div_rn_fixed(__fsub_rn(w.y, norm_sub), norm_div, norm_rcp);
                        w.z = rf_div_rn_fixed(__fsub_rn(w.z, norm_sub), norm_div, norm_rcp); w.w = rf_div_rn_fixed(__fsub_rn(w.w, norm_sub), norm_div, norm_rcp);
                    }
                    *reinterpret_cast<float4*>(o + dsti[u]) = w;
                }
            }
        }
        return;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int z = i % ez, y = (i / ez) % ey, x = i / (ez * ey);
        float v = fill;
        const int sx = X0 + x, sy = Y0 + y, sz = Z0 + z;
        // numpy slicing clips at the array end; a block that is narrower than
        // the destination would raise in the reference, so only in-range
        // voxels are defined - anything else keeps the truncation value.
        if (scene >= 0 && scene < n_scenes && sx < X1 && sy < Y1 && sz < Z1 && sx >= 0 && sy >= 0 && sz >= 0 &&
            sx < ssz.v[0] && sy < ssz.v[1] && sz < ssz.v[2])
            v = __fmul_rn(s[((long)sx * ssz.v[1] + sy) * ssz.v[2] + sz], ratio);
        if (norm_div != 0.f) v = __fdiv_rn(__fsub_rn(v, norm_sub), norm_div);
        o[patch_major ? (long)i : ((long)(de[0] + x) * csz.v[1] + (de[2] + y)) * csz.v[2] + (de[4] + z)] = v;
    }
}

static int compose_launch(const float* rows, const int* dst_extents, const float* scene_store, float* out, int n_chunks, int P,
                          int K, int n_scenes, const int scene_size[3], const int chunk_size[3], float trunc, float ratio,
                          float norm_sub, float norm_div, int patch_major, void* stream) {
    RF_CHECK_ARG(rows && dst_extents && scene_store && out && scene_size && chunk_size, "rf_compose_gather: null pointer");
    RF_CHECK_ARG(n_chunks > 0 && P > 0 && K > 0 && n_scenes > 0 && K <= 65535 && n_chunks <= 65535,
                 "rf_compose_gather: bad sizes");
    Int3 s, c;
    for (int a = 0; a < 3; ++a) { s.v[a] = scene_size[a]; c.v[a] = chunk_size[a]; }
    dim3 grid(P, K, n_chunks);
    compose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rows, dst_extents, scene_store, out, P, K, n_scenes, s, c,
                                                          trunc, ratio, norm_sub, norm_div, rf_host_rcp_for_div(norm_div),
                                                          patch_major);
    RF_LAUNCH_OK("compose_kernel");
    return 0;
}

extern "C" int rf_compose_gather(const float* rows, const int* dst_extents, const float* scene_store, float* out,
                                 int n_chunks, int P, int K, int n_scenes, const int scene_size[3],
                                 const int chunk_size[3], float trunc, float ratio, float norm_sub,
                                 float norm_div, void* stream) {
    return compose_launch(rows, dst_extents, scene_store, out, n_chunks, P, K, n_scenes, scene_size, chunk_size, trunc, ratio,
                          norm_sub, norm_div, 0, stream);
}

extern "C" int rf_compose_gather_patches(const float* rows, const int* dst_extents, const float* scene_store, float* out,
                                         int n_chunks, int P, int K, int n_scenes, const int scene_size[3],
                                         const int chunk_size[3], float trunc, float ratio, float norm_sub,
                                         float norm_div, void* stream) {
    return compose_launch(rows, dst_extents, scene_store, out, n_chunks, P, K, n_scenes, scene_size, chunk_size, trunc, ratio,
                          norm_sub, norm_div, 1, stream);
}

// ---------------------------------------------------------------------------
// MaxPool3d(2) and nearest 2x upsample
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2_kernel(const float* __restrict__ x, float* __restrict__ y, long NC, int D,
                                                       int H, int W) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    const long total = NC * Do * Ho * Wo;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long t = i;
        const int w = (int)(t % Wo); t /= Wo;
        const int h = (int)(t % Ho); t /= Ho;
        const int d = (int)(t % Do); t /= Do;
        const float* p = x + ((t * D + 2 * d) * H + 2 * h) * (long)W + 2 * w;
        float m = p[0];
        m = fmaxf(m, p[1]);
        m = fmaxf(m, p[W]); m = fmaxf(m, p[W + 1]);
        const float* p2 = p + (long)H * W;
        m = fmaxf(m, p2[0]); m = fmaxf(m, p2[1]);
        m = fmaxf(m, p2[W]); m = fmaxf(m, p2[W + 1]);
        y[i] = m;
    }
}

extern "C" int rf_maxpool3d_2(const float* x, float* y, int N, int C, int D, int H, int W, void* stream) {
    RF_CHECK_ARG(x && y && N > 0 && C > 0 && D >= 2 && H >= 2 && W >= 2, "rf_maxpool3d_2: bad arguments");
    const long total = (long)N * C * (D / 2) * (H / 2) * (W / 2);
    maxpool2_kernel<<<rf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, (long)N * C, D, H, W);
    RF_LAUNCH_OK("maxpool2_kernel");
    return 0;
}

__global__ void __launch_bounds__(256) upsample2_kernel(const float* __restrict__ x, float* __restrict__ y, long NC, int D,
                                                        int H, int W) {
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
    const long total = NC * Do * Ho * Wo;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long t = i;
        const int w = (int)(t % Wo); t /= Wo;
        const int h = (int)(t % Ho); t /= Ho;
        const int d = (int)(t % Do); t /= Do;
        y[i] = x[((t * D + d / 2) * H + h / 2) * (long)W + w / 2];
    }
}

extern "C" int rf_upsample_nearest_2(const float* x, float* y, int N, int C, int D, int H, int W, void* stream) {
    RF_CHECK_ARG(x && y && N > 0 && C > 0 && D > 0 && H > 0 && W > 0, "rf_upsample_nearest_2: bad arguments");
    const long total = (long)N * C * D * H * W * 8;
    upsample2_kernel<<<rf_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, (long)N * C, D, H, W);
    RF_LAUNCH_OK("upsample2_kernel");
    return 0;
}
