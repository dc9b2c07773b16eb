// Exact brute-force kNN over the embedding bank under the canonical rule
//   d(q,x) = sum_i (double(q_i) - double(x_i))^2   (i ascending, no fma)
//   order  = ascending (d, global row id)
// plus the shard merge and the reference's source-scene demotion
// (util/retrieval.py:79-105).
//
// Data layout: bank [n_rows, 64] fp32 row-major (256 B / row, HBM resident,
// L2 resident up to ~490k rows), queries [Q, 64] fp32.  Each warp owns QW = 8
// queries (fp64 copies in shared memory, read as broadcasts) and sweeps its
// bank slice 32 rows at a time - lane l keeps row (base + l) in 64 registers,
// so every bank byte is read once per 8 queries.  The running top-k of a
// query lives ACROSS the warp's lanes: lane l holds the l-th best (d, id)
// pair; an insertion is one ballot + one shuffle-up.  No atomics, no shared
// memory list, fully deterministic.
#include <float.h>
#include <limits.h>

#include "rf_common.cuh"

namespace {

constexpr int D64 = 64;
constexpr int QW = 8;              // queries per warp
constexpr int WARPS_PER_CTA = 4;   // 32 queries per CTA
constexpr int QB = QW * WARPS_PER_CTA;

__device__ __forceinline__ bool cand_less(double d, int i, double d2, int i2) {
    return d < d2 || (d == d2 && i < i2);
}

// Insert (cd, ci) into the warp-distributed sorted list (ld, li); every lane
// calls this with the same (cd, ci).
__device__ __forceinline__ void warp_list_insert(double& ld, int& li, double cd, int ci, int lane) {
    const bool before = cand_less(ld, li, cd, ci);  // my entry stays in front of the new one
    const unsigned mask = __ballot_sync(0xffffffffu, before);
    const int pos = __popc(mask);  // sorted list => `before` is a prefix
    const double ud = __shfl_up_sync(0xffffffffu, ld, 1);
    const int ui = __shfl_up_sync(0xffffffffu, li, 1);
    if (lane == pos) { ld = cd; li = ci; }
    else if (lane > pos) { ld = ud; li = ui; }
}

// Query window of one launch.  Host-sized sweeps pass q_count = nullptr (queries [0, Q)).  The re-check of the
// tensor-core path's unproven queries passes the DEVICE counter instead (no host synchronisation): the launch then
// covers queries [q_first, min(*q_count, q_first + q_cap)) of the q_sel list, CTAs walk the query blocks with a grid
// stride and leave at once when there is nothing to do.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) knn_exact_f64_kernel(const float* __restrict__ bank, long n_rows,
                                                                           long row_offset, const float* __restrict__ q,
                                                                           long Q, const int* __restrict__ q_count,
                                                                           long q_first, long q_cap, int k, int nsplit,
                                                                           const int* __restrict__ q_sel,
                                                                           int* __restrict__ out_idx,
                                                                           double* __restrict__ out_d) {
    __shared__ double qs[QB][D64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long q_lo = 0, q_hi = Q, stride = Q;
    if (q_count) {
        const long n = (long)*q_count;
        q_lo = q_first;
        q_hi = n < q_first + q_cap ? n : q_first + q_cap;
        stride = q_cap;
    }
    const int split = blockIdx.y;
    const long per = ((n_rows + nsplit - 1) / nsplit + 31) / 32 * 32;
    const long r_begin = split * per;
    long r_end = r_begin + per;
    if (r_end > n_rows) r_end = n_rows;

    for (long q_cta = q_lo + (long)blockIdx.x * QB; q_cta < q_hi; q_cta += (long)gridDim.x * QB) {
        __syncthreads();  // the previous block's queries have been consumed
        for (int i = threadIdx.x; i < QB * D64; i += blockDim.x) {
            const long qi = q_cta + i / D64;
            // q_sel (optional) selects a subset of the query matrix: the re-check of unproven queries
            qs[i / D64][i % D64] = qi < q_hi ? (double)q[(q_sel ? (long)q_sel[qi] : qi) * D64 + (i % D64)] : 0.0;
        }
        __syncthreads();

        double ld[QW];
        int li[QW];
        double tau_d[QW];
        int tau_i[QW];
#pragma unroll
        for (int j = 0; j < QW; ++j) { ld[j] = DBL_MAX; li[j] = INT_MAX; tau_d[j] = DBL_MAX; tau_i[j] = INT_MAX; }

        const double(*qw)[D64] = &qs[warp * QW];
        // queries past the end of the list (the re-check of a handful of unproven queries fills a fraction of one CTA):
        // a warp without queries idles, the others skip their empty slots - the fp64 pipe is the bound of this kernel
        const long nq_l = q_hi - q_cta - (long)warp * QW;
        const int nq = nq_l >= QW ? QW : (int)nq_l;  // warp-uniform
        if (nq <= 0) continue;                       // (still meets the barriers at the top of the next trip)
        for (long base = r_begin; base < r_end; base += 32) {
            const long r = base + lane;
            const bool valid = r < r_end;
            // the row is converted to fp64 once and kept in registers for all 8 queries
            double xd[D64];
            const float4* src = reinterpret_cast<const float4*>(bank + (valid ? r : r_begin) * D64);
#pragma unroll
            for (int i = 0; i < D64 / 4; ++i) {
                const float4 v = __ldg(src + i);
                xd[4 * i] = (double)v.x; xd[4 * i + 1] = (double)v.y; xd[4 * i + 2] = (double)v.z; xd[4 * i + 3] = (double)v.w;
            }
            const int gid = (int)(row_offset + r);
#pragma unroll
            for (int j = 0; j < QW; ++j) {
                if (j >= nq) break;
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < D64; ++i) {
                    const double diff = __dsub_rn(qw[j][i], xd[i]);
                    acc = __dadd_rn(acc, __dmul_rn(diff, diff));
                }
                const bool hit = valid && cand_less(acc, gid, tau_d[j], tau_i[j]);
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int src_lane = __ffs(m) - 1;
                    m &= m - 1;
                    const double cd = __shfl_sync(0xffffffffu, acc, src_lane);
                    const int ci = __shfl_sync(0xffffffffu, gid, src_lane);
                    if (cand_less(cd, ci, tau_d[j], tau_i[j])) {  // warp-uniform
                        warp_list_insert(ld[j], li[j], cd, ci, lane);
                        tau_d[j] = __shfl_sync(0xffffffffu, ld[j], k - 1);
                        tau_i[j] = __shfl_sync(0xffffffffu, li[j], k - 1);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < QW; ++j) {
            const long qi = q_cta + warp * QW + j;
            if (qi < q_hi && lane < k) {
                // single-slice sweeps write the final rows directly (through q_sel when given)
                const long oq = (nsplit == 1 && q_sel) ? (long)q_sel[qi] : qi - q_lo;
                const long o = ((long)split * stride + oq) * k + lane;
                out_idx[o] = li[j];
                out_d[o] = ld[j];
            }
        }
    }
}

// One warp per query: fold S sorted k-lists into one under (d, id).
__global__ void __launch_bounds__(128) knn_merge_kernel(const int* __restrict__ parts_idx,
                                                        const double* __restrict__ parts_d, int S, long Q,
                                                        const int* __restrict__ q_count, long q_first, long q_cap, int k,
                                                        const int* __restrict__ q_sel, int* __restrict__ out_idx,
                                                        double* __restrict__ out_d) {
    const int lane = threadIdx.x & 31;
    long n = Q, stride = Q;
    if (q_count) {  // device-sized re-check window [q_first, q_first + q_cap) of the q_sel list; parts are window-relative
        n = (long)*q_count - q_first;
        if (n > q_cap) n = q_cap;
        stride = q_cap;
        q_sel += q_first;
    }
    for (long qi = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; qi < n; qi += ((long)gridDim.x * blockDim.x) >> 5) {
        double ld = DBL_MAX, tau_d = DBL_MAX;
        int li = INT_MAX, tau_i = INT_MAX;
        for (int s = 0; s < S; ++s) {
            const long o = ((long)s * stride + qi) * k + lane;
            const double md = lane < k ? parts_d[o] : DBL_MAX;
            const int mi = lane < k ? parts_idx[o] : INT_MAX;
            for (int t = 0; t < k; ++t) {
                const double cd = __shfl_sync(0xffffffffu, md, t);
                const int ci = __shfl_sync(0xffffffffu, mi, t);
                if (!cand_less(cd, ci, tau_d, tau_i)) break;  // part lists are sorted: the rest is worse (warp-uniform)
                warp_list_insert(ld, li, cd, ci, lane);
                tau_d = __shfl_sync(0xffffffffu, ld, k - 1);
                tau_i = __shfl_sync(0xffffffffu, li, k - 1);
            }
        }
        const long oq = q_sel ? (long)q_sel[qi] : qi;
        if (lane < k) { out_idx[oq * k + lane] = li; out_d[oq * k + lane] = ld; }
    }
}

// util/retrieval.py:93-100 per query: stable partition by "same scene as the
// query", keep K, emit [scene, extents(6), (float)d].
__global__ void __launch_bounds__(256) knn_demote_rows_kernel(const int* __restrict__ idx2k, const double* __restrict__ d2k,
                                                              const float* __restrict__ meta,
                                                              const int* __restrict__ query_scene, long Q, int K2, int K,
                                                              float* __restrict__ out_rows, int* __restrict__ out_idx) {
    const long qi = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    const int qs = query_scene ? query_scene[qi] : -1;
    int taken = 0;
    // pass 0: hits from other scenes (all hits when demotion is off); pass 1: same-scene hits
    for (int pass = 0; pass < 2 && taken < K; ++pass) {
        for (int t = 0; t < K2 && taken < K; ++t) {
            const int id = idx2k[qi * K2 + t];
            const float* mrow = meta + (long)id * 7;
            const bool same = qs >= 0 && mrow[0] == (float)qs;  // database[:,0] == index (fp32 compare, :95)
            if ((pass == 0) == same) continue;
            float* o = out_rows + (qi * K + taken) * 8;
#pragma unroll
            for (int c = 0; c < 7; ++c) o[c] = mrow[c];
            o[7] = (float)d2k[qi * K2 + t];
            if (out_idx) out_idx[qi * K + taken] = id;
            ++taken;
        }
    }
}

}  // namespace

int rf_knn_exact_launch(const float* bank, long n_rows, long row_offset, const float* q, long Q, int k, const int* q_sel,
                        int* out_idx, double* out_d, void* workspace, size_t workspace_bytes, cudaStream_t s);
int rf_knn_exact_nsplit(long Q, long n_rows);

// Number of bank slices for the exact sweep: enough warps to fill the chip.
int rf_knn_exact_nsplit(long Q, long n_rows) {
    const long q_ctas = (Q + QB - 1) / QB;
    long want = (148L * 8 + q_ctas - 1) / q_ctas;  // ~8 CTAs of 4 warps per SM
    const long max_split = (n_rows + 1023) / 1024;   // keep >= 1024 rows per slice
    if (want > max_split) want = max_split;
    if (want > 256) want = 256;
    if (want < 1) want = 1;
    return (int)want;
}

// q_sel == NULL: queries 0..Q-1 of q.  q_sel != NULL: the Q queries q[q_sel[i]], results written to rows q_sel[i].
int rf_knn_exact_launch(const float* bank, long n_rows, long row_offset, const float* q, long Q, int k, const int* q_sel,
                        int* out_idx, double* out_d, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    const int nsplit = rf_knn_exact_nsplit(Q, n_rows);
    long gx = (Q + QB - 1) / QB;
    if (gx > 65535L * 16) gx = 65535L * 16;  // the kernel walks further query blocks with a grid stride
    dim3 grid((unsigned)gx, nsplit);
    if (nsplit == 1) {
        knn_exact_f64_kernel<<<grid, WARPS_PER_CTA * 32, 0, s>>>(bank, n_rows, row_offset, q, Q, nullptr, 0, 0, k, 1, q_sel, out_idx, out_d);
        RF_LAUNCH_OK("knn_exact_f64_kernel");
        return 0;
    }
    const size_t need = (size_t)nsplit * Q * k * (sizeof(int) + sizeof(double));
    RF_CHECK_ARG(workspace && workspace_bytes >= need, "rf_knn_l2_topk: workspace too small (%zu < %zu)", workspace_bytes, need);
    double* pd = (double*)workspace;
    int* pi = (int*)(pd + (size_t)nsplit * Q * k);
    knn_exact_f64_kernel<<<grid, WARPS_PER_CTA * 32, 0, s>>>(bank, n_rows, row_offset, q, Q, nullptr, 0, 0, k, nsplit, q_sel, pi, pd);
    RF_LAUNCH_OK("knn_exact_f64_kernel");
    knn_merge_kernel<<<(unsigned)rf_cdivl(Q * 32, 128), 128, 0, s>>>(pi, pd, nsplit, Q, nullptr, 0, 0, k, q_sel, out_idx, out_d);
    RF_LAUNCH_OK("knn_merge_kernel");
    return 0;
}

// The exact sweep for a DEVICE-sized list of queries (q_sel[0 .. *q_count)): the re-check of the tensor-core path's
// unproven queries, enqueued unconditionally so that no host synchronisation is needed.  The list is covered by
// windows: the first RF_RECHECK_CAP queries in RF_RECHECK_SLICES bank slices (+ merge), so that a handful of queries
// still spreads over the chip, then windows of RF_RECHECK_WINDOW queries in 8 slices.  Every launch leaves at once
// when its window starts beyond *q_count (the common case: nothing, or a few queries, to re-check).
static const long RF_RECHECK_CAP = 2048;
static const int RF_RECHECK_SLICES = 64;
static const long RF_RECHECK_WINDOW = 65536;
static const int RF_RECHECK_WINDOW_SLICES = 8;
size_t rf_knn_recheck_workspace_bytes(int k) {
    const size_t rows = (size_t)RF_RECHECK_SLICES * RF_RECHECK_CAP > (size_t)RF_RECHECK_WINDOW_SLICES * RF_RECHECK_WINDOW
                            ? (size_t)RF_RECHECK_SLICES * RF_RECHECK_CAP : (size_t)RF_RECHECK_WINDOW_SLICES * RF_RECHECK_WINDOW;
    return rows * k * (sizeof(int) + sizeof(double)) + 256;
}
int rf_knn_recheck_launch(const float* bank, long n_rows, long row_offset, const float* q, long Q, int k, const int* q_sel,
                          const int* q_count, int* out_idx, double* out_d, void* workspace, size_t workspace_bytes,
                          cudaStream_t s) {
    RF_CHECK_ARG(workspace && workspace_bytes >= rf_knn_recheck_workspace_bytes(k) - 256, "rf_knn_l2_topk: re-check workspace too small");
    const long max_split = (n_rows + 1023) / 1024;
    if (max_split <= 1) {  // a bank of at most 1024 rows: one slice, the sweep writes the final rows itself
        long gx = (Q + QB - 1) / QB;
        if (gx > 148 * 8) gx = 148 * 8;
        knn_exact_f64_kernel<<<dim3((unsigned)gx, 1), WARPS_PER_CTA * 32, 0, s>>>(bank, n_rows, row_offset, q, Q, q_count, 0, Q, k, 1, q_sel,
                                                                                 out_idx, out_d);
        RF_LAUNCH_OK("knn_exact_f64_kernel(re-check)");
        return 0;
    }
    for (long first = 0; first < Q;) {
        const bool head = first == 0;
        long cap = head ? RF_RECHECK_CAP : RF_RECHECK_WINDOW;
        if (cap > Q - first) cap = Q - first;
        int ns = head ? RF_RECHECK_SLICES : RF_RECHECK_WINDOW_SLICES;
        if (ns > max_split) ns = (int)max_split;
        double* pd = (double*)workspace;
        int* pi = (int*)(pd + (size_t)ns * cap * k);
        dim3 grid((unsigned)((cap + QB - 1) / QB), ns);
        knn_exact_f64_kernel<<<grid, WARPS_PER_CTA * 32, 0, s>>>(bank, n_rows, row_offset, q, Q, q_count, first, cap, k, ns, q_sel, pi, pd);
        RF_LAUNCH_OK("knn_exact_f64_kernel(re-check)");
        knn_merge_kernel<<<(unsigned)rf_cdivl(cap * 32, 128), 128, 0, s>>>(pi, pd, ns, Q, q_count, first, cap, k, q_sel, out_idx, out_d);
        RF_LAUNCH_OK("knn_merge_kernel(re-check)");
        first += cap;
    }
    return 0;
}

extern "C" int rf_knn_merge(const int* parts_idx, const double* parts_d, int S, long Q, int k, int* out_idx,
                            double* out_d, void* stream) {
    RF_CHECK_ARG(parts_idx && parts_d && out_idx && out_d, "rf_knn_merge: null pointer");
    RF_CHECK_ARG(S > 0 && Q > 0 && k > 0 && k <= 32, "rf_knn_merge: bad sizes S=%d Q=%ld k=%d", S, Q, k);
    knn_merge_kernel<<<(unsigned)rf_cdivl(Q * 32, 128), 128, 0, (cudaStream_t)stream>>>(parts_idx, parts_d, S, Q, nullptr, 0, 0, k, nullptr, out_idx, out_d);
    RF_LAUNCH_OK("knn_merge_kernel");
    return 0;
}

extern "C" int rf_knn_demote_rows(const int* idx2k, const double* d2k, const float* meta, const int* query_scene, long Q,
                                  int K2, int K, float* out_rows, int* out_idx, void* stream) {
    RF_CHECK_ARG(idx2k && d2k && meta && out_rows, "rf_knn_demote_rows: null pointer");
    RF_CHECK_ARG(Q > 0 && K > 0 && K2 >= K, "rf_knn_demote_rows: bad sizes Q=%ld K2=%d K=%d", Q, K2, K);
    knn_demote_rows_kernel<<<(unsigned)rf_cdivl(Q, 256), 256, 0, (cudaStream_t)stream>>>(idx2k, d2k, meta, query_scene, Q, K2, K,
                                                                                        out_rows, out_idx);
    RF_LAUNCH_OK("knn_demote_rows_kernel");
    return 0;
}
