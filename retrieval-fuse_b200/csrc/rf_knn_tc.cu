// tcgen05 candidate pass for the exact kNN (methods 2 and 3 of rf_knn_l2_topk).
//
// Idea: the canonical ranking is by the fp64 distance, but almost all of the
// 2*Q*N*64 flops are only needed to REJECT rows.  So:
//   1. evaluate an approximate score s~ ~ q.x on the 5th-gen tensor cores
//      (tcgen05.mma, fp32 accumulators in TMEM) with a proven error bound
//      |s~ - q.x| <= eps_rel |q| |x|:
//        method 2: one fp16 GEMM, K = 64          (eps_rel = 1.0e-3)
//        method 3: bf16 hi/lo split, K = 3*64: q_hi.x_hi + q_hi.x_lo + q_lo.x_hi
//                  (x = hi + lo + r, |r| <= 2^-18 |x|; eps_rel = 4e-5)
//   2. each query keeps its CAND best s~ per bank slice (a sorted list in shared
//      memory owned by the epilogue thread that reads the query's TMEM lane; hits
//      are inserted cooperatively by the warp when few lanes have one, by every
//      lane for itself when many do).  CAND = 16 for k <= 8, 32 above: the list must be LONGER than k,
//      otherwise the proof below compares the k-th candidate with itself.
//      k > 16 additionally splits the bank into >= 2 slices (2 x 32 candidates).
//   3. re-rank the candidates with the canonical fp64 arithmetic, and PROVE
//      completeness: every rejected row has s~ <= tau (the CAND-th best of its
//      slice), hence d >= |q|^2 + min|x|^2 - 2 (tau + eps); if the k-th best
//      candidate's exact d is strictly below that bound, no rejected row can
//      enter or tie the top-k.  Queries that cannot be proven (dense ties,
//      duplicates beyond the slack) are re-done by the exact fp64 sweep
//      (rf_knn.cu), which is enqueued unconditionally and sized by a DEVICE
//      counter: the call never synchronises the stream.
// The result is therefore bit-identical to method 1 by construction.
//
// Operand staging: both operands are pre-arranged in HBM as byte images of
// the shared-memory tiles the MMA reads - [tile][kb][128 rows][128 B], K-major,
// 128-byte swizzle (16-byte chunk index XOR row%8) - so that one
// cp.async.bulk (TMA bulk copy, 16 KiB) per K-block lands a ready-to-use
// SWIZZLE_128B operand; no tensor map, no register staging.
//
// CTA = 256 queries (two M=128 sub-tiles, so every bank byte pulled from L2
// feeds 2x128 rows of MMA) x a slice of bank tiles; 12 warps:
//   warp 0  producer   : bulk copies into a ring of bank-tile stages
//   warp 1  MMA issuer : one thread, 2 x KBLK x 4 tcgen05.mma (M128 N128 K16) per tile
//   warp 2  TMEM alloc : 512 columns = 2 buffers x 2 sub-tiles x 128 columns
//   warps 4-11 epilogue: tcgen05.ld 32 columns at a time, running top-16
// smem<->MMA and MMA<->epilogue hand-offs are mbarriers (full/empty rings).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#include <cub/device/device_radix_sort.cuh>

#include "rf_common.cuh"

size_t rf_knn_recheck_workspace_bytes(int k);
int rf_knn_recheck_launch(const float* bank, long n_rows, long row_offset, const float* q, long Q, int k, const int* q_sel,
                          const int* q_count, int* out_idx, double* out_d, void* workspace, size_t workspace_bytes,
                          cudaStream_t s);

namespace {

constexpr int TILE = 128;              // rows per operand tile (UMMA M and N)
constexpr int MQ = 2;                  // query sub-tiles per CTA
constexpr int KB_BYTES = TILE * 128;   // one K block: 64 x 16-bit per row, 16 KiB
constexpr int NACC = 2;                // TMEM accumulator buffers (MQ x 128 columns each)
constexpr int MAX_SPLIT = 8;
constexpr int NTHREADS = 128 + MQ * 128;

template <int KBLK, int CAND = 16> struct Cfg {
    static constexpr int TILE_BYTES = KBLK * KB_BYTES;
    static constexpr int NSTAGE = KBLK == 1 ? 6 : 2;
    // query sub-tile in shared memory: the bf16 split's K blocks are (hi, hi, lo) - the duplicate hi block is staged once
    static constexpr int A_BLOCKS = KBLK == 3 ? 2 : 1;
    static constexpr int A_BYTES = A_BLOCKS * KB_BYTES;
    // per epilogue warp: 32 candidate lists (score, position) + a 32-float staging row for the cooperative insertion
    static constexpr int LIST_BYTES = 32 * CAND * 8 + 128;
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + A_BYTES * MQ + TILE_BYTES * NSTAGE + 256 /*barriers*/ + 8 * LIST_BYTES;
};
// bound on |s~ - q.x| / (|q| |x|), see DESIGN.md "kNN proof"
// (+4e-6: the epilogue tags scores with their column in the low 5 mantissa bits)
__host__ __device__ constexpr float eps_rel(int kblk) { return kblk == 1 ? 1.004e-3f : 4.4e-5f; }

struct Stats {            // first 256 bytes of the call workspace
    int n_flagged;        // queries whose top-k could not be proven
    unsigned max_err_bits;  // max observed |s~ - s| over candidates (float bits)
};
struct BankStats {        // first 256 bytes of the prepared bank image
    unsigned nmin_bits;   // min |x|^2 over the bank (float bits)
    unsigned nmax_bits;   // max |x|^2
    int has_perm;         // rows of the image are in scan order (perm[] follows the header)
};

// ---------------------------------------------------------------- PTX wrappers
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
// try_wait suspends the thread for a hardware time slice per attempt; a
// pipeline bug must never hang the GPU, so give up (trap) after ~seconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 26)) {
            printf("rf_knn_tc: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
                   threadIdx.x, bar, parity);
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
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// Issue a 32-lane x 32-column TMEM load; the registers are valid after tc_ld_wait().
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
          "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
          "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory operand descriptor (sm_100 UMMA):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for
//   swizzled K-major; 1) | [32,46) stride byte offset >> 4 = 1024 B between
//   8-row groups | [46,48) version = 1 | [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32 (bits 4-5 = 1), A = B = bf16 (bits
// 7-9, 10-12 = 1), both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
// (A/B format 0 = f16, 1 = bf16)
__host__ __device__ constexpr uint32_t idesc(int ab_format) {
    return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) | ((uint32_t)(TILE >> 3) << 17) |
           ((uint32_t)(TILE >> 4) << 24);
}

// ------------------------------------------------------------------ scan order
// The epilogue's running top-16 pays for every row that beats the current 16th best score (insertion path, taken
// by the whole warp when any of its 32 queries inserts).  Scanning the rows that score high for a TYPICAL query
// first makes the thresholds tight after a few tiles, so later insertions become rare.  The order is the bank
// rows' projection on the mean query direction, descending; it changes nothing in the result (the candidate
// lists carry scan positions, the re-rank maps them back through perm[] and ranks by the canonical fp64 rule).
__global__ void __launch_bounds__(256) knn_mean_query_kernel(const float* __restrict__ q, long Q, float* __restrict__ mean64) {
    __shared__ float sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 0.f;
    __syncthreads();
    const int c = threadIdx.x & 63;
    float acc = 0.f;
    for (long r = blockIdx.x * 4L + (threadIdx.x >> 6); r < Q; r += gridDim.x * 4L) acc += __ldg(q + r * 64 + c);
    atomicAdd(&sh[c], acc);
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(mean64 + threadIdx.x, sh[threadIdx.x]);
}

__global__ void __launch_bounds__(256) knn_scan_key_kernel(const float* __restrict__ bank, long n_rows, long n_padded,
                                                           const float* __restrict__ mean64, float* __restrict__ key,
                                                           int* __restrict__ idx) {
    const long row = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (row >= n_padded) return;
    float k = -FLT_MAX;  // padding rows sort last
    if (row < n_rows) {
        k = 0.f;
        const float4* xr = reinterpret_cast<const float4*>(bank + row * 64);
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const float4 x = __ldg(xr + i);
            k = fmaf(x.x, __ldg(mean64 + 4 * i), k);
            k = fmaf(x.y, __ldg(mean64 + 4 * i + 1), k);
            k = fmaf(x.z, __ldg(mean64 + 4 * i + 2), k);
            k = fmaf(x.w, __ldg(mean64 + 4 * i + 3), k);
        }
    }
    key[row] = k;
    idx[row] = (int)row;
}

// ------------------------------------------------------------------ prep
// src [n_rows, 64] fp32 -> swizzled 16-bit tile images [(n_tiles), KBLK, 128 rows, 128 B].
// KBLK = 1: fp16(x).  KBLK = 3: bf16 split, bank K blocks = (hi, lo, hi), queries (hi, hi, lo).
// For the bank the range of |x|^2 is recorded.
template <int KBLK>
__global__ void __launch_bounds__(256) knn_tc_prep_kernel(const float* __restrict__ src, long n_rows, long n_rows_padded,
                                                          int is_bank, uint8_t* __restrict__ img, BankStats* stats,
                                                          const int* __restrict__ perm) {
    const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;  // one thread per (row, 16-byte chunk)
    const long row = gid >> 3;
    const int c = (int)(gid & 7);
    float nrm = 0.f;
    bool valid = false;
    if (row < n_rows_padded) {
        float x[8];
        const long srow = perm ? (long)perm[row] : row;  // image position `row` holds bank row perm[row] (scan order)
        valid = srow < n_rows;
        if (valid) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src + srow * 64 + c * 8));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + srow * 64 + c * 8 + 4));
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 0.f;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (KBLK == 1) {
                hi[i] = (uint32_t)__half_as_ushort(__float2half_rn(x[2 * i])) |
                        ((uint32_t)__half_as_ushort(__float2half_rn(x[2 * i + 1])) << 16);
                lo[i] = 0u;
            } else {
                const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
                hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            nrm = fmaf(x[2 * i], x[2 * i], nrm);
            nrm = fmaf(x[2 * i + 1], x[2 * i + 1], nrm);
        }
        const long tile = row / TILE;
        const int r = (int)(row % TILE);
        const uint4 vhi = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        uint8_t* base = img + tile * (long)(KBLK * KB_BYTES) + (long)r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(base) = vhi;
        if (KBLK == 3) {
            const uint4 vlo = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(base + KB_BYTES) = is_bank ? vlo : vhi;
            *reinterpret_cast<uint4*>(base + 2 * KB_BYTES) = is_bank ? vhi : vlo;
        }
    }
    if (is_bank) {  // |x|^2 per row: reduce over the row's 8 chunk threads (consecutive lanes)
        nrm += __shfl_xor_sync(0xffffffffu, nrm, 1);
        nrm += __shfl_xor_sync(0xffffffffu, nrm, 2);
        nrm += __shfl_xor_sync(0xffffffffu, nrm, 4);
        if (c == 0 && valid) {  // non-negative floats order like their bit patterns
            atomicMin(&stats->nmin_bits, __float_as_uint(nrm));
            atomicMax(&stats->nmax_bits, __float_as_uint(nrm));
        }
    }
}

// ------------------------------------------------------------------ main kernel
// COOP = 1: lists in shared memory, hits inserted cooperatively by the warp, one hot lane at a time.
// COOP = 0: lists in registers, every lane inserts its own hits with a parallel shift-insert network.
template <int KBLK, int CAND, int COOP>
__global__ void __launch_bounds__(NTHREADS, 1) knn_tc_candidates_kernel(const uint8_t* __restrict__ q_img,
                                                                        const uint8_t* __restrict__ bank_img, long Q,
                                                                        long n_rows, int n_qtiles, int n_btiles,
                                                                        int nsplit, float* __restrict__ cand_s,
                                                                        int* __restrict__ cand_i, int coop_max) {
    using C = Cfg<KBLK, CAND>;
    constexpr int TILE_BYTES = C::TILE_BYTES;
    constexpr int NSTAGE = C::NSTAGE;
    constexpr int A_BYTES = C::A_BYTES;
    constexpr uint32_t IDESC = idesc(KBLK == 1 ? 0 : 1);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
    const uint32_t sA = smem_base;                          // MQ query sub-tiles
    const uint32_t sB = smem_base + MQ * A_BYTES;           // NSTAGE bank tiles
    const uint32_t bars = sB + TILE_BYTES * NSTAGE;
    const uint32_t bar_full = bars;                     // NSTAGE x 8 B
    const uint32_t bar_empty = bars + 8 * NSTAGE;       // NSTAGE x 8 B
    const uint32_t bar_a = bars + 16 * NSTAGE;          // 8 B
    const uint32_t bar_tfull = bar_a + 8;               // NACC x 8 B
    const uint32_t bar_tempty = bar_tfull + 8 * NACC;   // NACC x 8 B
    const uint32_t tmem_slot = bar_tempty + 8 * NACC;   // 4 B
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    uint8_t* list_base = smem_raw + (bars + 256 - smem_u32(smem_raw));  // 8 x LIST_BYTES, 16-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // slice `split` owns bank tiles split, split + nsplit, ...: with the image in scan order every slice sees the
    // promising rows first, and the slices' lists have statistically the same tail (the proof uses their maximum)
    const int qpair = blockIdx.x, split = blockIdx.y;
    const int n_iter = n_btiles > split ? (n_btiles - split + nsplit - 1) / nsplit : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_a, 1);
        for (int b = 0; b < NACC; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, MQ * 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // TMEM: all 512 columns (one CTA per SM by launch bounds)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0) {  // ---------------- producer
            // the query pair: sub-tile m of this CTA is query tile 2*qpair+m (the image is padded to an even tile count)
            mbar_arrive_expect_tx(bar_a, MQ * A_BYTES);
            for (int m = 0; m < MQ; ++m)
                for (int j = 0; j < C::A_BLOCKS; ++j)  // image K blocks (hi, hi, lo): blocks 0 and 2 are staged
                    bulk_g2s(sA + m * A_BYTES + j * KB_BYTES,
                             q_img + ((long)qpair * MQ + m) * TILE_BYTES + (long)(KBLK == 3 ? 2 * j : j) * KB_BYTES, KB_BYTES, bar_a);
            for (int it = 0; it < n_iter; ++it) {
                const int s = it % NSTAGE;
                const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);  // a fresh barrier passes the parity-1 wait
                mbar_arrive_expect_tx(bar_full + 8 * s, TILE_BYTES);
                const uint8_t* src = bank_img + (long)(split + it * nsplit) * TILE_BYTES;
                for (int kb = 0; kb < KBLK; ++kb)
                    bulk_g2s(sB + s * TILE_BYTES + kb * KB_BYTES, src + (long)kb * KB_BYTES, KB_BYTES, bar_full + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---------------- MMA issuer
            mbar_wait(bar_a, 0);
            for (int it = 0; it < n_iter; ++it) {
                const int s = it % NSTAGE, b = it % NACC;
                const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u, bph = (uint32_t)(it / NACC) & 1u;
                mbar_wait(bar_tempty + 8 * b, bph ^ 1u);  // epilogue has drained this accumulator buffer
                mbar_wait(bar_full + 8 * s, ph);          // bank tile landed
                tc_fence_after();
#pragma unroll
                for (int m = 0; m < MQ; ++m) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)((b * MQ + m) * TILE);
#pragma unroll
                    for (int kb = 0; kb < KBLK; ++kb) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {  // 4 x K=16 (32 B) steps inside one 128 B swizzle atom
                            const uint64_t ad = umma_desc(sA + m * A_BYTES + (KBLK == 3 ? (kb >> 1) : kb) * KB_BYTES + k * 32);
                            const uint64_t bd = umma_desc(sB + s * TILE_BYTES + kb * KB_BYTES + k * 32);
                            tc_mma_bf16(d_tmem, ad, bd, IDESC, (kb | k) ? 1u : 0u);
                        }
                    }
                }
                tc_commit(bar_empty + 8 * s);   // smem stage reusable once these MMAs retire
                tc_commit(bar_tfull + 8 * b);   // accumulators ready for the epilogue
            }
        }
    } else if (warp >= 4) {  // ---------------- epilogue: thread <-> query row (TMEM lane)
        const int ew = warp - 4;          // 0..7
        const int quad = ew & 3;          // a warp may only touch TMEM lanes [32*(warp%4), +32); warp%4 == ew%4
        const int m = ew >> 2;            // query sub-tile
      if constexpr (COOP) {
        // Candidate lists live in shared memory, entry-major and XOR-swizzled (entry i of query l at word
        // i * 32 + (l ^ i): conflict-free both when every lane walks its own list and when CAND lanes hold one list),
        // sorted descending; a thread keeps only tau, the CAND-th best score of ITS query, in a register.
        float* ls_sm = reinterpret_cast<float*>(list_base + ew * C::LIST_BYTES);
        int* li_sm = reinterpret_cast<int*>(ls_sm + 32 * CAND);
        float* stage = reinterpret_cast<float*>(li_sm + 32 * CAND);
        for (int t = lane; t < 32 * CAND; t += 32) { ls_sm[t] = -FLT_MAX; li_sm[t] = -1; }
        __syncwarp();
        float tau = -FLT_MAX;
        // Software pipeline over the tile's four 32-column chunks: the TMEM load of
        // chunk c+1 is in flight while chunk c is scanned (two register buffers).
        float va[32], vb[32];
        for (int it = 0; it < n_iter; ++it) {
            const int b = it % NACC;
            const uint32_t bph = (uint32_t)(it / NACC) & 1u;
            mbar_wait(bar_tfull + 8 * b, bph);
            tc_fence_after();
            const long col_tile = (long)(split + it * nsplit) * TILE;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((b * MQ + m) * TILE);
            tc_ld32_issue(taddr, va);
#pragma unroll
            for (int c = 0; c < TILE / 32; ++c) {
                float(&v)[32] = (c & 1) ? vb : va;
                float(&vn)[32] = (c & 1) ? va : vb;
                tc_ld_wait();
                if (c + 1 < TILE / 32) tc_ld32_issue(taddr + (uint32_t)((c + 1) * 32), vn);
                const long col0 = col_tile + c * 32;
                const bool ragged = col0 + 32 > n_rows;
                if (ragged) {  // padded rows of the last tile never compete
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j >= n_rows) v[j] = -FLT_MAX;
                }
                float mx = fmaxf(v[0], v[1]);
#pragma unroll
                for (int j = 2; j < 32; j += 2) mx = fmaxf(mx, fmaxf(v[j], v[j + 1]));
                uint32_t hot = __ballot_sync(0xffffffffu, mx > tau);
                if (hot == 0u) continue;
                // Rare path, executed by the whole warp when ANY of its 32 queries has a score above its threshold.
                // Per-thread insertion (tag the 32 scores with their columns, pull maxima, shift-insert) costs every
                // lane of the warp ~350 instructions for what is usually ONE lane's single hit; on banks without
                // structure ~38 % of the chunks come here and the kernel ran 2.4x slower.  So:
                //  * few hot lanes (the long sparse tail of a scan): the warp serves one hot lane at a time
                //    COOPERATIVELY - the lane drops its 32 scores into shared memory, lane j looks at column j, and a
                //    hit is inserted by CAND lanes at once (position = number of entries >= score, the neighbours
                //    shift by one shuffle): ~60 instructions per hot lane;
                //  * many hot lanes (the first tiles of a scan): every lane inserts its own hits in parallel, list in
                //    registers for the duration (shift-insert network, no carried dependency).
                if (__popc(hot) <= coop_max) {
                    while (hot) {
                        const int L = __ffs(hot) - 1;
                        hot &= hot - 1;
                        if (lane == L) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stage + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                        __syncwarp();
                        const float sj = stage[lane];
                        float tauL = __shfl_sync(0xffffffffu, tau, L);
                        uint32_t hm = __ballot_sync(0xffffffffu, sj > tauL);
                        const bool in = lane < CAND;
                        const int le = in ? lane : 0;
                        float* lsL = ls_sm + le * 32 + (L ^ le);
                        int* liL = li_sm + le * 32 + (L ^ le);
                        while (hm) {
                            const int j = __ffs(hm) - 1;
                            hm &= hm - 1;
                            const float sc = __shfl_sync(0xffffffffu, sj, j);
                            if (!(sc > tauL)) continue;  // the threshold rose since the mask was taken (warp-uniform)
                            const float e = in ? *lsL : -FLT_MAX;
                            const int ei = in ? *liL : -1;
                            const int pos = __popc(__ballot_sync(0xffffffffu, in && e >= sc));  // after its equals; < CAND as sc > tau
                            const float eu = __shfl_up_sync(0xffffffffu, e, 1);
                            const int eiu = __shfl_up_sync(0xffffffffu, ei, 1);
                            const float ne = lane < pos ? e : (lane == pos ? sc : eu);
                            const int nei = lane < pos ? ei : (lane == pos ? (int)(col0 + j) : eiu);
                            if (in) { *lsL = ne; *liL = nei; }
                            tauL = __shfl_sync(0xffffffffu, ne, CAND - 1);
                        }
                        if (lane == L) tau = tauL;
                        __syncwarp();
                    }
                } else if (mx > tau) {
                    float ls[CAND];
                    int li[CAND];
#pragma unroll
                    for (int i = 0; i < CAND; ++i) { ls[i] = ls_sm[i * 32 + (lane ^ i)]; li[i] = li_sm[i * 32 + (lane ^ i)]; }
                    float key[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float t = __uint_as_float((__float_as_uint(v[j]) & ~31u) | (uint32_t)j);
                        key[j] = (ragged && v[j] == -FLT_MAX) ? -FLT_MAX : t;  // masked columns can never beat tau
                    }
#pragma unroll 1
                    for (int guard = 0; guard < 32; ++guard) {
                        float km = fmaxf(key[0], key[1]);
#pragma unroll
                        for (int j = 2; j < 32; j += 2) km = fmaxf(km, fmaxf(key[j], key[j + 1]));
                        if (!(km > tau)) break;
                        const int ci = (int)(col0 + (long)(__float_as_uint(km) & 31u));
                        bool pb[CAND];
#pragma unroll
                        for (int i = 0; i < CAND; ++i) pb[i] = km > ls[i];
#pragma unroll
                        for (int i = CAND - 1; i >= 1; --i) {  // descending i: ls[i - 1] is still the old value
                            ls[i] = pb[i] ? (pb[i - 1] ? ls[i - 1] : km) : ls[i];
                            li[i] = pb[i] ? (pb[i - 1] ? li[i - 1] : ci) : li[i];
                        }
                        ls[0] = pb[0] ? km : ls[0];
                        li[0] = pb[0] ? ci : li[0];
                        tau = ls[CAND - 1];
#pragma unroll
                        for (int j = 0; j < 32; ++j) key[j] = (key[j] == km) ? -FLT_MAX : key[j];
                    }
#pragma unroll
                    for (int i = 0; i < CAND; ++i) { ls_sm[i * 32 + (lane ^ i)] = ls[i]; li_sm[i * 32 + (lane ^ i)] = li[i]; }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * b);
        }
        __syncwarp();
        const long q0 = ((long)qpair * MQ + m) * TILE + quad * 32;
        float* os = cand_s + ((long)split * Q + q0) * CAND;
        int* oi = cand_i + ((long)split * Q + q0) * CAND;
        for (int t = lane; t < 32 * CAND; t += 32) {
            const int ql = t / CAND, i = t % CAND;
            if (q0 + ql < Q) { os[t] = ls_sm[i * 32 + (ql ^ i)]; oi[t] = li_sm[i * 32 + (ql ^ i)]; }
        }
      } else {
        float ls[CAND];
        int li[CAND];
#pragma unroll
        for (int i = 0; i < CAND; ++i) { ls[i] = -FLT_MAX; li[i] = -1; }
        float tau = -FLT_MAX;
        float va[32], vb[32];
        for (int it = 0; it < n_iter; ++it) {
            const int b = it % NACC;
            const uint32_t bph = (uint32_t)(it / NACC) & 1u;
            mbar_wait(bar_tfull + 8 * b, bph);
            tc_fence_after();
            const long col_tile = (long)(split + it * nsplit) * TILE;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((b * MQ + m) * TILE);
            tc_ld32_issue(taddr, va);
#pragma unroll
            for (int c = 0; c < TILE / 32; ++c) {
                float(&v)[32] = (c & 1) ? vb : va;
                float(&vn)[32] = (c & 1) ? va : vb;
                tc_ld_wait();
                if (c + 1 < TILE / 32) tc_ld32_issue(taddr + (uint32_t)((c + 1) * 32), vn);
                const long col0 = col_tile + c * 32;
                const bool ragged = col0 + 32 > n_rows;
                if (ragged) {  // padded rows of the last tile never compete
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j >= n_rows) v[j] = -FLT_MAX;
                }
                float mx = fmaxf(v[0], v[1]);
#pragma unroll
                for (int j = 2; j < 32; j += 2) mx = fmaxf(mx, fmaxf(v[j], v[j + 1]));
                if (mx > tau) {
                    // Rare path, ONE code site (the whole warp takes it when ANY of its 32 queries has a hit): tag every
                    // score with its column (low 5 mantissa bits; costs 2^-18 relative, covered by the proof's eps),
                    // then repeatedly pull the maximum while it beats tau.  Every lane extracts ITS OWN maximum per
                    // trip, so the trip count is the largest number of hits of any one query in this chunk.
                    // The insertion is a shift-insert NETWORK: all CAND comparisons against the new score are
                    // independent, entry i becomes (score beats i) ? ((score beats i-1) ? old i-1 : score) : old i.
                    // The first version carried the displaced element through CAND dependent compare-exchange steps
                    // - a ~250-cycle latency chain per hit that two warps per scheduler cannot hide (banks without
                    // structure, where ~38 % of the chunks come here, ran 2.4x slower).
                    float key[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float t = __uint_as_float((__float_as_uint(v[j]) & ~31u) | (uint32_t)j);
                        key[j] = (ragged && v[j] == -FLT_MAX) ? -FLT_MAX : t;  // masked columns can never beat tau
                    }
#pragma unroll 1
                    for (int guard = 0; guard < 32; ++guard) {
                        float km = fmaxf(key[0], key[1]);
#pragma unroll
                        for (int j = 2; j < 32; j += 2) km = fmaxf(km, fmaxf(key[j], key[j + 1]));
                        if (!(km > tau)) break;
                        const int ci = (int)(col0 + (long)(__float_as_uint(km) & 31u));
                        bool pb[CAND];
#pragma unroll
                        for (int i = 0; i < CAND; ++i) pb[i] = km > ls[i];
#pragma unroll
                        for (int i = CAND - 1; i >= 1; --i) {  // descending i: ls[i - 1] is still the old value
                            ls[i] = pb[i] ? (pb[i - 1] ? ls[i - 1] : km) : ls[i];
                            li[i] = pb[i] ? (pb[i - 1] ? li[i - 1] : ci) : li[i];
                        }
                        ls[0] = pb[0] ? km : ls[0];
                        li[0] = pb[0] ? ci : li[0];
                        tau = ls[CAND - 1];
#pragma unroll
                        for (int j = 0; j < 32; ++j) key[j] = (key[j] == km) ? -FLT_MAX : key[j];
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * b);
        }
        const long q = ((long)qpair * MQ + m) * TILE + quad * 32 + lane;
        if (q < Q) {
            float* os = cand_s + ((long)split * Q + q) * CAND;
            int* oi = cand_i + ((long)split * Q + q) * CAND;
#pragma unroll
            for (int i = 0; i < CAND; ++i) { os[i] = ls[i]; oi[i] = li[i]; }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------ re-rank + proof
__device__ __forceinline__ bool cand_less(double d, int i, double d2, int i2) { return d < d2 || (d == d2 && i < i2); }

// Bitonic sort of LW (d, id) entries, one per lane of an LW-lane group (LW = 16: half a warp, 32: a warp), ascending
// under the canonical order (log^2 compare-exchange stages instead of LW serial list insertions).
template <int LW>
__device__ __forceinline__ void group_bitonic_sort(double& d, int& i, int gl) {
#pragma unroll
    for (int k = 2; k <= LW; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, d, j, LW);
            const int oi = __shfl_xor_sync(0xffffffffu, i, j, LW);
            const bool up = (gl & k) == 0, lower = (gl & j) == 0;
            const bool take = (lower == up) ? cand_less(od, oi, d, i) : cand_less(d, i, od, oi);
            if (take) { d = od; i = oi; }
        }
    }
}
// (ld, li) sorted ascending, (d, i) sorted ascending -> (ld, li) = the LW smallest of the 2 LW, sorted:
// min(list[l], new[LW - 1 - l]) is a bitonic sequence holding exactly those LW, log LW merge stages sort it.
template <int LW>
__device__ __forceinline__ void group_bitonic_merge(double& ld, int& li, double d, int i, int gl) {
    const double rd = __shfl_sync(0xffffffffu, d, LW - 1 - gl, LW);
    const int ri = __shfl_sync(0xffffffffu, i, LW - 1 - gl, LW);
    if (cand_less(rd, ri, ld, li)) { ld = rd; li = ri; }
#pragma unroll
    for (int j = LW >> 1; j > 0; j >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, ld, j, LW);
        const int oi = __shfl_xor_sync(0xffffffffu, li, j, LW);
        const bool lower = (gl & j) == 0;
        const bool take = lower ? cand_less(od, oi, ld, li) : cand_less(ld, li, od, oi);
        if (take) { ld = od; li = oi; }
    }
}

// LW lanes per query (LW = 16 for k <= 16: the fp64 pipe is not fed half-empty warps; LW = 32 above).  The group
// re-ranks its query's S x cand candidates LW at a time.  Candidate c of slice s: (cand_s, cand_i)[s][q][c].
template <int LW>
__global__ void __launch_bounds__(128) knn_tc_rerank_kernel(const float* __restrict__ bank, long row_offset,
                                                            const float* __restrict__ q, long Q, int k, int S, int cand,
                                                            const float* __restrict__ cand_s, const int* __restrict__ cand_i,
                                                            int* __restrict__ out_idx, double* __restrict__ out_d,
                                                            int* __restrict__ flagged, Stats* stats,
                                                            const BankStats* __restrict__ bstats, float eps_r,
                                                            const int* __restrict__ perm) {
    constexpr int GPW = 32 / LW;  // query groups per warp
    const long grp = (blockIdx.x * (long)blockDim.x + threadIdx.x) / LW;  // group index = query
    const int gl = threadIdx.x & (LW - 1), sub = (threadIdx.x & 31) / LW;
    const bool live = grp < Q;
    const long qi = live ? grp : Q - 1;  // a dead group shadows the last query (shuffles need all lanes), writes nothing
    // the query in fp64, 64 / LW elements per lane; |q|^2
    const float* qr = q + qi * 64;
    double qn = 0.0;
#pragma unroll
    for (int j = 0; j < 64 / LW; ++j) { const double v = (double)qr[gl + LW * j]; qn += v * v; }
#pragma unroll
    for (int o = LW >> 1; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o, LW);

    // Staging (per warp: GPW queries x LW candidate rows = 32 rows).  A lane needs its candidate's whole 256-byte row
    // in index order (the canonical sum is sequential), so reading rows straight from the bank makes every load
    // instruction touch 32 different lines; instead each group fetches its rows with contiguous 256-byte accesses
    // into shared memory (row pitch 272 B: the per-lane float4 reads are conflict-free per quarter warp), and the
    // query sits there once as fp64 (no per-candidate re-conversion).
    __shared__ __align__(16) float xs[4][32][68];
    __shared__ double qsd[4 * GPW][64];
    const int warp_l = threadIdx.x >> 5, grp_l = threadIdx.x / LW;
#pragma unroll
    for (int j = 0; j < 64 / LW; ++j) qsd[grp_l][gl + LW * j] = (double)qr[gl + LW * j];
    __syncwarp();
    const double* qd = qsd[grp_l];

    double ld = DBL_MAX;
    int li = INT_MAX;
    float tau = -FLT_MAX;       // max over slices of the slice's cand-th best approximate score
    float err = 0.f;
    const int total = S * cand;
    for (int base = 0; base < total; base += LW) {
        const int c = base + gl;
        int id = -1;
        float sc = -FLT_MAX;
        if (c < total) {
            const long o = ((long)(c / cand) * Q + qi) * cand + (c % cand);
            id = cand_i[o];
            if (bstats->has_perm && id >= 0) id = perm[id];  // scan position -> bank row
            sc = cand_s[o];
            if ((c % cand) == cand - 1 && id >= 0) tau = fmaxf(tau, sc);  // a full list: its tail bounds the rejected rows
        }
        __syncwarp();  // the previous round's rows have been consumed
#pragma unroll 4
        for (int cc = 0; cc < LW; ++cc) {
            const int idc = __shfl_sync(0xffffffffu, id, cc, LW);
            if (idc >= 0) {
                const float4* src = reinterpret_cast<const float4*>(bank + (long)idc * 64);
                if (LW == 16) {
                    *reinterpret_cast<float4*>(&xs[warp_l][sub * LW + cc][gl * 4]) = __ldg(src + gl);
                } else if (gl < 16) {
                    *reinterpret_cast<float4*>(&xs[warp_l][cc][gl * 4]) = __ldg(src + gl);
                }
            }
        }
        __syncwarp();
        double d = DBL_MAX;
        if (id >= 0) {  // canonical fp64 distance, same arithmetic as the exact sweep
            const float4* xr = reinterpret_cast<const float4*>(&xs[warp_l][sub * LW + gl][0]);
            double acc = 0.0;
            float xn = 0.f;  // |x|^2 only feeds the error diagnostic: fp32 is plenty
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const float4 x4 = xr[i];
                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double qv = qd[4 * i + j];
                    const double diff = __dsub_rn(qv, (double)xv[j]);
                    acc = __dadd_rn(acc, __dmul_rn(diff, diff));
                    xn = fmaf(xv[j], xv[j], xn);
                }
            }
            d = acc;
            // observed approximation error of the tensor-core score (diagnostic for EPS_REL)
            const float s_exact = 0.5f * (float)(qn + (double)xn - acc);
            err = fmaxf(err, fabsf(sc - s_exact));
        }
        // sort this round's candidates and merge them into the running list (invalid slots rank last)
        int ci = id >= 0 ? id + (int)row_offset : INT_MAX;
        group_bitonic_sort<LW>(d, ci, gl);
        if (base == 0) { ld = d; li = ci; }
        else group_bitonic_merge<LW>(ld, li, d, ci, gl);
    }
#pragma unroll
    for (int o = LW >> 1; o > 0; o >>= 1) {
        tau = fmaxf(tau, __shfl_xor_sync(0xffffffffu, tau, o, LW));
        err = fmaxf(err, __shfl_xor_sync(0xffffffffu, err, o, LW));
    }
    if (live && gl < k) { out_idx[qi * k + gl] = li; out_d[qi * k + gl] = ld; }
    // proof: rejected rows have s~ <= tau  =>  d >= |q|^2 + min|x|^2 - 2 (tau + eps)
    const double dk = __shfl_sync(0xffffffffu, ld, k - 1, LW);
    bool proven = true;
    if (tau > -FLT_MAX) {
        const double nmin = (double)__uint_as_float(bstats->nmin_bits);
        const double nmax = (double)__uint_as_float(bstats->nmax_bits);
        const double eps = (double)eps_r * sqrt(qn * nmax) + 2e-6;
        const double d_lb = qn + nmin * (1.0 - 1e-6) - 2.0 * ((double)tau + eps);
        proven = dk < d_lb - 1e-9;
    }
    if (live && gl == 0) {
        if (!proven) flagged[atomicAdd(&stats->n_flagged, 1)] = (int)qi;
        atomicMax(&stats->max_err_bits, __float_as_uint(err));
    }
}


size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <int KBLK, int CAND>
int set_candidates_attr() {
    RF_SMEM_OPT_IN((knn_tc_candidates_kernel<KBLK, CAND, 0>), (Cfg<KBLK, CAND>::SMEM_BYTES));
    RF_SMEM_OPT_IN((knn_tc_candidates_kernel<KBLK, CAND, 1>), (Cfg<KBLK, CAND>::SMEM_BYTES));
    return 0;
}

// ---- the prepared bank image: [BankStats 256 B][perm: bpad int32][tile images, 1024-aligned]
struct ImgLayout {
    size_t perm, img, total, key_in, key_out, idx_in, mean, sort_tmp, sort_tmp_bytes, scratch_total;
    int n_btiles;
};
ImgLayout img_layout(long n_rows, int kblk) {
    ImgLayout L;
    L.n_btiles = (int)((n_rows + TILE - 1) / TILE);
    const size_t bpad = (size_t)L.n_btiles * TILE;
    size_t off = 256;
    L.perm = off; off += align_up(bpad * sizeof(int), 1024);
    L.img = off; off += (size_t)L.n_btiles * kblk * KB_BYTES;
    L.total = off;
    // scratch of the preparation (scan order): mean direction, projection keys, radix-sort buffers
    size_t so = 0;
    L.mean = so; so += 256;
    L.key_in = so; so += align_up(bpad * sizeof(float), 256);
    L.key_out = so; so += align_up(bpad * sizeof(float), 256);
    L.idx_in = so; so += align_up(bpad * sizeof(int), 256);
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairsDescending((void*)nullptr, tmp, (const float*)nullptr, (float*)nullptr, (const int*)nullptr,
                                              (int*)nullptr, (int)bpad);
    L.sort_tmp_bytes = tmp;
    L.sort_tmp = so; so += align_up(tmp, 256);
    L.scratch_total = so;
    return L;
}

// image: 1024-aligned, img_layout().total bytes; scratch: img_layout().scratch_total bytes (only read during the call)
template <int KBLK>
int prepare_bank(const float* bank, long n_rows, const float* q_sample, long n_sample, char* image, char* scratch, cudaStream_t s) {
    const ImgLayout L = img_layout(n_rows, KBLK);
    BankStats* bstats = (BankStats*)image;
    const long bpad = (long)L.n_btiles * TILE;
    // scan order (see knn_scan_key_kernel) only when the caller supplies a sample of the queries the image will serve:
    // an order derived from OTHER queries can be worse than none (rows that look unpromising for the sample come last
    // and then beat every list).  Pointless for tiny banks.
    const bool order = q_sample != nullptr && n_rows >= 16 * TILE;
    const BankStats init = {0x7f7fffffu /*FLT_MAX*/, 0u, order ? 1 : 0};
    RF_CUDA_OK(cudaMemcpyAsync(bstats, &init, sizeof(BankStats), cudaMemcpyHostToDevice, s));
    const int* perm = nullptr;
    if (order) {
        float* mean = (float*)(scratch + L.mean);
        float* key_in = (float*)(scratch + L.key_in);
        float* key_out = (float*)(scratch + L.key_out);
        int* idx_in = (int*)(scratch + L.idx_in);
        int* perm_w = (int*)(image + L.perm);
        RF_CUDA_OK(cudaMemsetAsync(mean, 0, 64 * sizeof(float), s));
        long ns = n_sample;  // direction of a typical query: the mean of the sample
        if (ns > 65536) ns = 65536;
        knn_mean_query_kernel<<<(unsigned)(ns / 4 < 592 ? (ns + 3) / 4 : 592), 256, 0, s>>>(q_sample, ns, mean);
        RF_LAUNCH_OK("knn_mean_query_kernel");
        knn_scan_key_kernel<<<(unsigned)rf_cdivl(bpad, 256), 256, 0, s>>>(bank, n_rows, bpad, mean, key_in, idx_in);
        RF_LAUNCH_OK("knn_scan_key_kernel");
        size_t tmp = L.sort_tmp_bytes;
        RF_CUDA_OK(cub::DeviceRadixSort::SortPairsDescending((void*)(scratch + L.sort_tmp), tmp, (const float*)key_in, key_out,
                                                             (const int*)idx_in, perm_w, (int)bpad, 0, 32, s));
        perm = perm_w;
    }
    knn_tc_prep_kernel<KBLK><<<(unsigned)rf_cdivl(bpad * 8, 256), 256, 0, s>>>(bank, n_rows, bpad, 1, (uint8_t*)(image + L.img), bstats, perm);
    RF_LAUNCH_OK("knn_tc_prep_kernel(bank)");
    return 0;
}

// ---- the per-call workspace
struct TcLayout {
    size_t stats, q_img, cand_s, cand_i, flagged, recheck_ws, recheck_ws_bytes, total;
    int n_btiles, n_qtiles, n_qpairs, nsplit, cand, lw;
};

TcLayout tc_layout(long Q, long n_rows, int k, int kblk) {
    TcLayout L;
    const size_t tile_bytes = (size_t)kblk * KB_BYTES;
    L.n_btiles = (int)((n_rows + TILE - 1) / TILE);
    L.n_qpairs = (int)((Q + MQ * TILE - 1) / (MQ * TILE));
    L.n_qtiles = L.n_qpairs * MQ;
    L.cand = k <= 8 ? 16 : 32;
    L.lw = k <= 16 ? 16 : 32;
    int ns = (148 + L.n_qpairs - 1) / L.n_qpairs;  // fill the chip when there are few query tiles
    if (ns > MAX_SPLIT) ns = MAX_SPLIT;
    if (k > 16 && ns < 2) ns = 2;                   // 2 x 32 candidates for k up to 32
    if (ns > L.n_btiles) ns = L.n_btiles;
    if (ns < 1) ns = 1;
    L.nsplit = ns;
    size_t off = 0;
    L.stats = off; off += 256;
    off = align_up(off, 1024);
    L.q_img = off; off += (size_t)L.n_qtiles * tile_bytes;
    L.cand_s = off; off += align_up((size_t)L.nsplit * Q * L.cand * sizeof(float), 256);
    L.cand_i = off; off += align_up((size_t)L.nsplit * Q * L.cand * sizeof(int), 256);
    L.flagged = off; off += align_up((size_t)Q * sizeof(int), 256);
    L.recheck_ws_bytes = rf_knn_recheck_workspace_bytes(k);
    L.recheck_ws = off; off += L.recheck_ws_bytes;
    L.total = off;
    return L;
}

template <int KBLK, int CAND>
int launch_candidates(const TcLayout& L, const uint8_t* q_img, const uint8_t* bank_img, long Q, long n_rows, float* cand_s,
                      int* cand_i, cudaStream_t s) {
    if (int rc = set_candidates_attr<KBLK, CAND>()) return rc;
    dim3 grid(L.n_qpairs, L.nsplit);
    // tuning aids: RF_KNN_COOP=0 selects the register-list kernel, RF_KNN_COOP_MAX the largest number of hot lanes
    // the cooperative insertion serves before the per-lane path takes over
    static const int coop = [] { const char* e = getenv("RF_KNN_COOP"); return e ? atoi(e) : 1; }();
    static const int coop_max = [] { const char* e = getenv("RF_KNN_COOP_MAX"); return e ? atoi(e) : 2; }();
    if (coop)
        knn_tc_candidates_kernel<KBLK, CAND, 1><<<grid, NTHREADS, Cfg<KBLK, CAND>::SMEM_BYTES, s>>>(q_img, bank_img, Q, n_rows, L.n_qtiles,
                                                                                             L.n_btiles, L.nsplit, cand_s, cand_i, coop_max);
    else
        knn_tc_candidates_kernel<KBLK, CAND, 0><<<grid, NTHREADS, Cfg<KBLK, CAND>::SMEM_BYTES, s>>>(q_img, bank_img, Q, n_rows, L.n_qtiles,
                                                                                             L.n_btiles, L.nsplit, cand_s, cand_i, coop_max);
    RF_LAUNCH_OK("knn_tc_candidates_kernel");
    return 0;
}

template <int KBLK>
int tc_run(const TcLayout& L, const float* bank, long n_rows, long row_offset, const char* image, const float* q, long Q, int k,
           int* out_idx, double* out_d, char* ws, cudaStream_t s) {
    const ImgLayout IL = img_layout(n_rows, KBLK);
    Stats* stats = (Stats*)(ws + L.stats);
    const BankStats* bstats = (const BankStats*)image;
    const uint8_t* bank_img = (const uint8_t*)(image + IL.img);
    const int* perm = (const int*)(image + IL.perm);  // used when the image header says the rows are in scan order
    uint8_t* q_img = (uint8_t*)(ws + L.q_img);
    float* cand_s = (float*)(ws + L.cand_s);
    int* cand_i = (int*)(ws + L.cand_i);
    int* flagged = (int*)(ws + L.flagged);
    RF_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(Stats), s));
    const long qpad = (long)L.n_qtiles * TILE;
    knn_tc_prep_kernel<KBLK><<<(unsigned)rf_cdivl(qpad * 8, 256), 256, 0, s>>>(q, Q, qpad, 0, q_img, nullptr, nullptr);
    RF_LAUNCH_OK("knn_tc_prep_kernel(queries)");
    int rc = L.cand == 16 ? launch_candidates<KBLK, 16>(L, q_img, bank_img, Q, n_rows, cand_s, cand_i, s)
                          : launch_candidates<KBLK, 32>(L, q_img, bank_img, Q, n_rows, cand_s, cand_i, s);
    if (rc) return rc;
    if (L.lw == 16)
        knn_tc_rerank_kernel<16><<<(unsigned)rf_cdivl(Q * 16, 128), 128, 0, s>>>(bank, row_offset, q, Q, k, L.nsplit, L.cand, cand_s, cand_i,
                                                                                out_idx, out_d, flagged, stats, bstats, eps_rel(KBLK), perm);
    else
        knn_tc_rerank_kernel<32><<<(unsigned)rf_cdivl(Q * 32, 128), 128, 0, s>>>(bank, row_offset, q, Q, k, L.nsplit, L.cand, cand_s, cand_i,
                                                                                out_idx, out_d, flagged, stats, bstats, eps_rel(KBLK), perm);
    RF_LAUNCH_OK("knn_tc_rerank_kernel");
    // unproven queries (normally none): exact fp64 sweep sized by the device counter, results scattered through `flagged`
    return rf_knn_recheck_launch(bank, n_rows, row_offset, q, Q, k, flagged, &stats->n_flagged, out_idx, out_d, ws + L.recheck_ws,
                                 L.recheck_ws_bytes, s);
}

}  // namespace

int rf_knn_tc_init() {
    if (int rc = set_candidates_attr<1, 16>()) return rc;
    if (int rc = set_candidates_attr<1, 32>()) return rc;
    if (int rc = set_candidates_attr<3, 16>()) return rc;
    return set_candidates_attr<3, 32>();
}

// kblk = 1: fp16 single pass (method 2); kblk = 3: bf16 hi/lo split (method 3)
size_t rf_knn_tc_image_bytes(long n_rows, int kblk) { return img_layout(n_rows, kblk).total + 1024; }
size_t rf_knn_tc_prepare_scratch_bytes(long n_rows, int kblk) { return img_layout(n_rows, kblk).scratch_total + 256; }

int rf_knn_tc_prepare(const float* bank, long n_rows, int kblk, const float* q_sample, long n_sample, void* image,
                      size_t image_bytes, void* scratch, size_t scratch_bytes, cudaStream_t s) {
    const ImgLayout IL = img_layout(n_rows, kblk);
    char* img = (char*)(((uintptr_t)image + 1023) & ~(uintptr_t)1023);
    RF_CHECK_ARG(image && image_bytes >= IL.total + (size_t)(img - (char*)image), "rf_knn_bank_prepare: image buffer too small (%zu < %zu)",
                 image_bytes, IL.total + 1024);
    RF_CHECK_ARG(scratch && scratch_bytes >= IL.scratch_total, "rf_knn_bank_prepare: scratch too small (%zu < %zu)", scratch_bytes,
                 IL.scratch_total);
    return kblk == 1 ? prepare_bank<1>(bank, n_rows, q_sample, n_sample, img, (char*)scratch, s)
                     : prepare_bank<3>(bank, n_rows, q_sample, n_sample, img, (char*)scratch, s);
}

size_t rf_knn_tc_workspace_bytes(long Q, long n_rows, int k, int kblk, int with_image) {
    size_t b = tc_layout(Q, n_rows, k, kblk).total + 1024;
    if (with_image) b += rf_knn_tc_image_bytes(n_rows, kblk) + rf_knn_tc_prepare_scratch_bytes(n_rows, kblk) + 1024;
    return b;
}

// image == nullptr: the bank image is built inside the workspace first (one-off lookups); otherwise `image` is a buffer
// filled by rf_knn_tc_prepare for exactly this bank
int rf_knn_tc_launch(const float* bank, long n_rows, long row_offset, const void* image, const float* q, long Q, int k, int kblk,
                     int* out_idx, double* out_d, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    RF_CHECK_ARG(k <= 32, "rf_knn_l2_topk(tensor-core methods): k=%d exceeds 32", k);
    RF_CHECK_ARG(n_rows >= k, "rf_knn_l2_topk(tensor-core methods): fewer bank rows than k");
    const TcLayout L = tc_layout(Q, n_rows, k, kblk);
    char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);  // swizzled tile images need 1024-byte alignment
    size_t need = L.total + (size_t)(ws - (char*)workspace);
    const char* img = nullptr;
    if (image) {
        img = (const char*)(((uintptr_t)image + 1023) & ~(uintptr_t)1023);
    } else {
        char* own = ws + align_up(L.total, 1024);
        const size_t ib = align_up(img_layout(n_rows, kblk).total, 1024);
        need += ib + img_layout(n_rows, kblk).scratch_total + 1024;
        RF_CHECK_ARG(workspace && workspace_bytes >= need, "rf_knn_l2_topk(tensor-core methods): workspace too small (%zu < %zu)",
                     workspace_bytes, need);
        int rc = kblk == 1 ? prepare_bank<1>(bank, n_rows, q, Q, own, own + ib, s) : prepare_bank<3>(bank, n_rows, q, Q, own, own + ib, s);
        if (rc) return rc;
        img = own;
    }
    RF_CHECK_ARG(workspace && workspace_bytes >= need, "rf_knn_l2_topk(tensor-core methods): workspace too small (%zu < %zu)",
                 workspace_bytes, need);
    return kblk == 1 ? tc_run<1>(L, bank, n_rows, row_offset, img, q, Q, k, out_idx, out_d, ws, s)
                     : tc_run<3>(L, bank, n_rows, row_offset, img, q, Q, k, out_idx, out_d, ws, s);
}

// Diagnostics of the last tensor-core call that used `workspace` (synchronises the stream).
extern "C" int rf_knn_tc_stats(const void* workspace, int* n_unproven, float* max_score_err, void* stream) {
    RF_CHECK_ARG(workspace, "rf_knn_tc_stats: null workspace");
    Stats h;
    const void* ws = (const void*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    RF_CUDA_OK(cudaMemcpyAsync(&h, ws, sizeof(Stats), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RF_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    if (n_unproven) *n_unproven = h.n_flagged;
    if (max_score_err) memcpy(max_score_err, &h.max_err_bits, sizeof(float));
    return 0;
}
