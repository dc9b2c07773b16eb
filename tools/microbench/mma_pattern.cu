// Micro-benchmark (development aid): tensor-pipe time per tcgen05.mma (M128 K16 kind::f16) for the operand /
// accumulator SEQUENCES a shifted-window convolution can issue.  Whole-warp uniform issue loop + one elected
// lane, fully unrolled 15-MMA period, so the issue cost is a few instructions per MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mma_pattern tools/microbench/mma_pattern.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}

// PAT: 0 all same | 1 per tile (Ah,Bh)(Ah,Bl)(Al,Bh) same D, 5 tiles | 2 per product: 5 tiles (different A and D), B per product
//      3 A varies (15), B, D same | 4 A, B same, D rotates over 5 | 5 A varies, D rotates over 5, B same
//      6 like 2 but 15 distinct accumulators | 7 like 1 but the three products go to 3 different accumulators
template <int PAT>
__device__ __forceinline__ void period(uint32_t tm, uint32_t a0, uint32_t ah, uint32_t b0, uint32_t bh, uint32_t idesc, uint32_t N,
                                       uint32_t tile_step, uint32_t lo_off, uint32_t blk, uint32_t leader, uint32_t shift) {
#pragma unroll
    for (int j = 0; j < 15; ++j) {
        uint32_t a = a0 + shift, b = b0, d = tm;
        if (PAT == 1 || PAT == 7) {
            const int t = j / 3, p = j % 3;
            a += t * tile_step + (p == 2 ? lo_off : 0); b += (p == 1 ? blk : 0);
            d += (PAT == 7 ? (uint32_t)(p * 5 + t) : (uint32_t)t) * N;
        } else if (PAT == 2 || PAT == 6) {
            const int p = j / 5, t = j % 5;
            a += t * tile_step + (p == 2 ? lo_off : 0); b += (p == 1 ? blk : 0);
            d += (PAT == 6 ? (uint32_t)j : (uint32_t)t) * N;
        } else if (PAT == 3) {
            a += (j % 5) * tile_step + (j / 5) * 1;
        } else if (PAT == 4) {
            d += (j % 5) * N;
        } else if (PAT == 5) {
            a += (j % 5) * tile_step + (j / 5) * 1; d += (j % 5) * N;
        }
        if (leader) mma(d, a, ah, b, bh, idesc, 1u);
    }
}

template <int PAT>
__global__ void __launch_bounds__(256, 1) bench(int N, int layout, int reps, long long* out, int flags) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + 160 * 1024, bar = base + 200 * 1024, slot = bar + 16;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) {
        uint4 v = make_uint4(0x3c003c00, 0x3c003c00, 0x3c003c00, 0x3c003c00);
        if (flags & 2) {  // pseudo-random fp16 values in (-2, 2)
            uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
            uint32_t w[4];
            for (int k = 0; k < 4; ++k) { h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15; w[k] = (h & 0x83ff83ffu) | 0x38003800u; }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (flags & 8) v = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(al)[i] = v;
    }
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(al + (slot - base)), 0);
    if (warp == 1) {
        const uint32_t leader = elect_one();
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        uint32_t a0, ah, b0, bh, tile_step, lo_off, blk;
        if (layout == 0) {  // no swizzle, "lines" geometry of an 8^3 patch: SBO 160 B, LBO one plane (1032 slots)
            a0 = ((sA & 0x3FFFFu) >> 4) | (1032u << 16); ah = 10u | (1u << 14);
            b0 = ((sB & 0x3FFFFu) >> 4) | ((uint32_t)N << 16); bh = 8u | (1u << 14);
            tile_step = 160; lo_off = 2064; blk = (uint32_t)N * 2;
        } else {            // 128B swizzle, K-major, 1024 B per 8-row group
            a0 = ((sA & 0x3FFFFu) >> 4) | (1u << 16); ah = 64u | (1u << 14) | (2u << 29);
            b0 = ((sB & 0x3FFFFu) >> 4) | (1u << 16); bh = 64u | (1u << 14) | (2u << 29);
            tile_step = 1024; lo_off = 5120; blk = (uint32_t)N * 8;
        }
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t bsel = (flags & 4) ? (uint32_t)(r & 3) * 768u : 0u;  // rotate over 4 weight slots of 12 KiB
            period<PAT>(tmem, a0, ah, b0 + bsel, bh, idesc, (uint32_t)N, tile_step, lo_off, blk, leader, layout == 0 ? (uint32_t)(r & 7) * 10u : 0u);
            if ((flags & 1) && (r % 3) == 2 && leader)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8) : "memory");
        }
        long long t1 = clock64();
        if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        __syncwarp();
        while (!try_wait(bar, 0)) {}
        long long t2 = clock64();
        if (blockIdx.x == 0 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int PAT>
void run(const char* name, int N, int layout, long long* out, int flags = 0, int threads = 128) {
    const int reps = 200;
    cudaFuncSetAttribute(bench<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    long long h[2] = {0, 0};
    bench<PAT><<<148, threads, 210 * 1024>>>(N, layout, reps, out, flags);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-52s N%-3d %s: ERROR %s\n", name, N, layout ? "SW128" : "NS   ", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-52s N%-3d %s flags %d thr %d: issue %.1f  complete %.1f cyc/mma\n", name, N, layout ? "SW128" : "NS   ", flags, threads, (double)h[0] / (reps * 15), (double)h[1] / (reps * 15));
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    for (int layout = 0; layout < 2; ++layout) {
        run<2>("2 base", 64, layout, out, 0, 128);
        run<2>("2 +commit/45", 64, layout, out, 1, 128);
        run<2>("2 +random data", 64, layout, out, 2, 128);
        run<2>("2 +zero data", 64, layout, out, 8, 128);
        run<2>("2 +B slot rotation", 64, layout, out, 4, 128);
        run<2>("2 +256 threads", 64, layout, out, 0, 256);
        run<2>("2 all", 64, layout, out, 7, 256);
        run<1>("1 all", 64, layout, out, 7, 256);
        run<0>("0 random", 64, layout, out, 2, 128);
        run<0>("0 random", 16, layout, out, 2, 128);
        run<0>("0 random", 128, layout, out, 2, 128);
        run<0>("0 random", 256, layout, out, 2, 128);
    }
    return 0;
}
