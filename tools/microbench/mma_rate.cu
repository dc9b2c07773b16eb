// Micro-benchmark (development aid): cycles per tcgen05.mma (M128, K16, kind::f16, cta_group::1) as a function of
// the shared-memory operand layout (no swizzle / 32B / 64B / 128B), start-address alignment, SBO / LBO and N.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mma_rate tools/microbench/mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

struct Cfg {
    const char* name;
    int layout_a, layout_b;          // descriptor layout type: 0 none, 2 128B, 4 64B, 6 32B
    uint32_t a_off, a_lbo, a_sbo;    // bytes
    uint32_t a_step, a_wrap;         // A start advances by a_step per MMA, modulo a_wrap
    uint32_t b_lbo, b_sbo;
    int N, n_mma, pollers;           // pollers: number of extra warps spinning on an mbarrier meanwhile
    int use_base_offset;
    int n_acc;                       // independent accumulators cycled through (n_acc * N <= 512 columns)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout, int use_bo) {
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo & 0x3FFFFu) >> 4) << 16) |
                 ((uint64_t)((sbo & 0x3FFFFu) >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
    if (use_bo) d |= (uint64_t)((addr >> 7) & 7u) << 49;
    return d;
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

__global__ void __launch_bounds__(256, 1) bench(Cfg c, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* al = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + 160 * 1024, bar = base + 200 * 1024, slot = bar + 16, bar2 = bar + 8;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(al)[i] = make_uint4(0x3c003c00, 0x3c003c00, 0x3c003c00, 0x3c003c00);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
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
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(al + (slot - base));
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
        const uint64_t db = mk_desc(sB, c.b_lbo, c.b_sbo, c.layout_b, 0);
        long long t0 = clock64();
        uint32_t off = 0;
        for (int i = 0; i < c.n_mma; ++i) {
            const uint64_t da = mk_desc(sA + c.a_off + off, c.a_lbo, c.a_sbo, c.layout_a, c.use_base_offset);
            off += c.a_step;
            if (off >= c.a_wrap) off = 0;
            const uint32_t d = tmem + (uint32_t)((i % c.n_acc) * c.N);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(da), "l"(db), "r"(idesc), "r"(i >= c.n_acc ? 1u : 0u) : "memory");
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        while (!try_wait(bar, 0)) {}
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar2) : "memory");
    } else if (warp >= 1 && warp <= c.pollers) {
        while (!try_wait(bar2, 0)) {}
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

int main() {
    const int NM = 2048;
    static Cfg cfgs[64];
    static char names[64][64];
    int nc = 0;
    const int Ns[] = {16, 32, 64, 128, 256};
    const int accs[] = {1, 2, 3, 4, 5, 6, 8, 16};
    for (int N : Ns)
        for (int na : accs) {
            if (na * N > 512) continue;
            snprintf(names[nc], 64, "NS lines N%-3d acc%-2d", N, na);
            cfgs[nc] = Cfg{names[nc], 0, 0, 16, 16384, 160, 2560, 65536, (uint32_t)N * 16, 128, N, NM, 0, 0, na};
            ++nc;
        }
    snprintf(names[nc], 64, "SW128 N64 acc8"); cfgs[nc] = Cfg{names[nc], 2, 2, 0, 16, 1024, 32, 128, 16, 1024, 64, NM, 0, 0, 8}; ++nc;
    snprintf(names[nc], 64, "SW128 N16 acc8"); cfgs[nc] = Cfg{names[nc], 2, 2, 0, 16, 1024, 32, 128, 16, 1024, 16, NM, 0, 0, 8}; ++nc;
    snprintf(names[nc], 64, "SW128 N32 acc8"); cfgs[nc] = Cfg{names[nc], 2, 2, 0, 16, 1024, 32, 128, 16, 1024, 32, NM, 0, 0, 8}; ++nc;
    snprintf(names[nc], 64, "SW32 N16 acc8"); cfgs[nc] = Cfg{names[nc], 6, 6, 32, 16, 256, 32, 16384, 16, 256, 16, NM, 0, 0, 8}; ++nc;
    snprintf(names[nc], 64, "NS N64 acc8 +7poll"); cfgs[nc] = Cfg{names[nc], 0, 0, 16, 16384, 160, 2560, 65536, 1024, 128, 64, NM, 7, 0, 8}; ++nc;
    long long* out;
    cudaMalloc(&out, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    for (int ci = 0; ci < nc; ++ci) {
        Cfg& c = cfgs[ci];
        for (int grid : {148}) {
            long long h[2] = {0, 0};
            bench<<<grid, 256, 210 * 1024>>>(c, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%-26s grid %3d: ERROR %s\n", c.name, grid, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("%-26s grid %3d: issue %.1f cyc/mma, complete %.1f cyc/mma\n", c.name, grid, (double)h[0] / c.n_mma, (double)h[1] / c.n_mma);
        }
    }
    return 0;
}
