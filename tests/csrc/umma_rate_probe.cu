// Probe (test infrastructure): issue-to-completion cost of small tcgen05.mma.kind::tf32 instructions (M = 128, K = 8)
// as a function of N, of the number of independent accumulators R the stream rotates over, and of where A lives
// (tensor memory vs shared memory).  One CTA, one issuing thread, NITER MMAs, one commit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate_probe tests/csrc/umma_rate_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_of(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128) rate(int N, int R, int a_in_tmem, int niter, long long *out) {
    extern __shared__ __align__(128) float sm[];     // A tile [128 x 8] (4 KB) + B tile [256 x 8] (8 KB), zeros
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 3072; i += 128) sm[i] = 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t td = tbase;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = desc_of(s32(sm)), db = desc_of(s32(sm + 1024));
        const uint32_t ta = td + 448u;                   // A (8 columns) at the top of the allocation
        const long long t0 = clock64();
#pragma unroll 4
        for (int i = 0; i < niter; ++i) {
            const uint32_t d = td + (uint32_t)((i & (R - 1)) * N);     // R is a power of two
            if (a_in_tmem)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(1u) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(s32(&bar)), "r"(0u) : "memory");
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(td) : "memory");
}

int main() {
    long long *out; cudaMalloc(&out, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    const int niter = 512;
    printf("%-6s %-4s %-3s %14s %14s\n", "A", "N", "R", "issue cyc/MMA", "total cyc/MMA");
    for (int a_in_tmem = 1; a_in_tmem >= 0; --a_in_tmem)
        for (int N : {16, 32, 64, 128, 256})
            for (int R : {1, 2, 4, 8}) {
                if (R * N > 448) continue;
                rate<<<1, 128, 12288>>>(N, R, a_in_tmem, niter, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("%-6s %-4d %-3d %14.1f %14.1f\n", a_in_tmem ? "tmem" : "smem", N, R, (double)h[0] / niter, (double)h[1] / niter);
            }
    return 0;
}
