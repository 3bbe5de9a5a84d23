// Probe of the tcgen05 shared-memory operand layout (MN-major, no swizzle, kind::tf32, M=128, K=8).
// Test infrastructure: places one-hot values at every shared-memory offset of an operand image and reports
// which accumulator row / column / k index the tensor core reads it as.  Build: see scripts/gpu_umma_probe.sh
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_of(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// mode 0: A one-hot at offset blockIdx.x, B all ones   -> out[b] = {first nonzero row, #nonzero rows, first nonzero col, #cols}
// mode 1: A all ones, B one-hot at offset blockIdx.x   -> same
// mode 2: A one-hot at oa, B one-hot at blockIdx.x     -> same (nonzero iff same k)
__global__ void __launch_bounds__(128) probe(int mode, int oa, int n16, int a_floats, int b_floats, uint32_t lbo_a,
                                             uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b, int a_major, int b_major,
                                             int4 *out) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    __shared__ int first_row, nrows, first_col, ncols;
    float *As = sm, *Bs = sm + a_floats;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < a_floats; i += 128) As[i] = mode == 1 ? 1.f : (i == (mode == 0 ? (int)blockIdx.x : oa) ? 1.f : 0.f);
    for (int i = tid; i < b_floats; i += 128) Bs[i] = mode == 0 ? 1.f : (i == (int)blockIdx.x ? 1.f : 0.f);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        first_row = 1 << 20; nrows = 0; first_col = 1 << 20; ncols = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(s32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t td = tbase;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) |
                               ((uint32_t)(n16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = desc_of(s32(As), lbo_a, sbo_a), db = desc_of(s32(Bs), lbo_b, sbo_b);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(td), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0; int spins = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(s32(&bar)), "r"(0u) : "memory");
            if (++spins > (1 << 22)) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int my_cols = 0, my_first = 1 << 20;
    for (int c = 0; c < n16; c += 8) {
        uint32_t v[8];
        const uint32_t taddr = td + ((uint32_t)(32 * warp) << 16) + (uint32_t)c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 8; ++e) if (__uint_as_float(v[e]) != 0.f) { ++my_cols; if (c + e < my_first) my_first = c + e; }
    }
    if (my_cols) { atomicMin(&first_row, tid); atomicAdd(&nrows, 1); atomicMin(&first_col, my_first); atomicMax(&ncols, my_cols); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = make_int4(nrows ? first_row : -1, nrows, ncols ? first_col : -1, ncols);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(td) : "memory");
}

static void run(const char *name, int mode, int oa, int n16, int a_floats, int b_floats, uint32_t lbo_a, uint32_t sbo_a,
                uint32_t lbo_b, uint32_t sbo_b, int a_major, int b_major, int nprobe, int stride) {
    int4 *d; cudaMalloc(&d, sizeof(int4) * nprobe);
    const int smem = (a_floats + b_floats) * 4;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<nprobe, 128, smem>>>(mode, oa, n16, a_floats, b_floats, lbo_a, sbo_a, lbo_b, sbo_b, a_major, b_major, d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("## %s  (mode %d, lboA %u sboA %u lboB %u sboB %u, major %d/%d): %s\n", name, mode, lbo_a, sbo_a, lbo_b, sbo_b,
           a_major, b_major, cudaGetErrorString(e));
    if (e != cudaSuccess) exit(1);
    std::vector<int4> h(nprobe); cudaMemcpy(h.data(), d, sizeof(int4) * nprobe, cudaMemcpyDeviceToHost);
    for (int i = 0; i < nprobe; i += stride) printf("  off %4d -> row %3d (x%d) col %3d (x%d)\n", i, h[i].x, h[i].y, h[i].z, h[i].w);
    cudaFree(d);
}

int main() {
    const int n16 = 32;
    const int a_floats = 128 * 8 * 2, b_floats = n16 * 8 * 2;   // two k-blocks worth so stray reads are visible
    // MN-major, as the kernel uses it: unit (16 B) = 4 MN elements, k rows 16 B apart, MN units 128 B apart
    run("A one-hot, MN-major", 0, 0, n16, a_floats, b_floats, 4096, 128, n16 * 32, 128, 1, 1, 80, 1);
    run("A one-hot, MN-major (sparse scan)", 0, 0, n16, a_floats, b_floats, 4096, 128, n16 * 32, 128, 1, 1, 2048, 97);
    run("B one-hot, MN-major", 1, 0, n16, a_floats, b_floats, 4096, 128, n16 * 32, 128, 1, 1, 80, 1);
    run("k match: A@0 vs B scan", 2, 0, n16, a_floats, b_floats, 4096, 128, n16 * 32, 128, 1, 1, 64, 1);
    run("k match: A@4 vs B scan", 2, 4, n16, a_floats, b_floats, 4096, 128, n16 * 32, 128, 1, 1, 64, 1);
    run("k match: A@1 vs B scan", 2, 1, n16, a_floats, b_floats, 4096, 128, n16 * 32, 128, 1, 1, 64, 1);
    // swapped roles of LBO / SBO
    run("A one-hot, MN-major, LBO<->SBO", 0, 0, n16, a_floats, b_floats, 128, 4096, 128, n16 * 32, 1, 1, 80, 1);
    // K-major: 8 MN rows x 16 B (4 k) core matrix; MN groups of 8 SBO apart; the two k halves LBO apart
    run("A one-hot, K-major", 0, 0, n16, a_floats, b_floats, 128, 256, 128, 256, 0, 0, 80, 1);
    run("B one-hot, K-major", 1, 0, n16, a_floats, b_floats, 128, 256, 128, 256, 0, 0, 80, 1);
    run("k match K-major: A@0 vs B scan", 2, 0, n16, a_floats, b_floats, 128, 256, 128, 256, 0, 0, 80, 1);
    run("k match K-major: A@1 vs B scan", 2, 1, n16, a_floats, b_floats, 128, 256, 128, 256, 0, 0, 80, 1);
    run("k match K-major: A@32 vs B scan", 2, 32, n16, a_floats, b_floats, 128, 256, 128, 256, 0, 0, 80, 1);
    return 0;
}
