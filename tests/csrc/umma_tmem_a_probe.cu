// Probe (test infrastructure): tcgen05.mma.kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st),
// B in shared memory (K-major no-swizzle canonical layout), M = 128, N = 16, K = 8 per instruction.
// Confirms: A[m][k] lives at TMEM lane m, column a_base + k (one 32-bit column per tf32 element); a second
// k-block at a_base + 8; accumulation into D[m][n] at lane m, column d_base + n.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_tmem_a_probe tests/csrc/umma_tmem_a_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_of(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(128) probe(float *out, int nkb) {
    __shared__ __align__(128) float Bs[2 * 16 * 8];     // two k-blocks of B[16][8]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // B[n][k] of k-block q: 1 if n == k + 8q (q = 1: shifted so that D[m][8 + k] picks up A[m][8 + k])
    for (int i = tid; i < 2 * 128; i += 128) {
        const int q = i / 128, r = i % 128;
        const int n = (r / 64) * 8 + (r % 32) / 4, k = ((r % 64) / 32) * 4 + (r % 4);
        Bs[i] = (n == k + 8 * q) ? 1.f : 0.f;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t td = tbase;               // D at columns 0..15, A at columns 32..47
    const int m = 32 * warp + lane;
    {
        uint32_t v[16];
        for (int k = 0; k < 16; ++k) v[k] = __float_as_uint((float)(m * 16 + k) * 0.5f);
        const uint32_t ta = td + ((uint32_t)(32 * warp) << 16) + 32u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta), "r"(v[0]), "r"(v[1]),
                     "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta + 8u), "r"(v[8]), "r"(v[9]),
                     "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int q = 0; q < nkb; ++q) {
            const uint64_t db = desc_of(s32(Bs + q * 128));
            const uint32_t ta = td + 32u + 8u * q;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(td), "r"(ta), "l"(db), "r"(idesc), "r"((uint32_t)q) : "memory");
        }
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
    uint32_t r[16];
    const uint32_t tl = td + ((uint32_t)(32 * warp) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tl));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tl + 8u));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int k = 0; k < 16; ++k) out[m * 16 + k] = __uint_as_float(r[k]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(td) : "memory");
}

int main() {
    float *out; cudaMalloc(&out, 128 * 16 * 4);
    for (int nkb = 1; nkb <= 2; ++nkb) {
        cudaMemset(out, 0, 128 * 16 * 4);
        probe<<<1, 128>>>(out, nkb);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("nkb=%d: %s\n", nkb, cudaGetErrorString(e)); return 1; }
        float h[128 * 16]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int k = 0; k < 16; ++k) {
                const float want = (k < 8 * nkb) ? (float)(m * 16 + k) * 0.5f : 0.f;
                if (h[m * 16 + k] != want) { if (bad < 8) printf("  nkb=%d m=%d n=%d got %g want %g\n", nkb, m, k, h[m * 16 + k], want); ++bad; }
            }
        printf("A-in-TMEM probe, %d k-block(s): %d mismatches of 2048; D[5][0..15] =", nkb, bad);
        for (int k = 0; k < 16; ++k) printf(" %g", h[5 * 16 + k]);
        printf("\n");
    }
    return 0;
}
