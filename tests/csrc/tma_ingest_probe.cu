// Probe (test infrastructure): per-SM ingest rate of TMA bulk copies global(L2-resident) -> shared memory, in the
// access pattern of the tensor-core sweeps: every CTA of a 16-CTA cluster streams the SAME activation image
// (A_BYTES per layer) plus its OWN weight slice (W_BYTES per layer) through a 2-stage ring.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingest_probe tests/csrc/tma_ingest_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) {
    uint32_t ok = 0; uint32_t spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}

// nchunks chunks per layer; per chunk: a_chunk bytes of the shared image + w_chunk bytes of the CTA's slice
__global__ void __launch_bounds__(256, 1) ingest(const char *A, const char *W, int nchunks, int a_chunk, int w_chunk, int nstage,
                                                 int layers, int pieces, long long *out) {
    extern __shared__ __align__(128) char sm[];
    __shared__ __align__(8) uint64_t full[4], empty[4];
    const int tid = threadIdx.x;
    const int stage_bytes = a_chunk + w_chunk;
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (tid == 0) {                       // producer
        int s = 0; uint32_t par = 0; long long issued = 0;
        for (int l = 0; l < layers; ++l)
            for (int c = 0; c < nchunks; ++c) {
                if (issued >= nstage) mbar_wait(&empty[s], par ^ 1u);
                mbar_expect(&full[s], (uint32_t)stage_bytes);
                char *dst = sm + (size_t)s * stage_bytes;
                const int ap = a_chunk / pieces;
                for (int q = 0; q < pieces; ++q) bulk(dst + q * ap, A + (size_t)c * a_chunk + q * ap, ap, &full[s]);
                bulk(dst + a_chunk, W + ((size_t)blockIdx.x * nchunks + c) * w_chunk, w_chunk, &full[s]);
                ++issued;
                if (++s == nstage) { s = 0; par ^= 1u; }
            }
    }
    if (tid >= 32) {                      // 7 consumer warps + ... (8 arrivals: warps 1..7 and warp 0's lane 31 group below)
        int s = 0; uint32_t par = 0;
        for (int l = 0; l < layers; ++l)
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&full[s], par);
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty[s]);
                if (tid == 32) mbar_arrive(&empty[s]);       // the 8th arrival
                if (++s == nstage) { s = 0; par ^= 1u; }
            }
    }
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
    const size_t ABYTES = 1 << 20, WBYTES = 64 << 20;
    char *A, *W; long long *out;
    cudaMalloc(&A, ABYTES); cudaMalloc(&W, WBYTES); cudaMalloc(&out, 1024 * 8);
    cudaMemset(A, 0, ABYTES); cudaMemset(W, 0, WBYTES);
    cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(ingest, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    struct Cfg { const char *name; int nch, a, w, st, pieces; };
    const Cfg cfgs[] = {
        {"c5 tf32x3: A 512K + W 128K, 8 chunks, 2 stages", 8, 65536, 16384, 2, 2},
        {"c5 tf32x3: 16 chunks of 40K, 4 stages", 16, 32768, 8192, 4, 2},
        {"c5 fp32 A: A 256K + W 128K, 8 chunks", 8, 32768, 16384, 2, 1},
        {"c5 bf16x3: A 384K + W 96K, 8 chunks", 8, 49152, 12288, 2, 3},
        {"c3 tf32x3: A 205K + W 26K, 5 chunks", 5, 40960, 5120, 2, 2},
        {"W only 128K", 8, 1024, 16384, 2, 1},
    };
    const int layers = 200;
    for (int csz : {1, 16}) {
        for (int nclusters : {1, 2, 8}) {
            if (csz == 1 && nclusters != 1) continue;
            for (const Cfg &c : cfgs) {
                const int grid = (csz == 1 ? 16 : csz) * nclusters;
                cudaLaunchConfig_t cfg = {};
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256);
                cfg.dynamicSmemBytes = (size_t)c.st * (c.a + c.w);
                cfg.attrs = attr; cfg.numAttrs = 1;
                cudaError_t e = cudaLaunchKernelEx(&cfg, ingest, (const char *)A, (const char *)W, c.nch, c.a, c.w, c.st, layers, c.pieces, out);
                cudaError_t e2 = cudaDeviceSynchronize();
                if (e != cudaSuccess || e2 != cudaSuccess) { printf("%s: launch error %s / %s\n", c.name, cudaGetErrorString(e), cudaGetErrorString(e2)); continue; }
                long long h[1024]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                const double bytes = (double)c.nch * (c.a + c.w);
                printf("cluster %2d x %d clusters (%3d CTAs)  %-48s  %8.0f cycles/layer  %6.1f B/clk/SM  (%.0f KB/layer)\n", csz, nclusters, grid,
                       c.name, (double)mx / layers, bytes * layers / (double)mx, bytes / 1024);
            }
        }
    }
    return 0;
}
