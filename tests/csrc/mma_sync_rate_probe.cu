// Throughput of the legacy warp-level mma.sync.m16n8k8 TF32 path on sm_100a (cycles per MMA per SM) with 4, 8, 16
// warps per CTA and 4 / 8 independent accumulators per warp; FFMA2 reference loop for comparison.
// nvcc -gencode arch=compute_100a,code=sm_100a -o mma_sync_rate_probe mma_sync_rate_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NACC>
__global__ void mma_loop(float *out, int iters, long long *cyc) {
    float d[NACC][4];
    for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u};
    uint32_t b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void ffma2_loop(float *out, int iters, long long *cyc) {
    float2 acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
    float2 w = make_float2(1.0001f, 0.9999f);
    float x = 1.f + threadIdx.x * 1e-6f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(make_float2(x, x), w, acc[i]);
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    for (int nt : {128, 256, 512}) {
        long long c4, c8;
        mma_loop<4><<<148, nt>>>(out, iters, cyc); cudaMemcpy(&c4, cyc, 8, cudaMemcpyDeviceToHost);
        mma_loop<4><<<148, nt>>>(out, iters, cyc); cudaMemcpy(&c4, cyc, 8, cudaMemcpyDeviceToHost);
        mma_loop<8><<<148, nt>>>(out, iters, cyc); cudaMemcpy(&c8, cyc, 8, cudaMemcpyDeviceToHost);
        const double m4 = (double)(nt / 32) * 4 * iters, m8 = (double)(nt / 32) * 8 * iters;
        printf("mma.sync m16n8k8 tf32, %2d warps/SM: 4 acc: %.2f cycles/MMA/SM (%.0f FMA/clk/SM); 8 acc: %.2f (%.0f FMA/clk/SM)\n", nt / 32,
               c4 / m4, 1024.0 * m4 / c4, c8 / m8, 1024.0 * m8 / c8);
    }
    for (int nt : {128, 512}) {
        long long c;
        ffma2_loop<<<148, nt>>>(out, iters, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        ffma2_loop<<<148, nt>>>(out, iters, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("FFMA2, %2d warps/SM: %.0f FMA/clk/SM\n", nt / 32, (double)(nt) * 16 * 2 * iters / c);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
