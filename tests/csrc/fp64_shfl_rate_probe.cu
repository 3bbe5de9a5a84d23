// Latency / throughput of the instructions the in-kernel moment matching leans on (B200): dependent DADD / DFMA
// chains, independent DFMA streams, IEEE double and float division, SHFL.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_shfl_rate_probe fp64_shfl_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double *out, long long *cyc, int iters) {
    const int tid = threadIdx.x;
    double a = 1.0 + tid * 1e-9, b = 1.0000001, c = 0.5;
    long long t0, t1;
    // dependent DADD chain
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < iters; ++i) a += b;
    __syncthreads(); t1 = clock64(); if (tid == 0) cyc[0] = t1 - t0;
    // dependent DFMA chain
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, b, c);
    __syncthreads(); t1 = clock64(); if (tid == 0) cyc[1] = t1 - t0;
    // 8 independent DFMA streams per thread
    double v[8]; for (int k = 0; k < 8; ++k) v[k] = a + k;
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fma(v[k], b, c);
    __syncthreads(); t1 = clock64(); if (tid == 0) cyc[2] = t1 - t0;
    for (int k = 0; k < 8; ++k) a += v[k];
    // dependent double division
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < iters; ++i) a = c / a + b;
    __syncthreads(); t1 = clock64(); if (tid == 0) cyc[3] = t1 - t0;
    // dependent float division + sqrt
    float f = (float)a;
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < iters; ++i) f = 0.5f / f + 1.0000001f;
    __syncthreads(); t1 = clock64(); if (tid == 0) cyc[4] = t1 - t0;
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < iters; ++i) f = sqrtf(f) + 1.0000001f;
    __syncthreads(); t1 = clock64(); if (tid == 0) cyc[5] = t1 - t0;
    // dependent SHFL chain and 8 independent SHFL streams
    float s = f;
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < iters; ++i) s += __shfl_xor_sync(0xffffffffu, s, 1);
    __syncthreads(); t1 = clock64(); if (tid == 0) cyc[6] = t1 - t0;
    float w[8]; for (int k = 0; k < 8; ++k) w[k] = s + k;
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __shfl_xor_sync(0xffffffffu, w[k], 1 + (k & 3));
    __syncthreads(); t1 = clock64(); if (tid == 0) cyc[7] = t1 - t0;
    for (int k = 0; k < 8; ++k) s += w[k];
    out[blockIdx.x * blockDim.x + tid] = a + f + s;
}
int main() {
    double *out; long long *cyc; long long h[8];
    cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 64);
    const int iters = 1000;
    for (int nt : {32, 256}) {
        probe<<<1, nt>>>(out, cyc, iters); probe<<<1, nt>>>(out, cyc, iters);
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("%3d threads/SM: DADD chain %.1f cyc/op, DFMA chain %.1f, DFMA x8 independent %.2f cyc/warp-instr/SM, ddiv chain %.1f, "
               "fdiv chain %.1f, fsqrt chain %.1f, SHFL chain %.1f, SHFL x8 independent %.2f cyc/warp-instr/SM\n",
               nt, h[0] / (double)iters, h[1] / (double)iters, h[2] / (8.0 * iters * (nt / 32)), h[3] / (double)iters, h[4] / (double)iters,
               h[5] / (double)iters, h[6] / (double)iters, h[7] / (8.0 * iters * (nt / 32)));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
