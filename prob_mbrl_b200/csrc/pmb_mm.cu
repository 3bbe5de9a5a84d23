// Moment matching of the rewards (reference utils/rollout.py:135-145): a 1x1 instance of mm_resample_,
//   r' = mean r + zhat * sqrt(var_unbiased(r) + 1e-12),   zhat = standardised z_rr[(t + n) mod N].
// Nothing in the recurrence consumes rewards, so both directions run as whole-horizon kernels off the
// serial chain: one CTA per (step, group).
#include "pmb_host.h"
#include "pmb_internal.cuh"

namespace pmb {

__device__ __forceinline__ double block_sum(double v, double *sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
    return t;
}

// rstat[t][g] = {mean r, sigma, mean z, 1/std z}
__global__ void __launch_bounds__(128) reward_mm_fwd_kernel(const float *__restrict__ rpre, float *__restrict__ rout,
                                                            const float *__restrict__ z_rr, float *__restrict__ rstat,
                                                            int N, int G, int *status) {
    __shared__ double sm[4];
    const int t = blockIdx.x, g = blockIdx.y, Ng = N / G, base = g * Ng;
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        a += (double)rpre[(size_t)t * N + base + i];
        b += (double)z_rr[(t + base + i) % N];
    }
    const double mr = block_sum(a, sm) / Ng, mz = block_sum(b, sm) / Ng;
    a = b = 0.0;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        const double dr = (double)rpre[(size_t)t * N + base + i] - mr;
        const double dz = (double)z_rr[(t + base + i) % N] - mz;
        a += dr * dr;
        b += dz * dz;
    }
    const float var = (float)(block_sum(a, sm) / (Ng - 1)) + 1e-12f;
    const float zistd = 1.f / sqrtf((float)(block_sum(b, sm) / (Ng - 1)));
    if (!(var > 0.f) && threadIdx.x == 0 && status) atomicCAS(status, 0, 1 + t);
    const float sigma = sqrtf(var);
    if (threadIdx.x == 0) {
        float *st = rstat + ((size_t)t * G + g) * 4;
        st[0] = (float)mr; st[1] = sigma; st[2] = (float)mz; st[3] = zistd;
    }
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        const float zh = (z_rr[(t + base + i) % N] - (float)mz) * zistd;
        rout[(size_t)t * N + base + i] = (float)mr + zh * sigma;
    }
}

// g_k = (1/M) sum_n g'_n + (sum_n g'_n zhat_n) (r_k - mean) / ((M-1) sigma)
__global__ void __launch_bounds__(128) reward_mm_bwd_kernel(const float *__restrict__ gout, const float *__restrict__ rpre,
                                                            const float *__restrict__ z_rr, const float *__restrict__ rstat,
                                                            float *__restrict__ gin, int N, int G) {
    __shared__ double sm[4];
    const int t = blockIdx.x, g = blockIdx.y, Ng = N / G, base = g * Ng;
    const float *st = rstat + ((size_t)t * G + g) * 4;
    const float mr = st[0], sigma = st[1], mz = st[2], zistd = st[3];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < Ng; i += blockDim.x) {
        const double go = (double)gout[(size_t)t * N + base + i];
        a += go;
        b += go * (double)((z_rr[(t + base + i) % N] - mz) * zistd);
    }
    const float sa = (float)(block_sum(a, sm) / Ng);
    const float sb = (float)(block_sum(b, sm) / ((double)(Ng - 1) * (double)sigma));
    for (int i = threadIdx.x; i < Ng; i += blockDim.x)
        gin[(size_t)t * N + base + i] = sa + sb * (rpre[(size_t)t * N + base + i] - mr);
}

cudaError_t launch_reward_mm_fwd(const float *rpre, float *rout, const float *z_rr, float *rstat, int N, int H, int G,
                                 int *status, cudaStream_t stream) {
    dim3 grid(H, G);
    reward_mm_fwd_kernel<<<grid, 128, 0, stream>>>(rpre, rout, z_rr, rstat, N, G, status);
    return cudaGetLastError();
}

cudaError_t launch_reward_mm_bwd(const float *gout, const float *rpre, const float *z_rr, const float *rstat, float *gin,
                                 int N, int H, int G, cudaStream_t stream) {
    dim3 grid(H, G);
    reward_mm_bwd_kernel<<<grid, 128, 0, stream>>>(gout, rpre, z_rr, rstat, gin, N, G);
    return cudaGetLastError();
}

}  // namespace pmb
