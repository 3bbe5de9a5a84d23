// Whole-GPU kernels around the sweeps:
//   * pack_kernel      : lays weights / biases / dropout masks out the way the sweeps stream them
//                        (transposed + zero-padded to multiples of 4 columns);
//   * wgrad_tile_kernel: batched policy weight gradient  dW[m][n] = sum_r delta[r][m] * inp[r][n]
//                        over the r = (step, particle) axis, split-K across the grid -- the only piece of
//                        the backward pass that is not on the sequential chain;
//                        bias gradients ride along as an implicit column of ones;
//   * reduce_partials_kernel: fixed-order sum of the split-K partials (deterministic, no atomics).
// Replaces the weight-gradient part of loss.backward() (reference algorithms/mc_pilco.py:197).
#include "pmb_internal.cuh"
#include "pmb_host.h"

namespace pmb {

// ---------------------------------------------------------------------------------------------
__global__ void pack_kernel(const __grid_constant__ PackJobs jobs) {
    const PackJob &j = jobs.job[blockIdx.y];
    const long long total = (long long)j.R * j.C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int r = (int)(i / j.C), c = (int)(i - (long long)r * j.C);
        int sr = j.transpose ? c : r, sc = j.transpose ? r : c;
        float v = 0.f;
        if (sr < j.SR && sc < j.SC) v = __ldg(j.src + (long long)sr * j.src_ld + sc);
        j.dst[i] = v;
    }
}

cudaError_t launch_pack(const PackJobs &jobs, cudaStream_t stream) {
    if (jobs.n == 0) return cudaSuccess;
    dim3 grid(64, jobs.n);
    pack_kernel<<<grid, 256, 0, stream>>>(jobs);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Partial over rows [r0, r1) of  C = A^T [B | 1]:  A:[R][lda] (M columns = layer output adjoints),
// B:[R][ldb] (Nc columns = layer inputs), plus an implicit column of ones when bias_out != nullptr, so
// the bias gradient (column sums of A) falls out of the same pass:
//   w_out[m][n] = sum_r A[r][m] B[r][n]   (n < Nc, row stride Nc),   bias_out[m] = sum_r A[r][m].
// 64x64 output tile per CTA, 4x4 per thread on the packed FP32 pipe (FFMA2), register-staged double
// buffering of the 16-row k-slab.
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) wgrad_tile_kernel(const float *__restrict__ A, int lda, int M,
                                                         const float *__restrict__ B, int ldb, int Nc,
                                                         long long R, int nsplit, float *__restrict__ w_out,
                                                         float *__restrict__ bias_out, long long part_stride) {
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int ncols = Nc + (bias_out ? 1 : 0);
    const int tiles_n = (ncols + BN - 1) / BN;
    const int tm = blockIdx.x / tiles_n, tn = blockIdx.x - tm * tiles_n;
    const int m0 = tm * BM, n0 = tn * BN;
    const long long rows_per = (R + nsplit - 1) / nsplit;
    const long long r0 = (long long)blockIdx.y * rows_per;
    const long long r1 = min(R, r0 + rows_per);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads, 4x4 outputs each
    const int lc = tid & 63, lr = tid >> 6;      // loader: column 0..63, rows lr, lr+4, lr+8, lr+12
    float2 acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
    const bool a_ok = m0 + lc < M;
    const int bcol = n0 + lc;
    const int b_kind = bcol < Nc ? 0 : (bcol == Nc && bias_out) ? 1 : 2;   // data, ones, padding
    float ra[4], rb[4];
    auto gload = [&](long long rbase) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = rbase + lr + 4 * i;
            const bool okr = r < r1;
            ra[i] = (okr && a_ok) ? __ldg(A + r * lda + m0 + lc) : 0.f;
            rb[i] = !okr ? 0.f : b_kind == 0 ? __ldg(B + r * ldb + bcol) : b_kind == 1 ? 1.f : 0.f;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            As[buf][lr + 4 * i][lc] = ra[i];
            Bs[buf][lr + 4 * i][lc] = rb[i];
        }
    };
    int buf = 0;
    if (r0 < r1) {
        gload(r0);
        sstore(0);
    }
    __syncthreads();
    for (long long rb0 = r0; rb0 < r1; rb0 += BK) {
        const bool more = rb0 + BK < r1;
        if (more) gload(rb0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[buf][k][4 * ty]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][k][4 * tx]);
            const float2 b01 = make_float2(b.x, b.y), b23 = make_float2(b.z, b.w);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 a2 = make_float2(av[i], av[i]);
                acc[i][0] = __ffma2_rn(a2, b01, acc[i][0]);
                acc[i][1] = __ffma2_rn(a2, b23, acc[i][1]);
            }
        }
        if (more) sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    float *wo = w_out + (long long)blockIdx.y * part_stride;
    float *bo = bias_out ? bias_out + (long long)blockIdx.y * part_stride : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + 4 * ty + i;
        if (m >= M) continue;
        const float v[4] = {acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 4 * tx + j;
            if (n < Nc) wo[(long long)m * Nc + n] = v[j];
            else if (n == Nc && bo) bo[m] = v[j];
        }
    }
}

__global__ void reduce_partials_kernel(const float *__restrict__ part, long long n, int nsplit,
                                       float *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nsplit; ++k) s += __ldg(part + (long long)k * n + i);
        out[i] = s;
    }
}

// partials of one linear layer: A = output adjoints [R][lda] (M columns), B = layer input [R][ldb] (Nc columns)
cudaError_t launch_wgrad(const float *A, int lda, int M, const float *B, int ldb, int Nc, long long R,
                         int nsplit, float *w_part, float *bias_part, long long part_stride, cudaStream_t stream) {
    const int ncols = Nc + (bias_part ? 1 : 0);
    dim3 grid(((M + BM - 1) / BM) * ((ncols + BN - 1) / BN), nsplit);
    wgrad_tile_kernel<<<grid, 256, 0, stream>>>(A, lda, M, B, ldb, Nc, R, nsplit, w_part, bias_part, part_stride);
    return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const float *part, long long n, int nsplit, float *out, cudaStream_t stream) {
    long long nb = (n + 255) / 256;
    int blocks = (int)(nb < 592 ? nb : 592);
    reduce_partials_kernel<<<blocks, 256, 0, stream>>>(part, n, nsplit, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// clip_grad_norm_ + Adam (reference algorithms/mc_pilco.py:209-214)
// ---------------------------------------------------------------------------------------------
constexpr int NORM_BLOCKS = 128;

__global__ void __launch_bounds__(256) gradnorm_kernel(const pmb_adam_tensor *__restrict__ tab, int nt,
                                                       float *__restrict__ scratch, long long *step_dev) {
    __shared__ float sm[8];
    if (step_dev && blockIdx.x == 0 && threadIdx.x == 0) step_dev[0] += 1;   // read by adam_kernel (next launch)
    float s = 0.f;
    for (int t = 0; t < nt; ++t) {
        const float *g = tab[t].grad;
        const long long n = tab[t].n;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += (long long)gridDim.x * blockDim.x) {
            float v = g[i];
            s = fmaf(v, v, s);
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int i = 0; i < 8; ++i) tot += sm[i];
        scratch[1 + blockIdx.x] = tot;
    }
}

__global__ void __launch_bounds__(256) adam_kernel(const pmb_adam_tensor *__restrict__ tab, int nt, float max_norm,
                                                   float beta1, float beta2, float eps, float lr, float step_size,
                                                   float bc2_sqrt, const long long *__restrict__ step_dev,
                                                   float *__restrict__ scratch, int norm_blocks) {
    __shared__ float s_coef, s_step_size, s_bc2_sqrt;
    if (threadIdx.x == 0) {
        if (step_dev) {   // bias corrections from the device-side step counter (CUDA-graph replay)
            double st = (double)step_dev[0];
            s_step_size = (float)((double)lr / (1.0 - pow((double)beta1, st)));
            s_bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, st));
        } else {
            s_step_size = step_size;
            s_bc2_sqrt = bc2_sqrt;
        }
        float tot = 0.f;
        for (int i = 0; i < norm_blocks; ++i) tot += scratch[1 + i];   // fixed order: deterministic
        float norm = sqrtf(tot);
        float coef = 1.f;
        if (max_norm > 0.f) {
            coef = max_norm / (norm + 1e-6f);     // torch.nn.utils.clip_grad_norm_
            if (coef > 1.f) coef = 1.f;
        }
        s_coef = coef;
        if (blockIdx.x == 0) scratch[0] = norm;
    }
    __syncthreads();
    const float coef = s_coef;
    step_size = s_step_size;
    bc2_sqrt = s_bc2_sqrt;
    for (int t = 0; t < nt; ++t) {
        float *p = tab[t].param, *m = tab[t].exp_avg, *v = tab[t].exp_avg_sq;
        const float *gp = tab[t].grad;
        const long long n = tab[t].n;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += (long long)gridDim.x * blockDim.x) {
            float g = gp[i] * coef;
            float mi = m[i], vi = v[i];
            mi = mi + (g - mi) * (1.f - beta1);                 // exp_avg.lerp_(grad, 1 - beta1)
            vi = vi * beta2 + (1.f - beta2) * g * g;            // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
            float denom = sqrtf(vi) / bc2_sqrt + eps;
            m[i] = mi;
            v[i] = vi;
            p[i] = p[i] - step_size * (mi / denom);
            const_cast<float *>(gp)[i] = g;                     // clip_grad_norm_ scales .grad in place
        }
    }
}

cudaError_t launch_clip_adam(const pmb_adam_tensor *tab, int nt, float max_norm, float lr, float beta1, float beta2,
                             float eps, long long step, long long *step_dev, float *scratch,
                             cudaStream_t stream) {
    gradnorm_kernel<<<NORM_BLOCKS, 256, 0, stream>>>(tab, nt, scratch, step_dev);
    if (step < 1) step = 1;
    double bc1 = 1.0 - pow((double)beta1, (double)step);
    double bc2 = 1.0 - pow((double)beta2, (double)step);
    float step_size = (float)((double)lr / bc1);
    float bc2_sqrt = (float)sqrt(bc2);
    adam_kernel<<<NORM_BLOCKS, 256, 0, stream>>>(tab, nt, max_norm, beta1, beta2, eps, lr, step_size, bc2_sqrt,
                                                 step_dev, scratch, NORM_BLOCKS);
    return cudaGetLastError();
}

}  // namespace pmb
