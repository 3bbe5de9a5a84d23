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

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM version of the same contraction for the hidden x hidden layers (the one genuinely dense
// GEMM of the path: [M x R] . [R x N] with R = H*N rows).  fp32 parity on the tensor pipe needs a split:
// x = hi + lo with hi = tf32(x); C += A_hi B_hi + A_hi B_lo + A_lo B_hi (3 x kind::tf32, fp32 accumulate in
// TMEM; the dropped lo*lo term is ~2^-22 relative).  In global memory the reduction index r is the slow axis
// of both delta[r][m] and x[r][n] (MN-major), but kind::tf32 reads MN-major no-swizzle operands as zeros on
// this part (tests/csrc/umma_probe.cu, profiles/r01_umma_layout_probe.txt), so the loader transposes while
// staging into the K-major no-swizzle canonical layout the probe confirmed: core matrix = 8 MN rows x 16 B
// (4 k), the two k halves of a K=8 MMA 128 B apart (LBO), groups of 8 MN rows 256 B apart (SBO).
// One CTA = one 128-row M tile x all N (<= 256) columns x one split-K slice; 256 threads load, split and
// store; one elected thread issues the MMAs; 4 warps drain TMEM with tcgen05.ld.
// ---------------------------------------------------------------------------------------------
constexpr int UM = 128;        // UMMA M
constexpr int UKB = 4;         // k-blocks (of 8 rows) per stage

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);   // version 1 (sm_100), SWIZZLE_NONE
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(256, 1) wgrad_umma_kernel(const float *__restrict__ A, int lda, int M,
                                                            const float *__restrict__ B, int ldb, int Nc,
                                                            long long R, int nsplit, float *__restrict__ w_out,
                                                            float *__restrict__ bias_out, long long part_stride,
                                                            int n16, int ncw) {
    extern __shared__ __align__(128) float usm[];
    __shared__ __align__(8) uint64_t stage_free[2], done_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * UM;
    const long long rows_per = (((R + nsplit - 1) / nsplit) + 7) & ~7LL;
    const long long r0 = (long long)blockIdx.y * rows_per;
    const long long r1 = min(R, r0 + rows_per);
    const int a_kb = UM * 8;                 // floats per k-block of A
    const int b_kb = n16 * 8;                // floats per k-block of B
    const int stage_floats = UKB * 2 * (a_kb + b_kb);
    const int ncols = Nc + (bias_out ? 1 : 0);
    // column chunk of this CTA (blockIdx.z): global columns n_off .. n_off + ncw - 1 of [B | 1]; the accumulator holds
    // <= 256 columns, wider layers (c4: 400, c5: 512 inputs) take several chunks
    const int n_off = blockIdx.z * ncw;

    if (tid == 0) {
        mbar_init(&stage_free[0], 1);
        mbar_init(&stage_free[1], 1);
        mbar_init(&done_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    // fp32 accumulate, tf32 x tf32, both operands K-major
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n16 >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);

    constexpr int A_ITEMS = UKB * 2 * UM / 256, B_ITEMS = UKB * 2 * 256 / 256;
    int b_n[B_ITEMS], b_rq[B_ITEMS];          // this thread's (column, row quad) items of a B stage
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
        const int i = tid + 256 * j;
        b_n[j] = i % n16;
        b_rq[j] = i / n16;                    // >= 2*UKB: past the stage (n16 < 256)
    }

    int s = 0;
    for (long long rs = r0; rs < r1; rs += 8 * UKB, ++s) {
        const int b = s & 1;
        float *st = usm + (size_t)b * stage_floats;
        float *Ahi = st, *Alo = st + UKB * a_kb, *Bhi = st + 2 * UKB * a_kb, *Blo = Bhi + UKB * b_kb;
        // Lanes run over the MN index (coalesced 128 B row segments), each thread gathers 4 consecutive k rows of
        // one column and stores them as one 16 B K-major core-matrix row (conflict-free per quarter warp).
        // All of a stage's loads are issued before the first conversion so ~44 requests per thread are in flight.
        float va[A_ITEMS][4], vb[B_ITEMS][4];
#pragma unroll
        for (int j = 0; j < A_ITEMS; ++j) {
            const int i = tid + 256 * j, ml = i % UM, rq = i / UM;
            const long long r = rs + 4 * rq;
            const int m = m0 + ml;
#pragma unroll
            for (int e = 0; e < 4; ++e) va[j][e] = (m < M && r + e < r1) ? __ldg(A + (r + e) * lda + m) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < B_ITEMS; ++j) {
            const long long r = rs + 4 * b_rq[j];
            const int n = b_n[j];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                vb[j][e] = (b_rq[j] < 2 * UKB && r + e < r1 && n < ncw)
                               ? (n_off + n < Nc ? __ldg(B + (r + e) * ldb + n_off + n) : ((n_off + n == Nc && bias_out) ? 1.f : 0.f))
                               : 0.f;
        }
        if (s >= 2) mbar_wait(&stage_free[b], ((s >> 1) - 1) & 1);   // the MMAs that read this buffer retired
#pragma unroll
        for (int j = 0; j < A_ITEMS; ++j) {
            const int i = tid + 256 * j, ml = i % UM, rq = i / UM;
            const float4 h = make_float4(tf32_hi(va[j][0]), tf32_hi(va[j][1]), tf32_hi(va[j][2]), tf32_hi(va[j][3]));
            const int off = (rq >> 1) * a_kb + (ml >> 3) * 64 + (rq & 1) * 32 + (ml & 7) * 4;
            *reinterpret_cast<float4 *>(Ahi + off) = h;
            *reinterpret_cast<float4 *>(Alo + off) = make_float4(va[j][0] - h.x, va[j][1] - h.y, va[j][2] - h.z, va[j][3] - h.w);
        }
#pragma unroll
        for (int j = 0; j < B_ITEMS; ++j) {
            if (b_rq[j] < 2 * UKB) {
                const int n = b_n[j], rq = b_rq[j];
                const float4 h = make_float4(tf32_hi(vb[j][0]), tf32_hi(vb[j][1]), tf32_hi(vb[j][2]), tf32_hi(vb[j][3]));
                const int off = (rq >> 1) * b_kb + (n >> 3) * 64 + (rq & 1) * 32 + (n & 7) * 4;
                *reinterpret_cast<float4 *>(Bhi + off) = h;
                *reinterpret_cast<float4 *>(Blo + off) = make_float4(vb[j][0] - h.x, vb[j][1] - h.y, vb[j][2] - h.z, vb[j][3] - h.w);
            }
        }
        fence_proxy_async();            // generic-proxy stores -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long left = (r1 - rs + 7) / 8;
            const int nkb = left < UKB ? (int)left : UKB;
            for (int kb = 0; kb < nkb; ++kb) {
                const uint64_t dAh = umma_desc(smem_u32(Ahi + kb * a_kb), 128, 256);
                const uint64_t dAl = umma_desc(smem_u32(Alo + kb * a_kb), 128, 256);
                const uint64_t dBh = umma_desc(smem_u32(Bhi + kb * b_kb), 128, 256);
                const uint64_t dBl = umma_desc(smem_u32(Blo + kb * b_kb), 128, 256);
                umma_tf32(tmem_d, dAh, dBh, idesc, (s | kb) ? 1u : 0u);
                umma_tf32(tmem_d, dAh, dBl, idesc, 1u);
                umma_tf32(tmem_d, dAl, dBh, idesc, 1u);
            }
            umma_commit(&stage_free[b]);
        }
    }
    if (tid == 0) umma_commit(&done_bar);
    float *wo = w_out + (long long)blockIdx.y * part_stride;
    float *bo = bias_out ? bias_out + (long long)blockIdx.y * part_stride : nullptr;
    if (s == 0) {           // empty slice: contribute zeros
        for (int i = tid; i < UM * ncw; i += 256) {
            const int m = m0 + i / ncw, n = n_off + i % ncw;
            if (m < M && n < ncols) { if (n < Nc) wo[(long long)m * Nc + n] = 0.f; else bo[m] = 0.f; }
        }
    } else {
        mbar_wait(&done_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp < 4) {         // warp w drains TMEM lanes 32w .. 32w+31 = rows m0 + 32w + lane
            const int m = m0 + 32 * warp + lane;
            for (int c = 0; c < n16; c += 8) {
                uint32_t v[8];
                const uint32_t taddr = tmem_d + ((uint32_t)(32 * warp) << 16) + (uint32_t)c;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m < M) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int n = n_off + c + e;
                        if (c + e >= ncw) continue;
                        if (n < Nc) wo[(long long)m * Nc + n] = __uint_as_float(v[e]);
                        else if (n == Nc && bo) bo[m] = __uint_as_float(v[e]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_d) : "memory");
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
                         int nsplit, float *w_part, float *bias_part, long long part_stride, cudaStream_t stream,
                         int use_umma) {
    const int ncols = Nc + (bias_part ? 1 : 0);
    // mode 0 (auto): tensor cores while one split-K slice accumulates <= 1024 rows in TMEM -- the tensor core's
    // fp32 accumulation rounds coarser than FFMA (measured 2.1e-6 vs 3.7e-7 gradient error at 625 rows per
    // slice against a 1e-5 budget); 1: always; 2: never.
    const bool umma = use_umma == 1 || (use_umma == 0 && (R + nsplit - 1) / nsplit <= 1024);
    if (umma && M > 16 && Nc > 16) {
        // <= 256 accumulator columns per CTA: wider layers are cut into equal column chunks (grid.z)
        const int nchunks = (ncols + 255) / 256;
        const int ncw = (((ncols + nchunks - 1) / nchunks) + 7) & ~7;
        const int n16 = (ncw + 15) & ~15;
        const int smem = 2 * UKB * 2 * (UM * 8 + n16 * 8) * (int)sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        dim3 grid((M + UM - 1) / UM, nsplit, (ncols + ncw - 1) / ncw);
        wgrad_umma_kernel<<<grid, 256, smem, stream>>>(A, lda, M, B, ldb, Nc, R, nsplit, w_part, bias_part, part_stride, n16, ncw);
        return cudaGetLastError();
    }
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
                                                       float *__restrict__ scratch, long long *step_dev,
                                                       const int *__restrict__ skip) {
    __shared__ float sm[8];
    if (skip && *skip != 0) return;      // failed rollout (status word): no update at all
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
                                                   float *__restrict__ scratch, int norm_blocks,
                                                   const int *__restrict__ skip) {
    __shared__ float s_coef, s_step_size, s_bc2_sqrt;
    if (skip && *skip != 0) return;
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
                             float eps, long long step, long long *step_dev, float *scratch, const int *skip,
                             cudaStream_t stream) {
    gradnorm_kernel<<<NORM_BLOCKS, 256, 0, stream>>>(tab, nt, scratch, step_dev, skip);
    if (step < 1) step = 1;
    double bc1 = 1.0 - pow((double)beta1, (double)step);
    double bc2 = 1.0 - pow((double)beta2, (double)step);
    float step_size = (float)((double)lr / bc1);
    float bc2_sqrt = (float)sqrt(bc2);
    adam_kernel<<<NORM_BLOCKS, 256, 0, stream>>>(tab, nt, max_norm, beta1, beta2, eps, lr, step_size, bc2_sqrt,
                                                 step_dev, scratch, NORM_BLOCKS, skip);
    return cudaGetLastError();
}

}  // namespace pmb
