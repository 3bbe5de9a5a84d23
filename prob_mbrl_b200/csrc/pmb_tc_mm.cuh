// Moment matching of the states inside the tensor-core cluster sweeps (reference utils/rollout.py:20-29,121-132):
//   x' = m + zhat chol(S)^T,  m = mean_n x,  S = (x-m)^T (x-m)/(M-1) + 1e-12 I,
//   zhat = (z - mean z) / std_unbiased(z) per column, z = z_mm[(t + n) mod N] (rollout.py:53-59).
// Every CTA of the cluster holds the state of ALL particles of the tile (and the tile holds every matching group
// completely), so each CTA evaluates the group statistics redundantly from its own shared memory: no exchange, no
// barrier beyond the CTA's own, and -- fixed summation order -- bit-identical results on every CTA.  Same
// arithmetic as the streaming variant (pmb_mm.cuh): fp64 sums, the 1e-12-jittered Cholesky in fp32 like the
// reference's, a non-positive pivot reported through the status word (reference: cholesky() raises,
// rollout.py:25,154-157).  The reverse step is the hand-derived adjoint of oracle/rollout_oracle.py::mm_backward.
// Called by the 256 compute threads only (CTA_SYNC = their named barrier).
#pragma once
#include "pmb_tc.cuh"

namespace pmb {

// scratch layout (floats)
constexpr int TCMM_ZT = 0;                                   // [128][TC_SDP] the z_mm table (rows 0..N-1), loaded once
constexpr int TCMM_ZS = TCMM_ZT + TC_M * TC_SDP;             // [128][TC_SDP] z rows of the tile at this step (rotated)
constexpr int TCMM_XS = TCMM_ZS + TC_M * TC_SDP;             // [128][TC_SDP] pre-matching particles (reverse)
constexpr int TCMM_DBL = TCMM_XS + TC_M * TC_SDP;            // doubles: per group 2*SD means
constexpr int TCMM_MAXG = 4;                                 // groups per tile
constexpr int TCMM_GST = TCMM_DBL + 2 * (TCMM_MAXG * 2 * SD);   // per group: m[SD], zm[SD], zistd[SD], L[SD*SD], A, X, Sb, dm[SD]
constexpr int TCMM_GSTRIDE = 3 * SD + 4 * SD * SD + SD;
constexpr int TCMM_FLOATS = TCMM_GST + TCMM_MAXG * TCMM_GSTRIDE;

// sum over the particles of group g of f(particle) in fp64 by `nl` neighbouring lanes (power of two; fixed order)
template <typename F>
__device__ __forceinline__ double tcmm_group_sum(int sub, int nl, int base, int Ng, F f) {
    double a = 0.0;
    for (int i = sub; i < Ng; i += nl) a += f(base + i);
    for (int o = 1; o < nl; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    return a;
}
// lanes per reduction item: as many as keep all items in one pass over the 256 compute threads
__device__ __forceinline__ int tcmm_lanes(int items) { return items <= 8 ? 32 : items <= 16 ? 16 : items <= 32 ? 8 : 4; }

// the z_mm rows 0..N-1 (rollout.py:53-59 only ever reads those), once per kernel
__device__ __forceinline__ void tc_mm_load_table(const TcParams &prm, float *scr) {
    for (int i = threadIdx.x; i < prm.N * prm.D; i += TC_NT) {
        const int r = i / prm.D, d = i - r * prm.D;
        scr[TCMM_ZT + r * TC_SDP + d] = __ldg(prm.z_mm + i);
    }
}

// forward: st[p][d] (pre-matching particles of the tile) -> moment-matched particles, in place
__device__ __forceinline__ void tc_mm_forward(const TcParams &prm, float *st, float *scr, int t, int n0, int nval, int rank) {
    const int tid = threadIdx.x, D = prm.D, N = prm.N, G = max(prm.mm_G, 1), Ng = prm.mm_Ng;
    float *zs = scr + TCMM_ZS;
    double *mean = reinterpret_cast<double *>(scr + TCMM_DBL);
    if (tid < nval) {
        const float *zr = scr + TCMM_ZT + ((t + n0 + tid) % N) * TC_SDP;
        for (int d = 0; d < D; ++d) zs[tid * TC_SDP + d] = zr[d];
    }
    CTA_SYNC();
    int nl = tcmm_lanes(G * 2 * D);
    int sub = tid & (nl - 1), slot = tid / nl, nslot = TC_NT / nl;
    // ---- means of x and z per group (whole warps walk the loop: the shuffles are convergent) ----
    for (int it0 = 0; it0 < G * 2 * D; it0 += nslot) {
        const int it = it0 + slot;
        const bool on = it < G * 2 * D;
        const int g = on ? it / (2 * D) : 0, q = on ? it - g * 2 * D : 0;
        const float *src = q < D ? st + q : zs + (q - D);
        const double s = tcmm_group_sum(sub, nl, g * Ng, on ? Ng : 0, [&](int i) { return (double)src[i * TC_SDP]; });
        if (on && sub == 0) mean[g * 2 * SD + (q < D ? q : SD + q - D)] = s / Ng;
    }
    CTA_SYNC();
    // ---- unbiased covariance (lower triangle) + jitter, z variance ----
    const int nq = D * D + D;
    nl = tcmm_lanes(G * nq);
    sub = tid & (nl - 1); slot = tid / nl; nslot = TC_NT / nl;
    for (int it0 = 0; it0 < G * nq; it0 += nslot) {
        const int it = it0 + slot;
        const bool on = it < G * nq;
        const int g = on ? it / nq : 0, q = on ? it - g * nq : 0;
        float *gs_ = scr + TCMM_GST + g * TCMM_GSTRIDE;
        const double *mg = mean + g * 2 * SD;
        if (q < D * D) {
            const int i = q / D, j = q - i * D;
            const bool low = on && j <= i;
            const double mi = mg[i], mj = mg[j];
            const double s = tcmm_group_sum(sub, nl, g * Ng, low ? Ng : 0, [&](int k) {
                return ((double)st[k * TC_SDP + i] - mi) * ((double)st[k * TC_SDP + j] - mj);
            });
            // rollout.py:24; handed to the fp32 Cholesky (A = 3*SD + SD*SD floats into the group block)
            if (low && sub == 0) gs_[3 * SD + SD * SD + i * SD + j] = (float)(s / (double)(Ng - 1)) + (i == j ? 1e-12f : 0.f);
        } else {
            const int d = q - D * D;
            const double mz = mg[SD + d];
            const double s = tcmm_group_sum(sub, nl, g * Ng, on ? Ng : 0, [&](int k) {
                const double dz = (double)zs[k * TC_SDP + d] - mz;
                return dz * dz;
            });
            if (on && sub == 0) {
                gs_[d] = (float)mg[d];
                gs_[SD + d] = (float)mz;
                gs_[2 * SD + d] = 1.f / sqrtf((float)(s / (double)(Ng - 1)));     // 1 / z.std(unbiased)
            }
        }
    }
    CTA_SYNC();
    // ---- fp32 Cholesky, one thread per group ----
    if (tid < G) {
        float *gs_ = scr + TCMM_GST + tid * TCMM_GSTRIDE;
        float *Lm = gs_ + 3 * SD;
        const float *A = Lm + SD * SD;
        bool ok = true;
        for (int i = 0; i < D; ++i) {
            for (int j = 0; j <= i; ++j) {
                float s = A[i * SD + j];
                for (int k = 0; k < j; ++k) s -= Lm[i * SD + k] * Lm[j * SD + k];
                if (i == j) {
                    if (!(s > 0.f)) { ok = false; s = 1.f; }
                    Lm[i * SD + i] = sqrtf(s);
                } else {
                    Lm[i * SD + j] = s / Lm[j * SD + j];
                }
            }
            for (int j = i + 1; j < D; ++j) Lm[i * SD + j] = 0.f;
        }
        if (!ok && prm.status && rank == 0) atomicCAS(prm.status, 0, 1 + t);
    }
    CTA_SYNC();
    // keep (m, z statistics, L) of this step for the reverse sweep
    if (rank == 0) {
        for (int i = tid; i < G * (3 * SD + SD * SD); i += TC_NT) {
            const int g = i / (3 * SD + SD * SD), k = i - g * (3 * SD + SD * SD);
            prm.mmstat[((size_t)t * G + g) * (3 * SD + SD * SD) + k] = scr[TCMM_GST + g * TCMM_GSTRIDE + k];
        }
    }
    if (tid < nval) {
        const int g = tid / Ng;
        const float *gs_ = scr + TCMM_GST + g * TCMM_GSTRIDE;
        float xo[SD];
        for (int d = 0; d < D; ++d) {
            float x = gs_[d];
            for (int j = 0; j <= d; ++j)
                x = fmaf((zs[tid * TC_SDP + j] - gs_[SD + j]) * gs_[2 * SD + j], gs_[3 * SD + d * SD + j], x);
            xo[d] = x;
        }
        for (int d = 0; d < D; ++d) st[tid * TC_SDP + d] = xo[d];
    }
    CTA_SYNC();
}

// reverse: gs[p][d] holds the cotangent of the moment-matched particles x'; on return the cotangent of the
// pre-matching particles x:  dx_n = dm/M + 2/(M-1) * sym(L^-T Phi(L^T dL) L^-1) (x_n - m),
// dm = sum_n g_n,  dL = tril(sum_n g_n zhat_n^T).
__device__ __forceinline__ void tc_mm_backward(const TcParams &prm, float *gs, float *scr, int t, int n0, int nval) {
    const int tid = threadIdx.x, D = prm.D, N = prm.N, G = max(prm.mm_G, 1), Ng = prm.mm_Ng;
    float *zs = scr + TCMM_ZS, *xs = scr + TCMM_XS;
    const int GS = 3 * SD + SD * SD;
    for (int i = tid; i < G * GS; i += TC_NT) {
        const int g = i / GS, k = i - g * GS;
        scr[TCMM_GST + g * TCMM_GSTRIDE + k] = __ldcg(prm.mmstat + ((size_t)t * G + g) * GS + k);
    }
    if (tid < nval) {
        const float *zr = scr + TCMM_ZT + ((t + n0 + tid) % N) * TC_SDP;
        for (int d = 0; d < D; ++d) zs[tid * TC_SDP + d] = zr[d];
    }
    for (int i = tid; i < nval * D; i += TC_NT)       // pre-matching particles of the step: one coalesced block
        xs[(i / D) * TC_SDP + (i % D)] = __ldcg(prm.s1pre + ((size_t)t * N + n0) * D + i);
    CTA_SYNC();
    const int nq = D * D + D;
    const int nl = tcmm_lanes(G * nq);
    const int sub = tid & (nl - 1), slot = tid / nl, nslot = TC_NT / nl;
    for (int it0 = 0; it0 < G * nq; it0 += nslot) {
        const int it = it0 + slot;
        const bool on = it < G * nq;
        const int g = on ? it / nq : 0, q = on ? it - g * nq : 0;
        float *gs_ = scr + TCMM_GST + g * TCMM_GSTRIDE;
        float *X = gs_ + 3 * SD + 2 * SD * SD, *dm = gs_ + 3 * SD + 4 * SD * SD;
        if (q < D * D) {
            const int i = q / D, j = q - i * D;
            const bool low = on && j <= i;
            const float zm = gs_[SD + j], zi = gs_[2 * SD + j];
            const double s = tcmm_group_sum(sub, nl, g * Ng, low ? Ng : 0, [&](int k) {
                return (double)gs[k * TC_SDP + i] * (double)((zs[k * TC_SDP + j] - zm) * zi);
            });
            if (on && sub == 0) X[i * SD + j] = low ? (float)s : 0.f;          // dL (lower triangle), staged in X
        } else {
            const int d = q - D * D;
            const double s = tcmm_group_sum(sub, nl, g * Ng, on ? Ng : 0, [&](int k) { return (double)gs[k * TC_SDP + d]; });
            if (on && sub == 0) dm[d] = (float)s;
        }
    }
    CTA_SYNC();
    // A = Phi(L^T dL): lower triangle, diagonal halved
    for (int it = tid; it < G * D * D; it += TC_NT) {
        const int g = it / (D * D), q = it - g * D * D, i = q / D, j = q - i * D;
        float *gs_ = scr + TCMM_GST + g * TCMM_GSTRIDE;
        const float *Lm = gs_ + 3 * SD, *X = gs_ + 3 * SD + 2 * SD * SD;
        float *A = gs_ + 3 * SD + SD * SD;
        float a = 0.f;
        if (j <= i) {
            for (int k = i; k < D; ++k) a = fmaf(Lm[k * SD + i], X[k * SD + j], a);   // dL[k][j] = 0 for j > k
            if (i == j) a *= 0.5f;
        }
        A[i * SD + j] = a;
    }
    CTA_SYNC();
    // X = L^-T A  (back substitution, one column per thread)
    for (int it = tid; it < G * D; it += TC_NT) {
        const int g = it / D, j = it - g * D;
        float *gs_ = scr + TCMM_GST + g * TCMM_GSTRIDE;
        const float *Lm = gs_ + 3 * SD, *A = gs_ + 3 * SD + SD * SD;
        float *X = gs_ + 3 * SD + 2 * SD * SD;
        for (int r = D - 1; r >= 0; --r) {
            float s = A[r * SD + j];
            for (int k = r + 1; k < D; ++k) s -= Lm[k * SD + r] * X[k * SD + j];
            X[r * SD + j] = s / Lm[r * SD + r];
        }
    }
    CTA_SYNC();
    // Sb = X L^-1  (one row per thread)
    for (int it = tid; it < G * D; it += TC_NT) {
        const int g = it / D, i = it - g * D;
        float *gs_ = scr + TCMM_GST + g * TCMM_GSTRIDE;
        const float *Lm = gs_ + 3 * SD, *X = gs_ + 3 * SD + 2 * SD * SD;
        float *Sb = gs_ + 3 * SD + 3 * SD * SD;
        for (int c = D - 1; c >= 0; --c) {
            float s = X[i * SD + c];
            for (int k = c + 1; k < D; ++k) s -= Sb[i * SD + k] * Lm[k * SD + c];
            Sb[i * SD + c] = s / Lm[c * SD + c];
        }
    }
    CTA_SYNC();
    if (tid < nval) {
        const int g = tid / Ng;
        const float *gs_ = scr + TCMM_GST + g * TCMM_GSTRIDE;
        const float *Sb = gs_ + 3 * SD + 3 * SD * SD, *dm = gs_ + 3 * SD + 4 * SD * SD;
        float out[SD];
        for (int d = 0; d < D; ++d) {
            float acc = 0.f;
            for (int j = 0; j < D; ++j)
                acc = fmaf(0.5f * (Sb[d * SD + j] + Sb[j * SD + d]), xs[tid * TC_SDP + j] - gs_[j], acc);
            out[d] = dm[d] / (float)Ng + (2.f / (float)(Ng - 1)) * acc;
        }
        for (int d = 0; d < D; ++d) gs[tid * TC_SDP + d] = out[d];
    }
    CTA_SYNC();
}

}  // namespace pmb
