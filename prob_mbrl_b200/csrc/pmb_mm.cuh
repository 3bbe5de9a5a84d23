// Moment matching inside the sweeps (reference utils/rollout.py:20-29,121-145):
//   x' = m + zhat chol(S)^T,  m = mean_n x,  S = (x-m)^T (x-m)/(M-1) + 1e-12 I,
//   zhat = (z - mean z) / std_unbiased(z) per column, z = z_mm[(t + n) mod N] (rollout.py:53-59).
// The particles of one moment-matching group live in several CTAs, so every step needs ONE
// cross-CTA reduction: each CTA publishes its block statistics (count, block mean, centred scatter;
// fp64) to global memory, a grid barrier (cooperative launch) follows, and every CTA combines the
// records of its group in a fixed order (Chan et al. pairwise update) -- deterministic and identical
// on every CTA.  The 1e-12-jittered Cholesky runs in fp32 like the reference's and reports a
// non-positive pivot through the status word (reference: cholesky() raises, rollout.py:25,154-157).
// The reverse step is the hand-derived adjoint checked against autograd in
// tests/test_oracle_backward.py (oracle/rollout_oracle.py::mm_backward).
#pragma once
#include "pmb_internal.cuh"

namespace pmb {

constexpr int MMREC = 288;   // doubles per CTA record (1 + D + D*D + 2D <= 271 for D <= 15)

// offsets (floats) inside the mm scratch block of shared memory
constexpr int MM_XS = 0;                 // [P][SD] pre-mm particles (fwd)
// after the two P-scaled tiles: fixed part
struct MMSmem {
    float *xs, *zs;      // [P][SD]
    double *dbl;         // [SD] + [SD] local means (x, z); [SD] + [SD] group means
    float *st;           // m[SD], zm[SD], zistd[SD]
    float *L;            // [SD*SD]
    float *A, *X, *Sb;   // [SD*SD] each (reverse step)
    float *dm;           // [SD]
    double *stage;       // [MM_STAGE] the group's per-CTA records, fetched with ONE round of parallel loads per step
    template <int P>
    __device__ __forceinline__ void carve(float *base) {
        xs = base;
        zs = xs + P * SD;
        dbl = reinterpret_cast<double *>(zs + P * SD);      // 4*SD doubles = 8*SD floats
        st = reinterpret_cast<float *>(dbl + 4 * SD);
        L = st + 3 * SD;
        A = L + SD * SD;
        X = A + SD * SD;
        Sb = X + SD * SD;
        dm = Sb + SD * SD;
        stage = reinterpret_cast<double *>(dm + SD + (((2 * P * SD) & 1) ? 1 : 0));     // 8-byte aligned (offsets are even)
    }
};
// The combine after the grid barrier used to walk the group's records with dependent L2 loads (one round trip per
// CTA and quantity: ~17 k cycles per step at 25 CTAs); the records are now staged in shared memory first.
constexpr int MM_STAGE = 1280;   // doubles: e.g. 31 CTAs x 41 doubles (D = 5); larger groups fall back to direct loads
constexpr int mm_smem_floats(int P) { return 2 * P * SD + 8 * SD + 3 * SD + 4 * SD * SD + SD + 34 + 2 * MM_STAGE; }

// All CTAs of the (cooperatively launched) grid meet here.  `epoch` counts arrivals expected so far.
__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned &epoch) {
    CTA_SYNC();
    epoch += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned v, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (++spins > (1u << 28)) __trap();
        } while (v < epoch);
        __threadfence();
    }
    __syncwarp();       // the polling lane rejoins its warp before the (aligned) barrier
    CTA_SYNC();
}

struct MMGroup {
    int gid, c0, c1, Ng, Pv;   // group id, CTA range of the group, particles per group, valid particles here
    template <int P>
    __device__ __forceinline__ void set(const SweepParams &prm, int n0) {
        Ng = prm.mm_Ng;
        gid = n0 / Ng;
        if (prm.mm_G <= 1) {
            c0 = 0;
            c1 = gridDim.x;
        } else {
            const int cpg = Ng / P;     // planner guarantees Ng % P == 0 when G > 1
            c0 = gid * cpg;
            c1 = c0 + cpg;
        }
        Pv = min(P, prm.N - n0);
    }
};

// ---- forward: replace this CTA's pre-mm particles (role-B registers) by the moment-matched ones ----
template <int P>
__device__ __forceinline__ void mm_states_forward(const SweepParams &prm, const MMSmem &M, const MMGroup &grp,
                                                  int t, int n0, bool roleB, int b_p, int b_d, int b_n,
                                                  float &s_reg, float zrow, unsigned &epoch) {
    const int D = prm.D, N = prm.N, tid = threadIdx.x;
    double *lm = M.dbl, *lzm = M.dbl + SD, *gm = M.dbl + 2 * SD, *gzm = M.dbl + 3 * SD;
    if (roleB) {
        M.xs[b_p * SD + b_d] = s_reg;
        M.zs[b_p * SD + b_d] = zrow;
        if (n0 + b_p < N) prm.s1pre[((size_t)t * N + b_n) * D + b_d] = s_reg;
    }
    CTA_SYNC();
    if (tid < D) {
        double a = 0.0, b = 0.0;
        for (int p = 0; p < grp.Pv; ++p) {
            a += (double)M.xs[p * SD + tid];
            b += (double)M.zs[p * SD + tid];
        }
        lm[tid] = a / grp.Pv;
        lzm[tid] = b / grp.Pv;
    }
    CTA_SYNC();
    double *rec = prm.mmrec + ((size_t)(t & 1) * gridDim.x + blockIdx.x) * MMREC;
    for (int idx = tid; idx < D * D; idx += NT) {
        const int i = idx / D, j = idx - i * D;
        if (j <= i) {
            double a = 0.0;
            for (int p = 0; p < grp.Pv; ++p)
                a += ((double)M.xs[p * SD + i] - lm[i]) * ((double)M.xs[p * SD + j] - lm[j]);
            rec[1 + D + idx] = a;
        }
    }
    if (tid < D) {
        double a = 0.0;
        for (int p = 0; p < grp.Pv; ++p) {
            const double d = (double)M.zs[p * SD + tid] - lzm[tid];
            a += d * d;
        }
        rec[1 + tid] = lm[tid];
        rec[1 + D + D * D + tid] = lzm[tid];
        rec[1 + D + D * D + D + tid] = a;
    }
    if (tid == 0) rec[0] = (double)grp.Pv;
    grid_barrier(prm.mmctr, epoch);
    // the records of this group: staged in shared memory when they fit (one round of parallel loads), else read in place
    const int ncta = grp.c1 - grp.c0, nrec = 1 + 3 * D + D * D;
    const bool staged = ncta * nrec <= MM_STAGE;
    const double *recs = prm.mmrec + ((size_t)(t & 1) * gridDim.x + grp.c0) * MMREC;
    int rstride = MMREC;
    if (staged) {
        for (int i = tid; i < ncta * nrec; i += NT) {
            const int c = i / nrec, k = i - c * nrec;
            M.stage[i] = __ldcg(recs + (size_t)c * MMREC + k);
        }
        CTA_SYNC();
        recs = M.stage;
        rstride = nrec;
    }
    auto rd = [&](const double *q) { return staged ? *q : __ldcg(q); };
    if (tid < D) {
        double n = 0.0, a = 0.0, b = 0.0;
        for (int c = 0; c < ncta; ++c) {
            const double *r = recs + (size_t)c * rstride;
            const double nc = rd(r);
            n += nc;
            a += nc * rd(r + 1 + tid);
            b += nc * rd(r + 1 + D + D * D + tid);
        }
        gm[tid] = a / n;
        gzm[tid] = b / n;
    }
    CTA_SYNC();
    for (int idx = tid; idx < D * D; idx += NT) {
        const int i = idx / D, j = idx - i * D;
        if (j <= i) {
            double a = 0.0;
            for (int c = 0; c < ncta; ++c) {
                const double *r = recs + (size_t)c * rstride;
                const double nc = rd(r);
                a += rd(r + 1 + D + idx) + nc * (rd(r + 1 + i) - gm[i]) * (rd(r + 1 + j) - gm[j]);
            }
            // unbiased covariance + jitter (rollout.py:24), handed to the fp32 Cholesky
            M.A[i * SD + j] = (float)(a / (double)(grp.Ng - 1)) + (i == j ? 1e-12f : 0.f);
        }
    }
    if (tid < D) {
        double a = 0.0;
        for (int c = 0; c < ncta; ++c) {
            const double *r = recs + (size_t)c * rstride;
            const double nc = rd(r);
            const double dz = rd(r + 1 + D + D * D + tid) - gzm[tid];
            a += rd(r + 1 + D + D * D + D + tid) + nc * dz * dz;
        }
        M.st[tid] = (float)gm[tid];
        M.st[SD + tid] = (float)gzm[tid];
        M.st[2 * SD + tid] = 1.f / sqrtf((float)(a / (double)(grp.Ng - 1)));   // 1 / z.std(unbiased)
    }
    CTA_SYNC();
    if (tid == 0) {
        bool ok = true;
        for (int i = 0; i < D; ++i) {
            for (int j = 0; j <= i; ++j) {
                float s = M.A[i * SD + j];
                for (int k = 0; k < j; ++k) s -= M.L[i * SD + k] * M.L[j * SD + k];
                if (i == j) {
                    if (!(s > 0.f)) { ok = false; s = 1.f; }
                    M.L[i * SD + i] = sqrtf(s);
                } else {
                    M.L[i * SD + j] = s / M.L[j * SD + j];
                }
            }
            for (int j = i + 1; j < D; ++j) M.L[i * SD + j] = 0.f;
        }
        if (!ok && prm.status) atomicCAS(prm.status, 0, 1 + t);
    }
    CTA_SYNC();
    // keep (m, L, z statistics) of this step for the reverse sweep
    if (blockIdx.x == grp.c0) {
        float *ms = prm.mmstat + ((size_t)t * max(prm.mm_G, 1) + grp.gid) * (3 * SD + SD * SD);
        for (int i = tid; i < 3 * SD; i += NT) ms[i] = M.st[i];
        for (int i = tid; i < SD * SD; i += NT) ms[3 * SD + i] = M.L[i];
    }
    if (roleB) {
        float x = M.st[b_d];
        for (int j = 0; j <= b_d; ++j)
            x = fmaf((M.zs[b_p * SD + j] - M.st[SD + j]) * M.st[2 * SD + j], M.L[b_d * SD + j], x);
        s_reg = x;
    }
}

// ---- reverse: turn the cotangent of the moment-matched particles (gs, in place) into the cotangent of
//      the pre-mm particles.  xs must hold the pre-mm particles of step t, zs their z rows. ----
template <int P>
__device__ __forceinline__ void mm_states_backward(const SweepParams &prm, const MMSmem &M, const MMGroup &grp,
                                                   int t, float *gs, bool roleB, int b_p, int b_d,
                                                   unsigned &epoch) {
    const int D = prm.D, tid = threadIdx.x;
    const float *ms = prm.mmstat + ((size_t)t * max(prm.mm_G, 1) + grp.gid) * (3 * SD + SD * SD);
    for (int i = tid; i < 3 * SD; i += NT) M.st[i] = __ldcg(ms + i);
    for (int i = tid; i < SD * SD; i += NT) M.L[i] = __ldcg(ms + 3 * SD + i);
    CTA_SYNC();
    // block partials: dm = sum_n g_n,  dL = tril(sum_n g_n zhat_n^T)
    double *rec = prm.mmrec + ((size_t)(t & 1) * gridDim.x + blockIdx.x) * MMREC;
    for (int idx = tid; idx < D * D + D; idx += NT) {
        double a = 0.0;
        if (idx < D * D) {
            const int i = idx / D, j = idx - i * D;
            if (j <= i)
                for (int p = 0; p < grp.Pv; ++p)
                    a += (double)gs[p * SD + i] * (double)((M.zs[p * SD + j] - M.st[SD + j]) * M.st[2 * SD + j]);
        } else {
            const int d = idx - D * D;
            for (int p = 0; p < grp.Pv; ++p) a += (double)gs[p * SD + d];
        }
        rec[idx] = a;
    }
    grid_barrier(prm.mmctr, epoch);
    const int ncta = grp.c1 - grp.c0, nrec = D * D + D;
    const bool staged = ncta * nrec <= MM_STAGE;
    const double *recs = prm.mmrec + ((size_t)(t & 1) * gridDim.x + grp.c0) * MMREC;
    int rstride = MMREC;
    if (staged) {
        for (int i = tid; i < ncta * nrec; i += NT) {
            const int c = i / nrec, k = i - c * nrec;
            M.stage[i] = __ldcg(recs + (size_t)c * MMREC + k);
        }
        CTA_SYNC();
        recs = M.stage;
        rstride = nrec;
    }
    for (int idx = tid; idx < D * D + D; idx += NT) {
        double a = 0.0;
        for (int c = 0; c < ncta; ++c) a += staged ? recs[(size_t)c * rstride + idx] : __ldcg(recs + (size_t)c * rstride + idx);
        if (idx < D * D) M.X[(idx / D) * SD + (idx % D)] = (float)a;     // dL (lower triangle), staged in X
        else M.dm[idx - D * D] = (float)a;
    }
    CTA_SYNC();
    // A = Phi(L^T dL): lower triangle, diagonal halved
    for (int idx = tid; idx < D * D; idx += NT) {
        const int i = idx / D, j = idx - i * D;
        float a = 0.f;
        if (j <= i) {
            for (int k = i; k < D; ++k) a = fmaf(M.L[k * SD + i], M.X[k * SD + j], a);   // dL[k][j] = 0 for j > k
            if (i == j) a *= 0.5f;
        }
        M.A[i * SD + j] = a;
    }
    CTA_SYNC();
    // X = L^-T A  (back substitution, one column per thread)
    if (tid < D) {
        const int j = tid;
        for (int r = D - 1; r >= 0; --r) {
            float s = M.A[r * SD + j];
            for (int k = r + 1; k < D; ++k) s -= M.L[k * SD + r] * M.X[k * SD + j];
            M.X[r * SD + j] = s / M.L[r * SD + r];
        }
    }
    CTA_SYNC();
    // Sb = X L^-1  (one row per thread)
    if (tid < D) {
        const int i = tid;
        for (int c = D - 1; c >= 0; --c) {
            float s = M.X[i * SD + c];
            for (int k = c + 1; k < D; ++k) s -= M.Sb[i * SD + k] * M.L[k * SD + c];
            M.Sb[i * SD + c] = s / M.L[c * SD + c];
        }
    }
    CTA_SYNC();
    // dx_n = dm/M + 2/(M-1) * sym(Sb) (x_n - m)
    if (roleB) {
        float acc = 0.f;
        for (int j = 0; j < D; ++j)
            acc = fmaf(0.5f * (M.Sb[b_d * SD + j] + M.Sb[j * SD + b_d]), M.xs[b_p * SD + j] - M.st[j], acc);
        gs[b_p * SD + b_d] = M.dm[b_d] / (float)grp.Ng + (2.f / (float)(grp.Ng - 1)) * acc;
    }
    CTA_SYNC();
}

}  // namespace pmb
