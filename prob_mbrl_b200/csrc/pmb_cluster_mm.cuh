// Moment matching of the states inside the cluster-resident sweeps (reference utils/rollout.py:20-29,121-145):
//   x' = m + zhat chol(S)^T,  m = mean_n x,  S = (x-m)^T (x-m)/(N-1) + 1e-12 I,
//   zhat = (z - mean z) / std_unbiased(z) per column, z = z_mm[(t + n) mod N] (rollout.py:53-59).
// The particles live in <= 15 clusters, so every step needs ONE cross-cluster exchange: the owner CTA of a particle
// writes it to global memory, every particle tile (warp group) of every CTA arrives on a global counter, and after
// the barrier every tile reads ALL particles back (N x D floats out of L2) and forms the statistics itself, in a
// fixed order -- deterministic and identical everywhere, no per-block records to combine.  One matching group only
// (the whole particle set), N <= 128: with a single group the z rows of a step are a rotation of the same set, so
// their mean / std are constants of the launch.  The reverse step is the hand-derived adjoint of
// oracle/rollout_oracle.py::mm_backward (same formulas as pmb_mm.cuh).
#pragma once
#include "pmb_cluster.cuh"

namespace pmb {

constexpr int CMM_NMAX = 128;                 // particles (<= 8 x the co-resident clusters)
constexpr int CMM_PART = 256;                 // doubles: chunked partial sums
// shared memory of ONE particle tile (floats): xs, zs [128][SD]; partial sums; m, zm, zistd; L, A, X, Sb; dm; own rows
constexpr int CMM_FLOATS = 2 * CMM_NMAX * SD + 2 * CMM_PART + 2 * 2 * SD + 3 * SD + 4 * SD * SD + SD + 2 * CL_TS * SD;

struct CMM {
    float *xs, *zs;          // [N][SD] all particles (forward: pre-matching states; reverse: adjoints) / standardised z rows
    double *part;            // [CMM_PART]
    double *dmean;           // [SD] means in double, [SD] scratch
    float *st;               // m[SD], zm[SD], zistd[SD]
    float *L, *A, *X, *Sb;   // [SD*SD]
    float *dm;               // [SD]
    float *zrow;             // [4][SD] z rows of the tile's slots (forward) / pre-matching states of the slots (reverse)
    float *aux;              // [4][SD]
    __device__ __forceinline__ void carve(float *base) {
        xs = base;
        zs = xs + CMM_NMAX * SD;
        part = reinterpret_cast<double *>(zs + CMM_NMAX * SD);
        dmean = part + CMM_PART;
        st = reinterpret_cast<float *>(dmean + 2 * SD);
        L = st + 3 * SD;
        A = L + SD * SD;
        X = A + SD * SD;
        Sb = X + SD * SD;
        dm = Sb + SD * SD;
        zrow = dm + SD;
        aux = zrow + CL_TS * SD;
    }
};

// every active particle tile of the grid meets here; `target` = arrivals expected so far
__device__ __forceinline__ void cmm_barrier(unsigned *ctr, unsigned target, int g, int gtid) {
    __threadfence();
    CT_SYNC(g);
    if (gtid == 0) {
        atomicAdd(ctr, 1u);
        unsigned v, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (++spins > (1u << 26)) __trap();
        } while (v < target);
        __threadfence();
    }
    CT_SYNC(g);
}

// One reduced quantity: sum_n (P[n][i] - cp[i]) * (Q[n][j] - cq[j]);  j < 0: sum_n (P[n][i] - cp[i])
struct CmmQ {
    int i, j;
};
// lower-triangle index q = i (i + 1) / 2 + j  ->  (i, j)
__device__ __forceinline__ CmmQ cmm_tri(int q) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= q) ++i;
    return CmmQ{i, q - i * (i + 1) / 2};
}
// out[q] for q < nq, in double: chunked over the tile's 128 threads, combined in a fixed order (identical on every tile)
template <typename QF>
__device__ __forceinline__ void cmm_reduce(int nq, int N, QF qf, const float *P, const double *cp, const float *Q, const double *cq,
                                           double *part, double *out, int g, int gtid) {
    const int ch = max(1, min(8, CMM_PART / max(nq, 1)));
    for (int idx = gtid; idx < nq * ch; idx += CL_GT) {
        const int q = idx / ch, c = idx - q * ch;
        const CmmQ s = qf(q);
        const double ci = cp ? cp[s.i] : 0.0, cj = (cq && s.j >= 0) ? cq[s.j] : 0.0;
        double a = 0.0;
        if (s.j < 0) {
            for (int n = c; n < N; n += ch) a += (double)P[n * SD + s.i] - ci;
        } else {
            for (int n = c; n < N; n += ch) a += ((double)P[n * SD + s.i] - ci) * ((double)Q[n * SD + s.j] - cj);
        }
        part[idx] = a;
    }
    CT_SYNC(g);
    for (int q = gtid; q < nq; q += CL_GT) {
        double a = 0.0;
        for (int c = 0; c < ch; ++c) a += part[q * ch + c];
        out[q] = a;
    }
    CT_SYNC(g);
}

// active particle tiles of the launch (arrivals per step)
__device__ __forceinline__ unsigned cmm_active_tiles(int N, int PG, int C) {
    unsigned n = 0;
    for (int n0 = 0; n0 < N; n0 += PG) {
        const int nval = min(PG, N - n0);
        n += (unsigned)C * (1u + ((nval - ((nval + 1) >> 1)) > 0 ? 1u : 0u));
    }
    return n;
}

// constants of the launch: mean and 1 / unbiased std of the z_mm rows (all N rows take part in every step)
__device__ __forceinline__ void cmm_z_statistics(const ClusterParams &prm, const CMM &M, int g, int gtid) {
    const int D = prm.D, N = prm.N;
    for (int i = gtid; i < N * D; i += CL_GT) M.zs[(i / D) * SD + (i % D)] = __ldg(prm.z_mm + i);
    CT_SYNC(g);
    double *mean = M.dmean, *sq = M.dmean + SD;
    cmm_reduce(D, N, [](int q) { return CmmQ{q, -1}; }, M.zs, nullptr, nullptr, nullptr, M.part, mean, g, gtid);
    if (gtid < D) mean[gtid] = mean[gtid] / N;
    CT_SYNC(g);
    cmm_reduce(D, N, [](int q) { return CmmQ{q, q}; }, M.zs, mean, M.zs, mean, M.part, sq, g, gtid);
    if (gtid < D) {
        M.st[SD + gtid] = (float)mean[gtid];
        M.st[2 * SD + gtid] = 1.f / sqrtf((float)(sq[gtid] / (double)(N - 1)));
    }
    CT_SYNC(g);
}

// ---- forward: the tile's pre-matching states (role-B registers) -> the matched ones.  Called by all 128 threads of
//      the tile; M.zrow[slot][j] holds z_mm[(t + n) mod N][j] of the tile's slots. ----
__device__ __forceinline__ void cmm_forward(const ClusterParams &prm, const CMM &M, int g, int gtid, int t, unsigned target,
                                            bool roleB, int b_p, int b_d, int b_n, bool b_own, float &s_reg, bool leader) {
    const int D = prm.D, N = prm.N;
    float *s1pre = const_cast<float *>(prm.s1pre);
    if (roleB && b_own) s1pre[((size_t)t * N + b_n) * D + b_d] = s_reg;
    cmm_barrier(prm.mmctr, target, g, gtid);
    for (int i = gtid; i < N * D; i += CL_GT) {
        const int n = i / D, d = i - n * D;
        M.xs[n * SD + d] = __ldcg(s1pre + (size_t)t * N * D + i);
    }
    CT_SYNC(g);
    double *sum = M.dmean;
    cmm_reduce(D, N, [](int q) { return CmmQ{q, -1}; }, M.xs, nullptr, nullptr, nullptr, M.part, sum, g, gtid);
    if (gtid < D) sum[gtid] = sum[gtid] / N;
    CT_SYNC(g);
    // lower triangle of the centred scatter, quantity q = i (i + 1) / 2 + j
    const int nq = D * (D + 1) / 2;
    double *cov = reinterpret_cast<double *>(M.X);       // SD*SD floats = 128 doubles >= 120
    cmm_reduce(nq, N, [](int q) { return cmm_tri(q); }, M.xs, sum, M.xs, sum, M.part, cov, g, gtid);
    for (int q = gtid; q < nq; q += CL_GT) {
        const CmmQ s = cmm_tri(q);
        // unbiased covariance + jitter (rollout.py:24), handed to the fp32 Cholesky
        M.A[s.i * SD + s.j] = (float)(cov[q] / (double)(N - 1)) + (s.i == s.j ? 1e-12f : 0.f);
    }
    if (gtid < D) M.st[gtid] = (float)sum[gtid];
    CT_SYNC(g);
    if (gtid == 0) {
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
    CT_SYNC(g);
    if (leader) {       // (m, L, z statistics) of this step for the reverse sweep
        float *ms = prm.mmstat + (size_t)t * (3 * SD + SD * SD);
        for (int i = gtid; i < 3 * SD; i += CL_GT) ms[i] = M.st[i];
        for (int i = gtid; i < SD * SD; i += CL_GT) ms[3 * SD + i] = M.L[i];
    }
    if (roleB) {
        float x = M.st[b_d];
        for (int j = 0; j <= b_d; ++j)
            x = fmaf((M.zrow[b_p * SD + j] - M.st[SD + j]) * M.st[2 * SD + j], M.L[b_d * SD + j], x);
        s_reg = x;
    }
}

// ---- reverse: the cotangent of the matched particles (gs[slot][d], in place) -> the cotangent of the pre-matching
//      ones.  cmm_backward_prefetch (before the barrier, constant inputs): standardised z rows of ALL particles,
//      (m, L) of step t, the pre-matching states of the tile's slots. ----
__device__ __forceinline__ void cmm_backward_prefetch(const ClusterParams &prm, const CMM &M, int g, int gtid, int t, bool roleB,
                                                      int b_p, int b_d, int b_n) {
    const int D = prm.D, N = prm.N;
    const float *ms = prm.mmstat + (size_t)t * (3 * SD + SD * SD);
    for (int i = gtid; i < D; i += CL_GT) M.st[i] = __ldcg(ms + i);
    for (int i = gtid; i < SD * SD; i += CL_GT) M.L[i] = __ldcg(ms + 3 * SD + i);
    for (int i = gtid; i < N * D; i += CL_GT) {
        const int n = i / D, j = i - n * D;
        int r = t + n;
        r -= (r / N) * N;
        M.zs[n * SD + j] = (__ldg(prm.z_mm + (size_t)r * D + j) - M.st[SD + j]) * M.st[2 * SD + j];
    }
    if (roleB) M.zrow[b_p * SD + b_d] = __ldcg(prm.s1pre + ((size_t)t * N + b_n) * D + b_d);
}
__device__ __forceinline__ void cmm_backward(const ClusterParams &prm, const CMM &M, int g, int gtid, int t, unsigned target,
                                             float *gs, bool roleB, int b_p, int b_d, int b_n, bool b_own) {
    const int D = prm.D, N = prm.N;
    float *gbuf = prm.gbuf + (size_t)(t & 1) * N * SD;
    if (roleB && b_own) gbuf[(size_t)b_n * SD + b_d] = gs[b_p * SD + b_d];
    cmm_barrier(prm.mmctr, target, g, gtid);
    for (int i = gtid; i < N * D; i += CL_GT) {
        const int n = i / D, d = i - n * D;
        M.xs[n * SD + d] = __ldcg(gbuf + (size_t)n * SD + d);
    }
    CT_SYNC(g);
    // dm = sum_n g_n (q < D);  dL = tril(sum_n g_n zhat_n^T) (q = D + i (i + 1) / 2 + j)
    const int nq = D + D * (D + 1) / 2;
    double *acc = reinterpret_cast<double *>(M.A);       // A and X are contiguous: 2 * SD*SD floats = 256 doubles >= 135
    cmm_reduce(nq, N, [D](int q) {
        if (q < D) return CmmQ{q, -1};
        return cmm_tri(q - D);
    }, M.xs, nullptr, M.zs, nullptr, M.part, acc, g, gtid);
    // stage dL in Sb (A / X are overwritten next), dm in dm
    for (int q = gtid; q < nq; q += CL_GT) {
        if (q < D) {
            M.dm[q] = (float)acc[q];
        } else {
            const CmmQ s = cmm_tri(q - D);
            M.Sb[s.i * SD + s.j] = (float)acc[q];
        }
    }
    CT_SYNC(g);
    // A = Phi(L^T dL): lower triangle, diagonal halved
    for (int idx = gtid; idx < D * D; idx += CL_GT) {
        const int i = idx / D, j = idx - i * D;
        float a = 0.f;
        if (j <= i) {
            for (int k = i; k < D; ++k) a = fmaf(M.L[k * SD + i], M.Sb[k * SD + j], a);   // dL[k][j] = 0 for j > k
            if (i == j) a *= 0.5f;
        }
        M.A[i * SD + j] = a;
    }
    CT_SYNC(g);
    // X = L^-T A  (back substitution, one column per thread)
    if (gtid < D) {
        const int j = gtid;
        for (int r = D - 1; r >= 0; --r) {
            float s = M.A[r * SD + j];
            for (int k = r + 1; k < D; ++k) s -= M.L[k * SD + r] * M.X[k * SD + j];
            M.X[r * SD + j] = s / M.L[r * SD + r];
        }
    }
    CT_SYNC(g);
    // Sb = X L^-1  (one row per thread)
    if (gtid < D) {
        const int i = gtid;
        for (int c = D - 1; c >= 0; --c) {
            float s = M.X[i * SD + c];
            for (int k = c + 1; k < D; ++k) s -= M.Sb[i * SD + k] * M.L[k * SD + c];
            M.Sb[i * SD + c] = s / M.L[c * SD + c];
        }
    }
    CT_SYNC(g);
    // dx_n = dm/N + 2/(N-1) * sym(Sb) (x_n - m)
    if (roleB) {
        float a = 0.f;
        for (int j = 0; j < D; ++j)
            a = fmaf(0.5f * (M.Sb[b_d * SD + j] + M.Sb[j * SD + b_d]), M.zrow[b_p * SD + j] - M.st[j], a);
        gs[b_p * SD + b_d] = M.dm[b_d] / (float)N + (2.f / (float)(N - 1)) * a;
    }
    CT_SYNC(g);
}

}  // namespace pmb
