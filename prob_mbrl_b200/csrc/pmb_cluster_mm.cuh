// Moment matching of the states inside the cluster-resident sweeps (reference utils/rollout.py:20-29,121-145):
//   x' = m + zhat chol(S)^T,  m = mean_n x,  S = (x-m)^T (x-m)/(N-1) + 1e-12 I,
//   zhat = (z - mean z) / std_unbiased(z) per column, z = z_mm[(t + n) mod N] (rollout.py:53-59).
// The particles live in <= 15 clusters (per GPU), so every step needs ONE cross-cluster exchange: one CTA per cluster
// (rank 0 of the cluster) publishes the raw moments of each of its two particle tiles (sum x, sum x x^T in double:
// products of two floats are exact) as tagged 8-byte words, and every tile of every CTA polls the <= 30 records per
// GPU until they carry the step's tag and adds them in tile order -- deterministic and identical everywhere, no fence,
// counter or barrier (cmm_publish / cmm_fetch).  Across the GPUs of a node the same stores go to every rank's record
// area over NVLink.  One matching group only (the whole particle set): with a single group the z rows of a step are a
// rotation of the same set, so their mean / std are constants of the launch.  The reverse step is the hand-derived
// adjoint of oracle/rollout_oracle.py::mm_backward (same formulas as pmb_mm.cuh).
#pragma once
#include "pmb_cluster.cuh"

namespace pmb {

constexpr int CMM_NMAX = 128;                 // particles (<= 8 x the co-resident clusters)
constexpr int CMM_NQ = SD + SD * (SD + 1) / 2;  // reduced quantities (max): D sums + D (D + 1) / 2 products
constexpr int CMM_TILES = 2 * 32;             // particle tiles of the grid (max): two per cluster
// shared memory of ONE particle tile (floats): staged records (doubles); totals, means (doubles); (i, j) table;
// m, zm, zistd; L, A, X, Sb; dm; rows of the tile's slots (two sets)
constexpr int CMM_FLOATS = 2 * CMM_TILES * 16 + 2 * CMM_NQ + 2 * SD + CMM_NQ + 3 * SD + 4 * SD * SD + SD + 2 * CL_TS * SD + 24;

struct CMM {
    double *stage;           // [tiles][nq] records of the step (when they fit: tiles * nq <= CMM_TILES * 16), else read in place
    double *red;             // [NQ] totals
    double *dmean;           // [SD]
    int *qtab;               // [NQ] (i << 8 | j) of the lower-triangle quantity q
    float *st;               // m[SD], zm[SD], zistd[SD]
    float *L, *A, *X, *Sb;   // [SD*SD]
    float *dm;               // [SD]
    float *zrow;             // [4][SD] z rows of the tile's slots (forward: raw; reverse: standardised)
    float *xrow;             // [4][SD] pre-matching states of the tile's slots
    float *rinv;             // [SD] reciprocal pivots of L
    __device__ __forceinline__ void carve(float *base) {
        stage = reinterpret_cast<double *>(base);
        red = stage + CMM_TILES * 16;
        dmean = red + CMM_NQ;
        qtab = reinterpret_cast<int *>(dmean + SD);
        st = reinterpret_cast<float *>(qtab + CMM_NQ);
        L = st + 3 * SD;
        A = L + SD * SD;
        X = A + SD * SD;
        Sb = X + SD * SD;
        dm = Sb + SD * SD;
        zrow = dm + SD;
        xrow = zrow + CL_TS * SD;
        rinv = xrow + CL_TS * SD;
    }
};

// lower-triangle index q = i (i + 1) / 2 + j  ->  (i, j)
__device__ __forceinline__ void cmm_init_qtab(const CMM &M, int D, int gtid) {
    for (int q = gtid; q < D * (D + 1) / 2; q += CL_GT) {
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= q) ++i;
        M.qtab[q] = (i << 8) | (q - i * (i + 1) / 2);
    }
}
__device__ __forceinline__ unsigned cmm_clusters(int N, int PG) { return (unsigned)((N + PG - 1) / PG); }
// Exchange protocol ("LL": flag in the data): a record entry is a double cut into two 8-byte words, each carrying 32
// data bits and the 32-bit tag of the step -- an aligned 8-byte store is single-copy atomic, so a reader that sees the
// tag in both words has the value, with no fence, counter or barrier anywhere: the writer (rank 0 of a cluster, or of
// any cluster of any GPU of the node) never waits, a reader only waits for writers, and the two particle tiles of a CTA
// no longer meet.  Tags grow by one per step; entries are double-buffered by step parity because a writer can be at most
// one step ahead of a reader (it needs the reader's next record to go further).
__device__ __forceinline__ unsigned cmm_arrivals(const ClusterParams &) { return 1u; }     // tag increment per step
// tag before the first step of this launch: 0 on one GPU (the record area is zeroed before every launch), the count the
// previous launch left across GPUs (peer-mapped areas are never reset: a peer may already be a launch ahead)
__device__ __forceinline__ unsigned cmm_base(const ClusterParams &prm) {
    return prm.mm_base ? (unsigned)__ldcg(prm.mm_base) : 0u;
}
// record entry q of this tile (2 cluster + g) for step parity `par`: to the own record area, or to every rank's
__device__ __forceinline__ void cmm_publish(const ClusterParams &prm, int par, int g, int nq, int q, double a, unsigned tag) {
    const int tl = 2 * (int)cmm_clusters(prm.N, prm.PG);           // tiles of one rank
    const int world = max(prm.mm_world, 1);
    const size_t idx = ((size_t)par * world * tl + (size_t)prm.mm_rank * tl + 2 * (blockIdx.x / prm.C) + g) * nq + q;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(a), hi = (unsigned long long)tag << 32;
    const unsigned long long w0 = (bits & 0xffffffffull) | hi, w1 = (bits >> 32) | hi;
    for (int p = 0; p < world; ++p) {
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(world > 1 ? prm.mmrec_peer[p] : prm.mmrec) + idx;
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
    }
}
// entries i, i + 128, ... (< n, up to 4) of the record area once both words of each carry `tag`: all loads of a round in
// flight at once (one L2 / NVLink round trip per round, not per entry)
__device__ __forceinline__ void cmm_fetch4(const ulonglong2 *rec, int i, int n, unsigned tag, double *stage) {
    unsigned long long w0[4], w1[4], spins = 0;
    bool need[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) need[k] = i + k * CL_GT < n;
    for (;;) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (need[k]) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[k]), "=l"(w1[k]) : "l"(rec + i + k * CL_GT) : "memory");
        bool any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (need[k]) {
                if ((unsigned)(w0[k] >> 32) == tag && (unsigned)(w1[k] >> 32) == tag) {
                    stage[i + k * CL_GT] = __longlong_as_double((long long)((w1[k] << 32) | (w0[k] & 0xffffffffull)));
                    need[k] = false;
                } else {
                    any = true;
                }
            }
        }
        if (!any) break;
        if (++spins > (1ull << 27)) __trap();          // a writer that never shows up must not hang the GPU forever
    }
}

// totals of the step: M.red[q] = sum over the particle tiles of all ranks of their records, in tile order (identical on
// every tile).  Records: entry (2 cluster + g) * nq + q of a rank's block, written by rank 0 of the cluster.  Staged in
// passes of whole tiles (<= CMM_TILES * 16 entries); 16 quantities at a time: thread = (quantity, chunk of tiles c,
// c + 8, ...), then a 3-level butterfly over the 8 chunks -- a fixed tree.  Kept small on purpose: single-warp code runs
// at ~10 cycles per instruction here.
__device__ __forceinline__ void cmm_combine(const CMM &M, const double *recd, int ntiles, int nq, unsigned tag, int g, int gtid) {
    const ulonglong2 *rec = reinterpret_cast<const ulonglong2 *>(recd);
    const int tpp = (CMM_TILES * 16) / nq;                 // tiles per pass
#pragma unroll 1
    for (int t0 = 0; t0 < ntiles; t0 += tpp) {
        const int nt = min(tpp, ntiles - t0), n = nt * nq;
        for (int i = gtid; i < n; i += 4 * CL_GT) cmm_fetch4(rec + t0 * nq, i, n, tag, M.stage);
        CT_SYNC(g);
#pragma unroll 1
        for (int q0 = 0; q0 < nq; q0 += 16) {
            const int q = q0 + (gtid >> 3), c = gtid & 7;
            double a = 0.0;
            if (q < nq) {
#pragma unroll 1
                for (int tl = c; tl < nt; tl += 8) a += M.stage[tl * nq + q];
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            if (q < nq && c == 0) M.red[q] = t0 ? M.red[q] + a : a;
        }
        CT_SYNC(g);
    }
}

// an idle tile (no particles) still takes part in the per-step exchange: an all-zero record -- and it waits for the
// step's records like everybody else, which is what keeps it from running more than one step (= one parity slot) ahead
__device__ __forceinline__ void cmm_idle_step(const ClusterParams &prm, const CMM &M, int g, int gtid, int rank, int t, unsigned tag) {
    const int nq = prm.D + prm.D * (prm.D + 1) / 2;
    const int ntiles = 2 * (int)cmm_clusters(prm.N, prm.PG) * max(prm.mm_world, 1);
    if (rank == 0)
        for (int q = gtid; q < nq; q += CL_GT) cmm_publish(prm, t & 1, g, nq, q, 0.0, tag);
    cmm_combine(M, prm.mmrec + (size_t)(t & 1) * ntiles * nq * 2, ntiles, nq, tag, g, gtid);
}

// constants of the launch: mean and 1 / unbiased std of the z_mm rows (all N rows take part in every step)
__device__ __forceinline__ void cmm_z_statistics(const ClusterParams &prm, const CMM &M, int g, int gtid) {
    const int D = prm.D, N = prm.n_global;
    cmm_init_qtab(M, D, gtid);
    if (gtid < D) {
        double a = 0.0;
        for (int n = 0; n < N; ++n) a += (double)__ldg(prm.z_mm + (size_t)n * D + gtid);
        const double mean = a / N;
        double b = 0.0;
        for (int n = 0; n < N; ++n) {
            const double d = (double)__ldg(prm.z_mm + (size_t)n * D + gtid) - mean;
            b += d * d;
        }
        M.st[SD + gtid] = (float)mean;
        M.st[2 * SD + gtid] = 1.f / sqrtf((float)(b / (double)(N - 1)));
    }
    CT_SYNC(g);
}

// Cholesky factor of the D x D matrix A (lower triangle) in fp32, by one thread with compact loops.  Short dependent
// scalar code runs at ~10 cycles per instruction here (measured for D = 4 with IEEE sqrtf and divisions: 4.4 k cycles per
// step as loops, 4.4 k fully unrolled in registers, 6.2 k as one warp with lane = row and shuffles), so the instruction
// count is what matters: one rsqrtf per pivot (MUFU.RSQ, <= 2 ulp) replaces the IEEE square root and the column's
// divisions -- a deliberate deviation of a few ulp in L, the size of the reordering error between any two fp32
// Cholesky implementations.
// ... in registers for D <= DD (the matrix never touches shared memory between the first load and the last store:
// the loop version below spends its time in dependent LDS / STS round trips)
template <int DD>
__device__ __forceinline__ bool cmm_cholesky_reg(const CMM &M, int D) {
    float a[DD][DD], rinv[DD];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < DD; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) a[i][j] = (i < D) ? M.A[i * SD + j] : (i == j ? 1.f : 0.f);
#pragma unroll
    for (int j = 0; j < DD; ++j) {          // column by column (right-looking: same subtraction order per entry)
        float d = a[j][j];
        if (!(d > 0.f)) { if (j < D) ok = false; d = 1.f; }
        rinv[j] = rsqrtf(d);
        a[j][j] = d * rinv[j];
#pragma unroll
        for (int i = j + 1; i < DD; ++i) a[i][j] *= rinv[j];
#pragma unroll
        for (int k = j + 1; k < DD; ++k)
#pragma unroll
            for (int i = k; i < DD; ++i) a[i][k] = fmaf(-a[i][j], a[k][j], a[i][k]);
    }
#pragma unroll
    for (int i = 0; i < DD; ++i) {
        if (i < D) {
            M.rinv[i] = rinv[i];
#pragma unroll
            for (int j = 0; j < DD; ++j)
                if (j < D) M.L[i * SD + j] = j <= i ? a[i][j] : 0.f;
        }
    }
    return ok;
}
__device__ __forceinline__ bool cmm_cholesky(const CMM &M, int D) {
    if (D <= 6) return cmm_cholesky_reg<6>(M, D);
    bool ok = true;
#pragma unroll 1
    for (int i = 0; i < D; ++i) {
#pragma unroll 1
        for (int j = 0; j <= i; ++j) {
            float s = M.A[i * SD + j];
#pragma unroll 1
            for (int k = 0; k < j; ++k) s -= M.L[i * SD + k] * M.L[j * SD + k];
            if (i == j) {
                if (!(s > 0.f)) { ok = false; s = 1.f; }
                const float r = rsqrtf(s);
                M.L[i * SD + i] = s * r;
                M.rinv[i] = r;                      // 1 / L_ii for the rest of the column
            } else {
                M.L[i * SD + j] = s * M.rinv[j];
            }
        }
#pragma unroll 1
        for (int j = i + 1; j < D; ++j) M.L[i * SD + j] = 0.f;
    }
    return ok;
}

// ---- forward: the tile's pre-matching states (role-B registers) -> the matched ones.  M.zrow[slot][j] holds
//      z_mm[(t + n) mod N][j] of the tile's slots; nvg = particles of this tile. ----
__device__ __forceinline__ void cmm_forward(const ClusterParams &prm, const CMM &M, int g, int gtid, int rank, int t, unsigned target,
                                            int nvg, bool roleB, int b_p, int b_d, int b_n, bool b_own, float &s_reg, bool leader,
                                            long long *dbgp) {
#define CMM_MARK(i) do { if (dbgp && (threadIdx.x & 31) == 0) dbgp[(i) * 8 + (threadIdx.x >> 5)] = clock64(); } while (0)
    const int D = prm.D, N = prm.n_global;          // statistics over the particles of all ranks
    const int nq = D + D * (D + 1) / 2;
    const int ntiles = 2 * (int)cmm_clusters(prm.N, prm.PG) * max(prm.mm_world, 1);
    float *s1pre = const_cast<float *>(prm.s1pre);
    const double *rec = prm.mmrec + (size_t)(t & 1) * ntiles * nq * 2;      // 16-byte entries
    if (roleB) {
        M.xrow[b_p * SD + b_d] = s_reg;
        if (b_own) s1pre[((size_t)t * prm.N + b_n) * D + b_d] = s_reg;
    }
    CT_SYNC(g);
    // record of this tile (raw moments in double: products of two floats are exact): sum x_q, sum x_i x_j
    if (rank == 0) {
        for (int q = gtid; q < nq; q += CL_GT) {        // nq <= 135: one or two quantities per thread
            double a = 0.0;
            if (q < D) {
                for (int p = 0; p < nvg; ++p) a += (double)M.xrow[p * SD + q];
            } else {
                const int ij = M.qtab[q - D], i = ij >> 8, j = ij & 255;
                for (int p = 0; p < nvg; ++p) a += (double)M.xrow[p * SD + i] * (double)M.xrow[p * SD + j];
            }
            cmm_publish(prm, t & 1, g, nq, q, a, target);
        }
    }
    CMM_MARK(0);
    cmm_combine(M, rec, ntiles, nq, target, g, gtid);
    CMM_MARK(1);
    if (gtid < D) {
        M.dmean[gtid] = M.red[gtid] / N;
        M.st[gtid] = (float)M.dmean[gtid];
    }
    CT_SYNC(g);
    CMM_MARK(2);
    for (int q = D + gtid; q < nq; q += CL_GT) {
        const int ij = M.qtab[q - D], i = ij >> 8, j = ij & 255;
        // unbiased covariance + jitter (rollout.py:24), handed to the fp32 Cholesky
        const double c = (M.red[q] - (double)N * M.dmean[i] * M.dmean[j]) / (double)(N - 1);
        M.A[i * SD + j] = (float)c + (i == j ? 1e-12f : 0.f);
    }
    CT_SYNC(g);
    CMM_MARK(3);
    if (gtid == 0) {
        const bool ok = cmm_cholesky(M, D);
        if (!ok && prm.status) atomicCAS(prm.status, 0, 1 + t);
    }
    CMM_MARK(5);
    CT_SYNC(g);
    CMM_MARK(6);
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
    CMM_MARK(4);
#undef CMM_MARK
}

// ---- reverse: the cotangent of the matched particles (gs[slot][d], in place) -> the cotangent of the pre-matching
//      ones.  cmm_backward_prefetch (constant inputs of the step): (m, L) of step t, the standardised z rows and the
//      pre-matching states of the tile's slots. ----
__device__ __forceinline__ void cmm_backward_prefetch(const ClusterParams &prm, const CMM &M, int g, int gtid, int t, bool roleB,
                                                      int b_p, int b_d, int b_n) {
    const int D = prm.D, N = prm.n_global;
    const float *ms = prm.mmstat + (size_t)t * (3 * SD + SD * SD);
    for (int i = gtid; i < D; i += CL_GT) M.st[i] = __ldcg(ms + i);
    for (int i = gtid; i < SD * SD; i += CL_GT) M.L[i] = __ldcg(ms + 3 * SD + i);
    if (roleB) {
        int r = t + prm.n_off + b_n;
        r -= (r / N) * N;
        M.zrow[b_p * SD + b_d] = (__ldg(prm.z_mm + (size_t)r * D + b_d) - M.st[SD + b_d]) * M.st[2 * SD + b_d];
        M.xrow[b_p * SD + b_d] = __ldcg(prm.s1pre + ((size_t)t * prm.N + b_n) * D + b_d);
    }
}
__device__ __forceinline__ void cmm_backward(const ClusterParams &prm, const CMM &M, int g, int gtid, int rank, int t, unsigned target,
                                             int nvg, float *gs, bool roleB, int b_p, int b_d) {
    const int D = prm.D, N = prm.n_global;
    const int nq = D + D * (D + 1) / 2;
    const int ntiles = 2 * (int)cmm_clusters(prm.N, prm.PG) * max(prm.mm_world, 1);
    const double *rec = prm.mmrec + (size_t)(t & 1) * ntiles * nq * 2;      // 16-byte entries
    CT_SYNC(g);         // the prefetched rows are in place
    // record of this tile: dm = sum_p g_p (q < D);  dL = tril(sum_p g_p zhat_p^T) (q = D + i (i + 1) / 2 + j)
    if (rank == 0) {
        for (int q = gtid; q < nq; q += CL_GT) {
            double a = 0.0;
            if (q < D) {
                for (int p = 0; p < nvg; ++p) a += (double)gs[p * SD + q];
            } else {
                const int ij = M.qtab[q - D], i = ij >> 8, j = ij & 255;
                for (int p = 0; p < nvg; ++p) a += (double)gs[p * SD + i] * (double)M.zrow[p * SD + j];
            }
            cmm_publish(prm, t & 1, g, nq, q, a, target);
        }
    }
    cmm_combine(M, rec, ntiles, nq, target, g, gtid);
    for (int q = gtid; q < nq; q += CL_GT) {
        if (q < D) {
            M.dm[q] = (float)M.red[q];
        } else {
            const int ij = M.qtab[q - D], i = ij >> 8, j = ij & 255;
            M.Sb[i * SD + j] = (float)M.red[q];   // dL staged in Sb (rewritten below, after its last use)
        }
    }
    if (gtid < D) M.rinv[gtid] = 1.f / M.L[gtid * SD + gtid];
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
    // X = L^-T A  (back substitution, one column per thread; reciprocal pivots M.rinv formed once per step above);
    // in registers for D <= 6 (dependent LDS / STS round trips are what the loop version spends its time on)
    if (D <= 6) {
        constexpr int DD = 6;
        float l[DD][DD], ri[DD], x[DD];
        if (gtid < D) {
#pragma unroll
            for (int i = 0; i < DD; ++i) {
                ri[i] = i < D ? M.rinv[i] : 0.f;
#pragma unroll
                for (int k = 0; k <= i; ++k) l[i][k] = i < D ? M.L[i * SD + k] : 0.f;
            }
#pragma unroll
            for (int r = DD - 1; r >= 0; --r) {
                float sacc = r < D ? M.A[r * SD + gtid] : 0.f;
#pragma unroll
                for (int k = r + 1; k < DD; ++k) sacc = fmaf(-l[k][r], x[k], sacc);
                x[r] = sacc * ri[r];
            }
#pragma unroll
            for (int r = 0; r < DD; ++r)
                if (r < D) M.X[r * SD + gtid] = x[r];
        }
        CT_SYNC(g);
        // Sb = X L^-1  (one row per thread)
        if (gtid < D) {
#pragma unroll
            for (int c = DD - 1; c >= 0; --c) {
                float sacc = c < D ? M.X[gtid * SD + c] : 0.f;
#pragma unroll
                for (int k = c + 1; k < DD; ++k) sacc = fmaf(-x[k], l[k][c], sacc);
                x[c] = sacc * ri[c];
            }
#pragma unroll
            for (int c = 0; c < DD; ++c)
                if (c < D) M.Sb[gtid * SD + c] = x[c];
        }
        CT_SYNC(g);
    } else {
    if (gtid < D) {
        const int j = gtid;
#pragma unroll 1
        for (int r = D - 1; r >= 0; --r) {
            float s = M.A[r * SD + j];
#pragma unroll 1
            for (int k = r + 1; k < D; ++k) s -= M.L[k * SD + r] * M.X[k * SD + j];
            M.X[r * SD + j] = s * M.rinv[r];
        }
    }
    CT_SYNC(g);
    // Sb = X L^-1  (one row per thread)
    if (gtid < D) {
        const int i = gtid;
#pragma unroll 1
        for (int c = D - 1; c >= 0; --c) {
            float s = M.X[i * SD + c];
#pragma unroll 1
            for (int k = c + 1; k < D; ++k) s -= M.Sb[i * SD + k] * M.L[k * SD + c];
            M.Sb[i * SD + c] = s * M.rinv[c];
        }
    }
    CT_SYNC(g);
    }
    // dx_n = dm/N + 2/(N-1) * sym(Sb) (x_n - m)
    if (roleB) {
        float a = 0.f;
        for (int j = 0; j < D; ++j)
            a = fmaf(0.5f * (M.Sb[b_d * SD + j] + M.Sb[j * SD + b_d]), M.xrow[b_p * SD + j] - M.st[j], a);
        gs[b_p * SD + b_d] = M.dm[b_d] / (float)N + (2.f / (float)(N - 1)) * a;
    }
    CT_SYNC(g);
}

}  // namespace pmb
