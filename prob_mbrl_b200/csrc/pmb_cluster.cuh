// Cluster-resident sweeps (sm_100a): the imagined rollout with ALL weights resident in the shared memory of a
// thread-block cluster, no per-step weight traffic at all.
//
// A cluster of C CTAs owns PG <= 8 particles for the whole horizon.  For the reference's two-hidden-layer
// nets (Policy / DynamicsModel built by models.mlp, reference models/core.py:39-73) one net pass is
//   thin   : K <= 16 inputs  -> hidden a (full width)      computed redundantly by every CTA of the cluster
//   wide   : hidden a (K <= 256) -> hidden b                COLUMN-SPLIT: CTA r owns `hs` columns of hidden b;
//                                                            its [K][hs] slice of the matrix lives in its smem
//   narrow : hidden b -> <= 16 outputs                      every CTA forms the partial sums over ITS columns
//   exchange                                                the partials (PG x outputs floats) go to every CTA of
//                                                            the cluster with st.async (DSMEM store + mbarrier
//                                                            complete_tx in one instruction); every CTA adds the C
//                                                            partials in rank order (deterministic, identical)
// so the only inter-CTA traffic per net pass is C x PG x outputs floats, and each CTA keeps a full copy of
// the (tiny) per-particle state.  The reverse sweep has the same shape with the transposed matrices
// (thin = output-projection adjoint, wide = hidden x hidden adjoint split over the columns of hidden 0,
// narrow = input-projection adjoint).  Inner products run on the packed FP32 pipe (FFMA2).
#pragma once
#include "pmb_internal.cuh"

namespace pmb {

constexpr int CL_NT = 256;     // threads per CTA (8 warps, all compute)
constexpr int CL_PS = 8;       // particle slots per cluster (two tiles of CL_TS slots)
constexpr int CL_HS = 32;      // widest column slice per CTA
constexpr int CL_TW = 256;     // widest thin layer (= K of the wide layer)
constexpr int CL_NO = 16;      // narrow outputs / thin inputs (max)
constexpr int CL_TS = 4;                  // particle slots per tile (= warps per warp group; two tiles per cluster)
constexpr int CL_GT = 128;                // threads per warp group
constexpr int CL_MBOX = CL_TS * CL_NO;    // floats one (CTA, group) sends per exchange

// One net in one direction.
struct CNet {
    int tK, tW;               // thin layer: rows (inputs), padded width (= K of the wide layer)
    int tsl;                  // columns of the thin output this CTA stores to global (multiple of 4)
    int wN;                   // padded full width of the wide layer's output
    int hs;                   // columns of the wide layer per CTA (multiple of 4, <= CL_HS)
    int nN, nNp;              // narrow layer: outputs, padded to a multiple of 4
    long long t_goff, w_goff, n_goff;        // packed matrices (float offsets inside this sweep's packed area)
    long long tb_off, wb_off, nb_off;        // padded biases in the workspace, -1 = none (forward only)
    long long tm_off, wm_off;                // dropout masks [N][tW] / [N][wN] in the workspace, -1 = none
    float tkeep_inv, wkeep_inv;              // 1/keep of the thin-output / wide-output hidden layer
    long long tsav_off, wsav_off;            // stored activations [H][N][tW] / [H][N][wN]
    long long tdel_off, wdel_off, odel_off;  // backward, policy: adjoints kept for the weight gradient, -1 = none
    long long raw_off;                       // raw outputs of the net [H][N][nraw]
    int nraw;
    int has_density;
    float lmax;
    const float *z;
    long long zstride;
    int s_tw, s_ww, s_tb, s_wb, s_nb, s_tm, s_wm;         // shared-memory offsets (floats)
    int s_nwt;                // narrow matrix of this CTA's columns as [4][CL_HS][4]: (o >> 2, column, o & 3)
};

struct ClusterParams {
    int N, H, D, U;
    int PG;                     // particles per cluster
    int pingpong;               // != 0: the two particle tiles of a CTA alternate on the LSU-bound phases
    int C;                      // CTAs per cluster
    CNet pol, dyn;
    const float *wpack;         // packed weights of THIS sweep
    float *ws;                  // workspace base
    const float *act_scale, *act_bias, *mx, *iSx, *my, *Sy;
    int KR;
    const float *rew_C, *rew_c0, *rew_Q, *rew_R;
    float rew_scale, rew_offset;
    const float *x0;
    float *states, *actions, *rewards;
    const float *g_states, *g_actions, *g_rewards;
    float *dx0;
    float *da_total;            // backward: total dL/da_t [H][N][U] (nullable)
    float *pre;                 // backward: [H][N][2D + 3U] step-local adjoint factors (bwd_pre_kernel)
    const float *s1pre;         // next states BEFORE moment matching [H][N][D] (the reward acts on them), nullptr = states[t+1]
    long long *dbg;             // clock64() marks of cluster 0 / rank 0 at step H/2 (nullable)
    // moment matching of the states inside the cluster-resident sweeps (pmb_cluster_mm.cuh)
    int mm_states;              // != 0: one matching group = all N particles (N <= 128)
    const float *z_mm;          // [>= N][D]
    float *mmstat;              // [H][3*SD + SD*SD] mean, z mean, 1/z std, Cholesky factor of every step
    unsigned *mmctr;            // arrival counter of this sweep (zeroed before the launch)
    double *mmrec;              // [2][tiles][CMM_NQ] per-tile records of a step (double-buffered by step parity)
    // ... across GPUs (particles sharded over mm_world ranks of one node): every rank's record area and
    // arrival counter are mapped into every other rank (CUDA IPC); records and arrivals go to ALL ranks over NVLink
    int mm_world, mm_rank;      // 1, 0: single GPU
    int n_global, n_off;        // particles of all ranks / first global particle index of this rank
    double *mmrec_peer[16];     // [rank]: that rank's record area of THIS sweep (entry mm_rank == mmrec)
    unsigned *mmctr_peer[16];   // [rank]: that rank's arrival counter of THIS sweep (entry mm_rank == mmctr)
    const unsigned long long *mm_base;   // arrivals before this launch (counters across GPUs are never reset), or nullptr = 0
    unsigned long long *mm_base_next;    // where the launch leaves the arrivals after it (same word), or nullptr
    int *status;                // forward: 1 + first step whose covariance was not positive definite
    int off_mm;                 // shared memory: two CMM blocks (one per particle tile)
    unsigned *g1, *g2;          // wide cluster-resident sweeps (pmb_cw.cuh): ReLU/dropout gate bit words of hidden 0 / 1
    int ncl;                    // ... clusters of the launch
    int off_cst, off_xa, off_xb, off_act, off_red, off_inbox, off_misc;
    int smem_floats;
};

// ----------------------------------------------------------------------------------------
// cluster PTX helpers
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cl_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cl_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// 16-byte store into a peer CTA's shared memory that also signals 16 bytes on the peer's mbarrier
__device__ __forceinline__ void cl_st_async_v4(uint32_t daddr, float4 v, uint32_t dbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(daddr),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(dbar)
                 : "memory");
}
__device__ __forceinline__ float2 cl_fma2(float a, float2 w, float2 c) {
    return __ffma2_rn(make_float2(a, a), w, c);
}

// per-warp arrival marks (lane 0 of every warp), placed BEFORE barriers: dbg[i * 8 + warp]
#define CL_TMARK(i) do { if (dbg_step && (threadIdx.x & 31) == 0) prm.dbg[(i) * 8 + (threadIdx.x >> 5)] = clock64(); } while (0)

// ----------------------------------------------------------------------------------------
// resident operands of one net: thin matrix, this CTA's column slice of the wide and narrow matrices,
// biases (forward), the cluster's rows of the dropout masks (1.0 where the layer has no mask)
// ----------------------------------------------------------------------------------------
// Mask rows are indexed by the cluster's particle slot q: tile 0 = slots 0..3 holds the particles n0 .. n0+nv0-1,
// tile 1 = slots 4..7 the particles n0+nv0 ...; slots past the tile's share repeat a valid particle.
__device__ __forceinline__ int cl_slot_particle(int q, int n0, int nv0, int N) {
    return min(n0 + (q < CL_TS ? q : nv0 + q - CL_TS), N - 1);
}
__device__ __forceinline__ void cl_load_net(const ClusterParams &prm, const CNet &n, float *smem, int rank, int n0,
                                            int nv0, bool fwd) {
    const int tid = threadIdx.x;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const float4 *src = reinterpret_cast<const float4 *>(prm.wpack + n.t_goff);
        float4 *dst = reinterpret_cast<float4 *>(smem + n.s_tw);
        for (int i = tid; i < (n.tK * n.tW) / 4; i += CL_NT) dst[i] = __ldg(src + i);
    }
    const int hs4 = n.hs >> 2;
    for (int i = tid; i < n.tW * hs4; i += CL_NT) {
        const int k = i / hs4, c4 = i - k * hs4;
        const int gc = rank * n.hs + 4 * c4;
        float4 v = z4;
        if (gc < n.wN) v = __ldg(reinterpret_cast<const float4 *>(prm.wpack + n.w_goff + (long long)k * n.wN + gc));
        *reinterpret_cast<float4 *>(smem + n.s_ww + k * n.hs + 4 * c4) = v;
    }
    for (int i = tid; i < CL_NO * CL_HS; i += CL_NT) {
        const int o = i / CL_HS, c = i - o * CL_HS;
        const int gc = rank * n.hs + c;
        float v = 0.f;
        if (o < n.nN && c < n.hs && gc < n.wN) v = __ldg(prm.wpack + n.n_goff + (long long)o * n.wN + gc);
        smem[n.s_nwt + (o >> 2) * (CL_HS * 4) + c * 4 + (o & 3)] = v;
    }
    if (fwd) {
        for (int j = tid; j < n.tW; j += CL_NT) smem[n.s_tb + j] = n.tb_off >= 0 ? __ldg(prm.ws + n.tb_off + j) : 0.f;
        for (int c = tid; c < n.hs; c += CL_NT) {
            const int gc = rank * n.hs + c;
            smem[n.s_wb + c] = (n.wb_off >= 0 && gc < n.wN) ? __ldg(prm.ws + n.wb_off + gc) : 0.f;
        }
        for (int o = tid; o < CL_NO; o += CL_NT) smem[n.s_nb + o] = (n.nb_off >= 0 && o < n.nN) ? __ldg(prm.ws + n.nb_off + o) : 0.f;
    }
    for (int i = tid; i < CL_PS * n.tW; i += CL_NT) {
        const int p = i / n.tW, j = i - p * n.tW;
        const int nn = cl_slot_particle(p, n0, nv0, prm.N);
        smem[n.s_tm + i] = n.tm_off >= 0 ? __ldg(prm.ws + n.tm_off + (long long)nn * n.tW + j) : 1.f;
    }
    for (int i = tid; i < CL_PS * n.hs; i += CL_NT) {
        const int p = i / n.hs, c = i - p * n.hs;
        const int nn = cl_slot_particle(p, n0, nv0, prm.N);
        const int gc = rank * n.hs + c;
        float v = 0.f;
        if (gc < n.wN) v = n.wm_off >= 0 ? __ldg(prm.ws + n.wm_off + (long long)nn * n.wN + gc) : 1.f;
        smem[n.s_wm + i] = v;
    }
}

// ----------------------------------------------------------------------------------------
// Two independent particle tiles per CTA.  The 8 warps of a CTA form two groups of 4 warps; group g of every
// CTA of the cluster works on the cluster's particle slots [4g, 4g+4) and never synchronises with the other
// group (named barrier 1+g, its own exchange mailboxes and mbarriers): while one group waits on a barrier, on
// the exchange or on a dependent-latency chain, the other group's warps keep the LSU / FMA pipes busy.  Both
// groups read the same resident weights.
// ----------------------------------------------------------------------------------------
#define CT_SYNC(g) asm volatile("bar.sync %0, 128;" ::"r"((g) + 2) : "memory")   // barriers 2, 3 (0, 1 = whole CTA)
// LSU hand-over between the two tiles of a CTA (barriers 4, 5; both groups' threads are counted): the thin + wide
// phases are bound by shared-memory bandwidth, the epilogue / exchange / per-particle phases by latency, so the two
// groups alternate: a group enters its thin phase only after the other one finished its wide accumulate.
#define CT_LSU_ACQUIRE(g) asm volatile("bar.sync %0, 256;" ::"r"((g) + 4) : "memory")
#define CT_LSU_RELEASE(g) asm volatile("bar.arrive %0, 256;" ::"r"(5 - (g)) : "memory")

// wide layer, this CTA's column slice, one 4-slot tile: 16 k-slices = (warp of the group, quarter-warp); lane =
// (row quarter, column quad q).  Per row one LDS.128 of weights + one LDS.128 of activations (all 4 slots) feed
// 8 FFMA2 (4 slots x 4 columns per thread).  Partial sums go to red[slice][slot][32].
constexpr int CL_KS = 16;                 // k-slices of the wide layer per tile
__device__ __forceinline__ void ct_wide_accum(const float *__restrict__ ww, int K, int hs, const float *__restrict__ act,
                                              float *__restrict__ red, int gtid) {
    const int lane = gtid & 31, w = gtid >> 5;
    const int q = lane & 7;
    if (4 * q >= hs) return;
    const int slice = 4 * w + (lane >> 3);
    const float2 z2 = make_float2(0.f, 0.f);
    float2 a0 = z2, a1 = z2, b0 = z2, b1 = z2, c0 = z2, c1 = z2, d0 = z2, d1 = z2;   // slots 0..3 x column pairs
    const float *wp = ww + slice * hs + 4 * q;
    const float *ap = act + slice * CL_TS;
    const int n = (K - slice + CL_KS - 1) / CL_KS;
    const int wstride = CL_KS * hs;
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        const float4 wv = *reinterpret_cast<const float4 *>(wp);
        const float4 xv = *reinterpret_cast<const float4 *>(ap);
        wp += wstride;
        ap += CL_KS * CL_TS;
        const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
        a0 = cl_fma2(xv.x, w01, a0);
        a1 = cl_fma2(xv.x, w23, a1);
        b0 = cl_fma2(xv.y, w01, b0);
        b1 = cl_fma2(xv.y, w23, b1);
        c0 = cl_fma2(xv.z, w01, c0);
        c1 = cl_fma2(xv.z, w23, c1);
        d0 = cl_fma2(xv.w, w01, d0);
        d1 = cl_fma2(xv.w, w23, d1);
    }
    float *r = red + ((slice * CL_TS) << 5) + 4 * q;
    *reinterpret_cast<float4 *>(r) = make_float4(a0.x, a0.y, a1.x, a1.y);
    *reinterpret_cast<float4 *>(r + 32) = make_float4(b0.x, b0.y, b1.x, b1.y);
    *reinterpret_cast<float4 *>(r + 64) = make_float4(c0.x, c0.y, c1.x, c1.y);
    *reinterpret_cast<float4 *>(r + 96) = make_float4(d0.x, d0.y, d1.x, d1.y);
}
// finished value of (slot = warp of the group, column = lane): the 16 k-slices in a fixed order
__device__ __forceinline__ float ct_wide_reduce(const float *__restrict__ red, int gtid) {
    const float *r = red + ((gtid >> 5) << 5) + (gtid & 31);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int s = 0; s < CL_KS; s += 4) {
        s0 += r[(s + 0) * CL_TS * 32];
        s1 += r[(s + 1) * CL_TS * 32];
        s2 += r[(s + 2) * CL_TS * 32];
        s3 += r[(s + 3) * CL_TS * 32];
    }
    return (s0 + s1) + (s2 + s3);
}

// The shared::cluster window of CTA r is the window of CTA 0 shifted by r * stride (checked once per kernel),
// so one mapa per exchange replaces one per destination.
__device__ __forceinline__ uint32_t cl_window_stride(uint32_t probe_saddr, int C) {
    const uint32_t a0 = cl_mapa(probe_saddr, 0);
    const uint32_t stride = cl_mapa(probe_saddr, 1) - a0;
    for (int r = 2; r < C; ++r)
        if (cl_mapa(probe_saddr, r) != a0 + (uint32_t)r * stride) __trap();
    return stride;
}

// Narrow layer + exchange fused into the wide layer's epilogue.  Warp = particle slot, lane = column of this
// CTA's slice; `v` is the lane's finished hidden value (0 on idle lanes).  The warp forms
//   out[o] = sum_lanes v * nw[o][lane]   for o < NV
// with a reduce-scatter butterfly (NV values per lane -> 1), four neighbouring holders are gathered into one
// lane and that lane sends the 16-byte chunk (slot, 4 outputs) to every CTA of the cluster: st.async = DSMEM
// store + complete_tx on the destination's mbarrier in one instruction.  NV = 4, 8 or 16.
//   mbox_saddr : this (group, exchange)'s mailbox [C ranks][CL_MBOX] (same offset in every CTA)
//   slot_off   : byte offset of (sender rank, slot) inside the mailbox
template <int C, int NV>
__device__ __forceinline__ void ct_narrow_send(float v, const float *__restrict__ nwt, bool send_ok, int nN,
                                               uint32_t mbox_saddr, uint32_t slot_off, uint32_t bar_saddr,
                                               uint32_t wstride, long long *dbgp) {
    const int lane = threadIdx.x & 31;
    float pr[NV];
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
        const float4 w = *reinterpret_cast<const float4 *>(nwt + i * (CL_HS * 4) + lane * 4);
        pr[4 * i] = v * w.x;
        pr[4 * i + 1] = v * w.y;
        pr[4 * i + 2] = v * w.z;
        pr[4 * i + 3] = v * w.w;
    }
    int m = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1, m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = up ? pr[n / 2 + i] : pr[i];
            const float give = up ? pr[i] : pr[n / 2 + i];
            pr[i] = keep + __shfl_xor_sync(0xffffffffu, give, m);
        }
    }
#pragma unroll
    for (; m >= 1; m >>= 1) pr[0] += __shfl_xor_sync(0xffffffffu, pr[0], m);
    // output o = lane / s sits in every lane of its group of s = 32 / NV lanes.  Lane = (chunk, destination rank)
    // inside a group of G = 4s lanes fetches the chunk's four outputs straight from their holders, so that ONE
    // st.async instruction serves every destination.
    constexpr int s = 32 / NV;
    constexpr int G = 4 * s;
    const int src = lane & ~(G - 1);
    const float4 out = make_float4(__shfl_sync(0xffffffffu, pr[0], src), __shfl_sync(0xffffffffu, pr[0], src + s),
                                   __shfl_sync(0xffffffffu, pr[0], src + 2 * s), __shfl_sync(0xffffffffu, pr[0], src + 3 * s));
    if (dbgp && lane == 0) dbgp[0] = clock64();
    const int dst = lane & (G - 1), chunk = lane / G;
    if (send_ok && dst < C && 4 * chunk < nN) {
        const uint32_t a0 = cl_mapa(mbox_saddr + slot_off + (uint32_t)chunk * 16u, 0), b0 = cl_mapa(bar_saddr, 0);
        cl_st_async_v4(a0 + (uint32_t)dst * wstride, out, b0 + (uint32_t)dst * wstride);
    }
}
template <int C>
__device__ __forceinline__ void ct_narrow_send_any(float v, const float *__restrict__ nwt, bool send_ok, int nN,
                                                   uint32_t mbox_saddr, uint32_t slot_off, uint32_t bar_saddr,
                                                   uint32_t wstride, long long *dbgp) {
    if (nN <= 4) ct_narrow_send<C, 4>(v, nwt, send_ok, nN, mbox_saddr, slot_off, bar_saddr, wstride, dbgp);
    else if (nN <= 8) ct_narrow_send<C, 8>(v, nwt, send_ok, nN, mbox_saddr, slot_off, bar_saddr, wstride, dbgp);
    else ct_narrow_send<C, 16>(v, nwt, send_ok, nN, mbox_saddr, slot_off, bar_saddr, wstride, dbgp);
}
// value (slot, o) from a mailbox: the C partials added in a fixed pairwise order (identical on every CTA)
template <int C>
__device__ __forceinline__ float ct_gather(const float *mbox, int slot, int o) {
    const float *q = mbox + slot * CL_NO + o;
    float v[C];
#pragma unroll
    for (int r = 0; r < C; ++r) v[r] = q[r * CL_MBOX];
#pragma unroll
    for (int w = 1; w < C; w <<= 1)
#pragma unroll
        for (int r = 0; r + w < C; r += 2 * w) v[r] += v[r + w];
    return v[0];
}
// exp(clamp_logstd(l)) on the serial chain of the forward sweep: exp(lmax) / (1 + exp(lmax - l)) with the SFU
// exponential and reciprocal (a few ulp; the value only scales the density noise)
__device__ __forceinline__ float ct_exp_clamped_logstd(float l, float lmax, float elmax) {
    return __fdividef(elmax, 1.f + __expf(lmax - l));
}

cudaError_t launch_cluster_fwd(const ClusterParams &prm, int nclusters, cudaStream_t stream);
cudaError_t launch_cluster_bwd(const ClusterParams &prm, int nclusters, cudaStream_t stream);
cudaError_t launch_bwd_pre(const ClusterParams &prm, cudaStream_t stream);   // the adjoint-factor pre-pass alone
int cluster_max_active(int C, int smem_bytes, bool fwd);

}  // namespace pmb
