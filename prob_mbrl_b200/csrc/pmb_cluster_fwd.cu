// Forward sweep, cluster-resident variant (see pmb_cluster.cuh): H steps of
//   policy MLP -> Gaussian action sample -> tanh squash -> dynamics MLP -> Gaussian state sample -> reward
// for PG particles per cluster of C CTAs, all weights resident in the cluster's shared memory.
// Replaces the loop body of utils.rollout (reference utils/rollout.py:93-163) with Policy.forward
// (models/core.py:221-248), DynamicsModel.forward (models/core.py:265-303), B/CDropout masks
// (models/modules.py:61,160), DiagGaussianDensity (models/densities.py:87-121) and the env reward
// (envs/cartpole/env.py:41-86 et al.).  Same workspace layout as the streaming sweep (pmb_rollout_fwd.cu),
// so the reverse sweep and the weight-gradient kernels of either variant can follow.
#include "pmb_cluster.cuh"
#include "pmb_cluster_mm.cuh"
#include "pmb_host.h"

namespace pmb {

// Per-thread constants of a net pass, fixed for the whole horizon (registers).
//   thin layer : thread gtid of the group owns the columns gtid and gtid + 128 of hidden 0: their TK weights, bias
//                and the 4 particle slots' mask / keep
//   epilogue of the wide layer: thread = (particle slot = warp of the group, column = lane of this CTA's slice)
template <int TK>
struct FwdNetRegs {
    float tw[2][TK];
    float tb[2];
    float tmk[2][CL_TS];
    float wb, wmk;
    bool on0, on1, thin_store, wide_store, send_ok;
    float *sv_thin, *sv_wide;       // running global pointers of the stored activations (advance per step)
    size_t thin_step, wide_step;
    int tK, tcol;
    uint32_t slot_off;              // byte offset of (this rank, this slot) inside a mailbox
    __device__ __forceinline__ void init(const ClusterParams &prm, const CNet &n, const float *smem, int rank, int g,
                                         int gtid, int n0g, int nvg) {
        const int lane = gtid & 31, sl = gtid >> 5;     // slot of the tile
        const int ps = g * CL_TS + sl;                  // slot of the cluster (mask rows)
        on0 = gtid < n.tW;
        on1 = gtid + CL_GT < n.tW;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const bool on = c ? on1 : on0;
            const int j = gtid + c * CL_GT;
#pragma unroll
            for (int k = 0; k < TK; ++k) tw[c][k] = (on && k < n.tK) ? smem[n.s_tw + k * n.tW + j] : 0.f;
            tb[c] = on ? smem[n.s_tb + j] : 0.f;
#pragma unroll
            for (int q = 0; q < CL_TS; ++q) tmk[c][q] = on ? smem[n.s_tm + (g * CL_TS + q) * n.tW + j] * n.tkeep_inv : 0.f;
        }
        tK = n.tK;
        // hidden 0 is stored by warp = slot, lane = column of this CTA's share of the columns
        tcol = rank * n.tsl + lane;
        thin_store = sl < nvg && lane < n.tsl && tcol < n.tW;
        sv_thin = prm.ws + n.tsav_off + (size_t)(n0g + sl) * n.tW + tcol;
        thin_step = (size_t)prm.N * n.tW;
        const int gc = rank * n.hs + lane;
        const bool wide_on = lane < n.hs;
        wb = wide_on ? smem[n.s_wb + lane] : 0.f;
        wmk = wide_on ? smem[n.s_wm + ps * n.hs + lane] * n.wkeep_inv : 0.f;
        wide_store = wide_on && sl < nvg && gc < n.wN;
        sv_wide = prm.ws + n.wsav_off + (size_t)(n0g + sl) * n.wN + gc;
        wide_step = (size_t)prm.N * n.wN;
        send_ok = sl < nvg;
        slot_off = (uint32_t)(rank * CL_MBOX + sl * CL_NO) * 4u;
    }
};

// One net pass of one tile up to and including the send of the output partials.
//   x : [TK][4] input tile (rows >= tK are zero); on return the partial sums of the raw outputs are on their way
//   to every CTA of the cluster.  Two group barriers.
template <int C, int TK>
__device__ __forceinline__ void ct_net_forward(const ClusterParams &prm, const CNet &n, FwdNetRegs<TK> &R, float *smem,
                                               const float *x, float *act, float *red, int g, int gtid,
                                               uint32_t mbox_saddr, uint32_t bar_saddr, uint32_t wstride, bool dbg_step,
                                               int mark0, bool pingpong) {
    if (pingpong) CT_LSU_ACQUIRE(g);
    // ---- thin: hidden 0 = relu(x W0^T + b0) * mask0 / keep0 (full width, every CTA); two columns per thread ----
    if (R.on0) {
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            // rows >= tK are zero padding (weights and inputs): no branch, every load is issued up front
            const float4 xv = *reinterpret_cast<const float4 *>(x + k * CL_TS);
            a0 = cl_fma2(R.tw[0][k], make_float2(xv.x, xv.y), a0);
            a1 = cl_fma2(R.tw[0][k], make_float2(xv.z, xv.w), a1);
            b0 = cl_fma2(R.tw[1][k], make_float2(xv.x, xv.y), b0);
            b1 = cl_fma2(R.tw[1][k], make_float2(xv.z, xv.w), b1);
        }
        float4 v;
        v.x = fmaxf(a0.x + R.tb[0], 0.f) * R.tmk[0][0];
        v.y = fmaxf(a0.y + R.tb[0], 0.f) * R.tmk[0][1];
        v.z = fmaxf(a1.x + R.tb[0], 0.f) * R.tmk[0][2];
        v.w = fmaxf(a1.y + R.tb[0], 0.f) * R.tmk[0][3];
        *reinterpret_cast<float4 *>(act + gtid * CL_TS) = v;
        if (R.on1) {
            v.x = fmaxf(b0.x + R.tb[1], 0.f) * R.tmk[1][0];
            v.y = fmaxf(b0.y + R.tb[1], 0.f) * R.tmk[1][1];
            v.z = fmaxf(b1.x + R.tb[1], 0.f) * R.tmk[1][2];
            v.w = fmaxf(b1.y + R.tb[1], 0.f) * R.tmk[1][3];
            *reinterpret_cast<float4 *>(act + (gtid + CL_GT) * CL_TS) = v;
        }
    }
    CL_TMARK(mark0);
    CT_SYNC(g);
    // hidden 0 is kept for the reverse sweep: one coalesced row segment per warp (= particle slot), off the tile
    if (R.thin_store) *R.sv_thin = act[R.tcol * CL_TS + (gtid >> 5)];
    R.sv_thin += R.thin_step;
    // ---- wide: this CTA's columns of hidden 1, k-split 16 ways over the quarter-warps of the group ----
    ct_wide_accum(smem + n.s_ww, n.tW, n.hs, act, red, gtid);
    if (pingpong) CT_LSU_RELEASE(g);
    CL_TMARK(mark0 + 1);
    CT_SYNC(g);
    // ---- epilogue (warp = particle slot, lane = column) + narrow partial sums + exchange ----
    {
        float v = ct_wide_reduce(red, gtid);
        v = fmaxf(v + R.wb, 0.f) * R.wmk;           // idle lanes: wb = wmk = 0 and red holds zeros
        if (R.wide_store) *R.sv_wide = v;
        R.sv_wide += R.wide_step;
        CL_TMARK(mark0 + 10);
        ct_narrow_send_any<C>(v, smem + n.s_nwt, R.send_ok, n.nN, mbox_saddr, R.slot_off, bar_saddr, wstride,
                              dbg_step ? prm.dbg + (mark0 + 8) * 8 + (threadIdx.x >> 5) : nullptr);
    }
    CL_TMARK(mark0 + 2);
}

// MM: moment matching of the states compiled in (a separate instantiation: the plain sweeps keep their code size --
// they live off the instruction cache -- and their registers)
template <int C, int TK, bool MM = false>
__global__ void __launch_bounds__(CL_NT, 1) cluster_fwd_kernel(const __grid_constant__ ClusterParams prm) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t xbar[2][2];       // [group][0 policy exchange, 1 dynamics exchange]
    const int tid = threadIdx.x;
    const int g = tid >> 7, gtid = tid & (CL_GT - 1);  // particle tile (warp group) and thread of the group
    const int rank = (int)cl_rank();
    const int PG = prm.PG;
    const int n0 = (int)cl_id_x() * PG;
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H;
    const int nval = min(PG, N - n0);                  // particles of this cluster
    const int nv0 = (nval + 1) >> 1;                   // tile 0 takes the first half (rounded up), tile 1 the rest
    const int n0g = n0 + (g ? nv0 : 0);                // first particle of this tile
    const int nvg = g ? nval - nv0 : nv0;              // particles of this tile (0: the group idles)
    const CNet &pol = prm.pol;
    const CNet &dyn = prm.dyn;

    for (int i = tid; i < prm.smem_floats; i += CL_NT) smem[i] = 0.f;
    __syncthreads();
    float *cst = smem + prm.off_cst;
    const int tw_max = max(pol.tW, dyn.tW);
    float *xpol = smem + prm.off_xa + g * (CL_NO * CL_TS);     // [TK][4] policy input (raw state), rows >= D stay zero
    float *xdyn = smem + prm.off_xb + g * (CL_NO * CL_TS);     // [TK][4] dynamics input (scaled state, scaled action)
    float *act = smem + prm.off_act + g * (tw_max * CL_TS);    // [tW][4] hidden 0 of the tile
    float *red = smem + prm.off_red + g * (CL_KS * CL_TS * 32);    // [8 k-slices][4][32]
    const float *mbox_pol = smem + prm.off_inbox + g * (2 * C * CL_MBOX);
    const float *mbox_dyn = mbox_pol + C * CL_MBOX;
    const uint32_t bytes_pol = (uint32_t)(C * nvg * pol.nNp) * 4u, bytes_dyn = (uint32_t)(C * nvg * dyn.nNp) * 4u;
    if (gtid == 0) {
        mbar_init(&xbar[g][0], 1);
        mbar_init(&xbar[g][1], 1);
        fence_mbar_init();
        mbar_expect_tx(&xbar[g][0], bytes_pol);
        mbar_expect_tx(&xbar[g][1], bytes_dyn);
    }
    load_constants(prm, cst);
    // the mask rows of the cluster's 8 slots: tile 0 = slots 0..3, tile 1 = slots 4..7 -> particle of slot q
    cl_load_net(prm, pol, smem, rank, n0, nv0, true);
    cl_load_net(prm, dyn, smem, rank, n0, nv0, true);
    __syncthreads();
    FwdNetRegs<TK> Rp, Rd;
    Rp.init(prm, pol, smem, rank, g, gtid, n0g, nvg);
    Rd.init(prm, dyn, smem, rank, g, gtid, n0g, nvg);

    // ---- thread roles for the per-particle stages (fixed for the whole horizon) ----
    const bool roleA = gtid < CL_TS * U;                       // one (particle slot, action dim)
    const int a_p = roleA ? gtid / U : 0, a_u = roleA ? gtid - a_p * U : 0;
    const int a_n = min(n0g + a_p, N - 1);
    const bool a_own = roleA && a_p < nvg && ((g * CL_TS + a_p) % C) == rank;
    const bool roleB = gtid >= 64 && gtid - 64 < CL_TS * D;    // one (particle slot, state dim)
    const int b_p = roleB ? (gtid - 64) / D : 0, b_d = roleB ? (gtid - 64) - b_p * D : 0;
    const int b_n = min(n0g + b_p, N - 1);
    const bool b_own = roleB && b_p < nvg && ((g * CL_TS + b_p) % C) == rank;

    float s_reg = 0.f;            // role B: this thread's element of the current state
    float b_mx = 0.f, b_isx = 0.f, b_sy = 0.f, b_my = 0.f, b_nbm = 0.f, b_nbl = 0.f;
    float a_mx = 0.f, a_isx = 0.f, a_sc = 0.f, a_bi = 0.f, a_nbm = 0.f, a_nbl = 0.f;
    if (roleB) {
        b_mx = prm.mx[b_d]; b_isx = prm.iSx[b_d]; b_sy = prm.Sy[b_d]; b_my = prm.my[b_d];
        b_nbm = smem[dyn.s_nb + b_d];
        if (dyn.has_density) b_nbl = smem[dyn.s_nb + D + b_d];
        s_reg = prm.x0[(size_t)b_n * D + b_d];
        xpol[b_d * CL_TS + b_p] = s_reg;
        xdyn[b_d * CL_TS + b_p] = (s_reg - b_mx) * b_isx;
        if (b_own) prm.states[(size_t)b_n * D + b_d] = s_reg;
    }
    if (roleA) {
        a_mx = prm.mx[D + a_u]; a_isx = prm.iSx[D + a_u]; a_sc = prm.act_scale[a_u]; a_bi = prm.act_bias[a_u];
        a_nbm = smem[pol.s_nb + a_u];
        if (pol.has_density) a_nbl = smem[pol.s_nb + U + a_u];
    }
    const float elmax_pol = expf(pol.lmax), elmax_dyn = expf(dyn.lmax);
    float zA = 0.f, zB = 0.f;
    if (roleA && pol.has_density) zA = __ldg(pol.z + (size_t)a_n * U + a_u);
    if (roleB && dyn.has_density) zB = __ldg(dyn.z + (size_t)b_n * D + b_d);
    const uint32_t mbox_pol_saddr = smem_u32(mbox_pol), mbox_dyn_saddr = smem_u32(mbox_dyn);
    const uint32_t bar_pol = smem_u32(&xbar[g][0]), bar_dyn = smem_u32(&xbar[g][1]);
    const uint32_t wstride = cl_window_stride(bar_pol, C);
    // running global pointers of the role threads (advance per step)
    float *act_ptr = prm.actions + (size_t)a_n * U + a_u;
    float *rawp_ptr = prm.ws + pol.raw_off + (size_t)a_n * pol.nraw + a_u;
    float *st_ptr = prm.states + ((size_t)N + b_n) * D + b_d;
    float *rawd_ptr = prm.ws + dyn.raw_off + (size_t)b_n * dyn.nraw + b_d;
    const size_t act_step = (size_t)N * U, rawp_step = (size_t)N * pol.nraw, st_step = (size_t)N * D,
                 rawd_step = (size_t)N * dyn.nraw;

    // moment matching of the states: scratch of this tile, arrivals per step, z statistics (constants of the launch)
    constexpr bool mm = MM;
    CMM M;
    M.carve(smem + prm.off_mm + g * CMM_FLOATS);
    const unsigned mm_ncl = mm ? cmm_arrivals(prm) : 0u;         // arrivals per step: the clusters of all ranks
    const unsigned mm_base = mm ? cmm_base(prm) : 0u;
    const bool mm_leader = blockIdx.x == 0 && g == 0;

    __syncthreads();
    cl_sync();          // every CTA's barriers are initialised and armed before any peer may signal them

    const bool pingpong = prm.pingpong != 0 && nval - nv0 > 0;   // both tiles populated: alternate on the LSU phases
    if (nvg > 0) {
    if (mm) cmm_z_statistics(prm, M, g, gtid);
    if (pingpong && g == 1) CT_LSU_RELEASE(1);      // tile 0 goes first
#pragma unroll 1
    for (int t = 0; t < H; ++t) {
        const bool dbg_step = prm.dbg != nullptr && blockIdx.x == 0 && t == H / 2;
        const uint32_t par = (uint32_t)(t & 1);
        CL_TMARK(0);
        // per-step noise (only when the caller pre-drew [H, N, .] tables): issue the loads early
        if (pol.zstride != 0 && roleA && pol.has_density) zA = __ldg(pol.z + (size_t)t * pol.zstride + (size_t)a_n * U + a_u);
        if (dyn.zstride != 0 && roleB && dyn.has_density) zB = __ldg(dyn.z + (size_t)t * dyn.zstride + (size_t)b_n * D + b_d);
        if (mm && roleB) {      // z row of the matching step: z_mm[(t + n) mod N] (rollout.py:53-59); read after >= 4 tile barriers
            int r = t + prm.n_off + b_n;            // global particle index (the particles of all ranks are matched together)
            r -= (r / prm.n_global) * prm.n_global;
            M.zrow[b_p * SD + b_d] = __ldg(prm.z_mm + (size_t)r * D + b_d);
        }

        // ================= policy =================
        ct_net_forward<C, TK>(prm, pol, Rp, smem, xpol, act, red, g, gtid, mbox_pol_saddr, bar_pol, wstride, dbg_step, 1, pingpong);
        if (roleA) {
            // ---- Gaussian action sample + tanh squash (densities.py:95-119, core.py:243) ----
            mbar_wait(&xbar[g][0], par);
            CL_TMARK(20);
            if (gtid == 0) mbar_expect_tx(&xbar[g][0], bytes_pol);        // arm the next phase
            const float mu = a_nbm + ct_gather<C>(mbox_pol, a_p, a_u);
            float uu = mu, ls = 0.f;
            if (pol.has_density) {
                ls = a_nbl + ct_gather<C>(mbox_pol, a_p, U + a_u);
                uu += zA * ct_exp_clamped_logstd(ls, pol.lmax, elmax_pol);
            }
            const float a = a_sc * tanhf(uu) + a_bi;
            xdyn[(D + a_u) * CL_TS + a_p] = (a - a_mx) * a_isx;       // core.py:269,177
            if (a_own) {
                *act_ptr = a;
                rawp_ptr[0] = mu;
                if (pol.has_density) rawp_ptr[U] = ls;
            }
        }
        act_ptr += act_step;
        rawp_ptr += rawp_step;
        CL_TMARK(4);
        CT_SYNC(g);

        // ================= dynamics =================
        ct_net_forward<C, TK>(prm, dyn, Rd, smem, xdyn, act, red, g, gtid, mbox_dyn_saddr, bar_dyn, wstride, dbg_step, 5, pingpong);
        if (roleB) {
            // ---- Gaussian state sample, s' = s + delta (densities.py:100-119, core.py:293,298) ----
            mbar_wait(&xbar[g][1], par);
            CL_TMARK(21);
            if (gtid == 64) mbar_expect_tx(&xbar[g][1], bytes_dyn);
            const float mu = b_nbm + ct_gather<C>(mbox_dyn, b_p, b_d);
            float delta, ls = 0.f;
            if (dyn.has_density) {
                ls = b_nbl + ct_gather<C>(mbox_dyn, b_p, D + b_d);
                // exp(clamped log-std + log Sy) = Sy * exp(clamped log-std)   (densities.py:105)
                delta = (mu * b_sy + b_my) + zB * (b_sy * ct_exp_clamped_logstd(ls, dyn.lmax, elmax_dyn));
            } else {
                delta = mu * b_sy + b_my;
            }
            s_reg += delta;
            if (!mm) {
                xpol[b_d * CL_TS + b_p] = s_reg;
                xdyn[b_d * CL_TS + b_p] = (s_reg - b_mx) * b_isx;
            }
            if (b_own) {
                if (!mm) *st_ptr = s_reg;
                rawd_ptr[0] = mu;
                if (dyn.has_density) rawd_ptr[D] = ls;
            }
        }
        if (mm) {
            // ---- moment matching: every tile of the grid exchanges its pre-matching particles (rollout.py:121-128) ----
            cmm_forward(prm, M, g, gtid, rank, t, mm_base + (unsigned)(t + 1) * mm_ncl, nvg, roleB, b_p, b_d, b_n, b_own, s_reg, mm_leader,
                        dbg_step ? prm.dbg + 24 * 8 : nullptr);
            if (roleB) {
                xpol[b_d * CL_TS + b_p] = s_reg;
                xdyn[b_d * CL_TS + b_p] = (s_reg - b_mx) * b_isx;
                if (b_own) *st_ptr = s_reg;
            }
        }
        st_ptr += st_step;
        rawd_ptr += rawd_step;
        CL_TMARK(8);
        CT_SYNC(g);
    }
    // ---- rewards r_t = scale*exp(-0.5*(d^T Q d + a^T R a)) + offset on (s_{t+1}, a_t) for every step
    //      (envs/cartpole/env.py:62-86).  Nothing in the recurrence consumes them: evaluated here, off the serial
    //      chain, for the particles whose trajectory THIS group of THIS CTA wrote. ----
    for (int i = gtid; i < H * CL_TS; i += CL_GT) {
        const int tt = i / CL_TS, p = i - tt * CL_TS;
        if (p >= nvg || ((g * CL_TS + p) % C) != rank) continue;
        // with moment matching the reward sees the next state BEFORE the matching (models/core.py:293 runs inside dynamics())
        const float *s1 = mm ? prm.s1pre + ((size_t)tt * N + n0g + p) * D : prm.states + ((size_t)(tt + 1) * N + n0g + p) * D;
        const float *a = prm.actions + ((size_t)tt * N + n0g + p) * U;
        float dl[PMB_MAX_REWARD_ROWS];
        for (int r = 0; r < prm.KR; ++r) {
            float acc = cst[C_C0 + r];
            for (int d = 0; d < D; ++d) acc = fmaf(cst[C_C + r * SD + d], s1[d], acc);
            dl[r] = acc;
        }
        float cost = 0.f;
        for (int r = 0; r < prm.KR; ++r) {
            float q = 0.f;
            for (int j = 0; j < prm.KR; ++j) q = fmaf(dl[j], cst[C_Q + j * SD + r], q);
            cost = fmaf(q, dl[r], cost);
        }
        for (int u = 0; u < U; ++u) {
            float q = 0.f;
            for (int v = 0; v < U; ++v) q = fmaf(a[v], cst[C_R + v * SD + u], q);
            cost = fmaf(q, a[u], cost);
        }
        prm.rewards[(size_t)tt * N + n0g + p] = prm.rew_scale * expf(-0.5f * cost) + prm.rew_offset;
    }
    } else if (mm) {
        // an idle tile still takes part in the per-step exchange (cluster barrier + CTA barrier)
        for (int t = 0; t < H; ++t) cmm_idle_step(prm, M, g, gtid, rank, t, mm_base + (unsigned)(t + 1) * mm_ncl);
    }
    // across GPUs the arrival counters are never reset: leave the count this launch ends at for the next one (every CTA
    // read the old value before its first exchange, and cluster 0 is past the last one)
    if (mm && prm.mm_base_next && blockIdx.x == 0 && tid == 0) *prm.mm_base_next = (unsigned long long)(mm_base + (unsigned)H * mm_ncl);
    cl_sync();          // no CTA leaves while a peer could still address its shared memory
}

static cudaError_t cluster_launch_cfg(const void *fn, int C, int smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    if (C > 8) e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
}

cudaError_t launch_cluster_fwd(const ClusterParams &prm, int nclusters, cudaStream_t stream) {
    const int smem_bytes = prm.smem_floats * 4;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = prm.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(nclusters * prm.C);
    cfg.blockDim = dim3(CL_NT);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (prm.mm_states) {        // the per-step exchange spins on a global counter: every cluster must be resident
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.numAttrs = 2;
    }
    cudaError_t e;
    const int tkm = max(prm.pol.tK, prm.dyn.tK);
    const int tk = tkm <= 6 ? 6 : tkm <= 8 ? 8 : 16;
#define PMB_CL_FWD(CC, TT)                                                                                      \
    if (prm.C == CC && tk == TT && !prm.mm_states) {                                                            \
        if ((e = cluster_launch_cfg((const void *)cluster_fwd_kernel<CC, TT>, CC, smem_bytes)) != cudaSuccess)  \
            return e;                                                                                           \
        return cudaLaunchKernelEx(&cfg, cluster_fwd_kernel<CC, TT>, prm);                                       \
    }                                                                                                           \
    if (prm.C == CC && tk == TT && prm.mm_states) {                                                             \
        if ((e = cluster_launch_cfg((const void *)cluster_fwd_kernel<CC, TT, true>, CC, smem_bytes)) != cudaSuccess) \
            return e;                                                                                           \
        return cudaLaunchKernelEx(&cfg, cluster_fwd_kernel<CC, TT, true>, prm);                                 \
    }
    PMB_CL_FWD(8, 6)
    PMB_CL_FWD(8, 8)
    PMB_CL_FWD(8, 16)
    PMB_CL_FWD(4, 6)
    PMB_CL_FWD(4, 8)
    PMB_CL_FWD(4, 16)
#undef PMB_CL_FWD
    return cudaErrorInvalidValue;
}

// co-resident clusters of the forward kernel (0 when the query is unavailable, e.g. no device)
int cluster_max_active(int C, int smem_bytes, bool fwd) {
    (void)fwd;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(C * 64);
    cfg.blockDim = dim3(CL_NT);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    const void *fn = C == 8 ? (const void *)cluster_fwd_kernel<8, 8> : (const void *)cluster_fwd_kernel<4, 8>;
    if (cluster_launch_cfg(fn, C, smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

}  // namespace pmb
