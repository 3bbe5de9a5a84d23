// Reverse sweep, cluster-resident variant (see pmb_cluster.cuh): back-propagation through time without
// recompute for PG particles per cluster of C CTAs, all (transposed) weights resident in the cluster's shared
// memory.  Consumes what either forward variant stored, walks t = H-1 .. 0 and produces dL/dx0 plus the
// per-layer output adjoints of the POLICY net for every (t, particle), which pmb_wgrad.cu contracts over the
// (H*N) axis afterwards.  Replaces loss.backward() through utils.rollout (reference
// algorithms/mc_pilco.py:197); the adjoint formulas are those of oracle/rollout_oracle.py::manual_backward.
//
// Everything that depends only on forward values (reward adjoint, density / tanh derivative factors) is computed
// by a fully parallel pre-pass (cluster_bwd_pre_kernel); the sweep fetches those factors, the direct cotangents
// and the ReLU/dropout gates ONE STEP AHEAD into registers, so the serial chain never waits on global memory
// and holds no transcendental arithmetic.
#include "pmb_cluster.cuh"
#include "pmb_cluster_mm.cuh"
#include "pmb_host.h"

namespace pmb {

constexpr int CBL = CL_PS * SD;      // one [8][SD] block of the per-particle scratch area (`misc`)

// Step-local adjoint factors for every (step, particle), computed in one fully parallel pass BEFORE the sweep so
// that the serial chain holds no transcendental or reward arithmetic:
//   pre[t][n] = [ RS(D) | FD(D) | RA(U) | TP(U) | FP(U) ]
//   RS = g_r * dr/ds'        reward adjoint on the next state:  w * C^T (Q+Q^T) delta,  w = -0.5 * g_r * (r - offset)
//   FD = d s'/d log_std      dynamics density:  z * exp(clamped log_std + log Sy) * sigmoid(lmax - log_std)
//   RA = g_a + g_r * dr/da   direct action cotangent + reward adjoint  w * (R+R^T) a
//   TP = d a/d u = scale * (1 - tanh(u)^2),   FP = d u/d log_std = z * exp(clamped log_std) * sigmoid(lmax - log_std)
__global__ void __launch_bounds__(256) cluster_bwd_pre_kernel(const __grid_constant__ ClusterParams prm) {
    const int N = prm.N, D = prm.D, U = prm.U, KR = prm.KR;
    const CNet &pol = prm.pol;
    const CNet &dyn = prm.dyn;
    const long long total = (long long)prm.H * N;
    const int PW = 2 * D + 3 * U;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / N), n = (int)(i - (long long)t * N);
        float *out = prm.pre + i * PW;
        // the reward sees the next state BEFORE moment matching (models/core.py:293 runs inside dynamics())
        const float *s1 = prm.s1pre ? prm.s1pre + (size_t)i * D : prm.states + ((size_t)(t + 1) * N + n) * D;
        const float *a = prm.actions + ((size_t)t * N + n) * U;
        const float r = __ldg(prm.rewards + i);
        const float gr = prm.g_rewards ? __ldg(prm.g_rewards + i) : 0.f;
        const float w = -0.5f * gr * (r - prm.rew_offset);       // g_r * d r / d cost, r - off = scale*exp(-cost)
        float dl[PMB_MAX_REWARD_ROWS], qd[PMB_MAX_REWARD_ROWS];
        for (int k = 0; k < KR; ++k) {
            float acc = __ldg(prm.rew_c0 + k);
            for (int d = 0; d < D; ++d) acc = fmaf(__ldg(prm.rew_C + k * D + d), __ldg(s1 + d), acc);
            dl[k] = acc;
        }
        for (int k = 0; k < KR; ++k) {
            float acc = 0.f;
            for (int j = 0; j < KR; ++j)
                acc = fmaf(__ldg(prm.rew_Q + k * KR + j) + __ldg(prm.rew_Q + j * KR + k), dl[j], acc);
            qd[k] = acc;
        }
        for (int d = 0; d < D; ++d) {
            float acc = 0.f;
            for (int k = 0; k < KR; ++k) acc = fmaf(qd[k], __ldg(prm.rew_C + k * D + d), acc);
            out[d] = w * acc;
            float fd = 0.f;
            if (dyn.has_density) {
                const float ls = __ldg(prm.ws + dyn.raw_off + (size_t)i * dyn.nraw + D + d);
                const float z = __ldg(dyn.z + (size_t)t * dyn.zstride + (size_t)n * D + d);
                const float lst = clamp_logstd(ls, dyn.lmax) + logf(__ldg(prm.Sy + d));
                fd = z * expf(lst) * sigmoid_f(dyn.lmax - ls);
            }
            out[D + d] = fd;
        }
        for (int u = 0; u < U; ++u) {
            float acc = 0.f;
            for (int v = 0; v < U; ++v)
                acc = fmaf(__ldg(prm.rew_R + u * U + v) + __ldg(prm.rew_R + v * U + u), __ldg(a + v), acc);
            const float ga = prm.g_actions ? __ldg(prm.g_actions + (size_t)i * U + u) : 0.f;
            out[2 * D + u] = ga + w * acc;
            const float *op = prm.ws + pol.raw_off + (size_t)i * pol.nraw;
            const float mu = __ldg(op + u);
            const float sc = __ldg(prm.act_scale + u);
            float tp, fp = 0.f;
            if (pol.has_density) {
                const float ls = __ldg(op + U + u);
                const float z = __ldg(pol.z + (size_t)t * pol.zstride + (size_t)n * U + u);
                const float el = expf(clamp_logstd(ls, pol.lmax));
                const float th = tanhf(mu + z * el);
                tp = sc * (1.f - th * th);
                fp = z * el * sigmoid_f(pol.lmax - ls);
            } else {
                const float th = tanhf(mu);
                tp = sc * (1.f - th * th);
            }
            out[2 * D + U + u] = tp;
            out[2 * D + 2 * U + u] = fp;
        }
    }
}

// Per-thread constants and running pointers of one net's adjoint pass (one particle tile = one warp group).
//   thin layer : thread gtid owns the columns gtid and gtid + 128 of hidden 1
//   epilogue of the wide layer: thread = (slot = warp of the group, column = lane of this CTA's slice)
struct BwdNetRegs {
    bool on0, on1, thin_store, wide_on, wide_store, send_ok;
    float wmk;                       // mask / keep of (slot, column) of hidden 0
    const float *tw;                 // smem: thin matrix, column gtid (second column: + 128)
    const float *tmk;                // smem: mask rows of hidden 1 for the tile's slots, column gtid
    float tkinv;
    const float *gt_ptr, *gw_ptr;    // global: stored activations that gate the adjoints (next step to fetch)
    float *dl_thin, *dl_wide;        // global: policy adjoints kept for the weight gradient (current step)
    size_t thin_step, wide_step;
    int tW, tK, tcol;
    uint32_t slot_off;
    __device__ __forceinline__ void init(const ClusterParams &prm, const CNet &n, const float *smem, int rank, int g,
                                         int gtid, int n0g, int nvg, bool store) {
        const int lane = gtid & 31, sl = gtid >> 5;
        const int ps = g * CL_TS + sl;
        const int H = prm.H, N = prm.N;
        tW = n.tW;
        tK = (n.tK + 3) & ~3;
        on0 = gtid < n.tW;
        on1 = gtid + CL_GT < n.tW;
        tw = smem + n.s_tw + gtid;
        tmk = smem + n.s_tm + (g * CL_TS) * n.tW + gtid;
        tkinv = n.tkeep_inv;
        tcol = rank * n.tsl + lane;
        thin_store = store && sl < nvg && lane < n.tsl && tcol < n.tW;
        thin_step = (size_t)N * n.tW;
        gt_ptr = prm.ws + n.tsav_off + ((size_t)(H - 1) * N + n0g) * n.tW + gtid;
        dl_thin = store ? prm.ws + n.tdel_off + ((size_t)(H - 1) * N + n0g + sl) * n.tW + tcol : nullptr;
        const int gc = rank * n.hs + lane;
        wide_on = lane < n.hs && gc < n.wN;
        wmk = wide_on ? smem[n.s_wm + ps * n.hs + lane] * n.wkeep_inv : 0.f;
        wide_store = store && wide_on && sl < nvg;
        wide_step = (size_t)N * n.wN;
        const int np = min(n0g + sl, N - 1);
        gw_ptr = prm.ws + n.wsav_off + ((size_t)(H - 1) * N + np) * n.wN + gc;
        dl_wide = store ? prm.ws + n.wdel_off + ((size_t)(H - 1) * N + n0g + sl) * n.wN + gc : nullptr;
        send_ok = sl < nvg;
        slot_off = (uint32_t)(rank * CL_MBOX + sl * CL_NO) * 4u;
    }
};

// Adjoint pass of one tile through one net up to and including the send of the input-adjoint partials.
//   x   : [tK][4] adjoint of the net's raw outputs
//   gtf : [column 0/1][slot] (stored activation of hidden 1 != 0) ? mask / keep : 0
//   gwf : same for hidden 0 at (slot = warp, column = lane of this CTA's slice)
// y = relu(pre) * mask / keep  =>  dpre = (dy / keep) * mask * [pre > 0];  y != 0 <=> pre > 0, mask != 0
template <int C, bool kStore>
__device__ __forceinline__ void ct_net_backward(const ClusterParams &prm, const CNet &n, BwdNetRegs &R, float *smem,
                                                const float *x, float *act, float *red, const float (&gtf)[2][CL_TS],
                                                float gwf, int g, int gtid, uint32_t mbox_saddr, uint32_t bar_saddr,
                                                uint32_t wstride, bool dbg_step, int mark0, bool pingpong) {
    if (pingpong) CT_LSU_ACQUIRE(g);
    // ---- thin: adjoint of hidden 1 = (dout W2) * gate; two columns per thread ----
    if (R.on0) {
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
        const float *wp = R.tw;
        const float *xp = x;
        const int c1 = R.on1 ? CL_GT : 0;        // threads without a second column re-read the first (result unused)
        // whole 4-row blocks: rows >= tK of the matrix and of x are zero padding
#pragma unroll 4
        for (int k = 0; k < R.tK; ++k) {
            const float w0 = wp[0], w1 = wp[c1];
            const float4 xv = *reinterpret_cast<const float4 *>(xp);
            wp += R.tW;
            xp += CL_TS;
            a0 = cl_fma2(w0, make_float2(xv.x, xv.y), a0);
            a1 = cl_fma2(w0, make_float2(xv.z, xv.w), a1);
            b0 = cl_fma2(w1, make_float2(xv.x, xv.y), b0);
            b1 = cl_fma2(w1, make_float2(xv.z, xv.w), b1);
        }
        *reinterpret_cast<float4 *>(act + gtid * CL_TS) =
            make_float4(a0.x * gtf[0][0], a0.y * gtf[0][1], a1.x * gtf[0][2], a1.y * gtf[0][3]);
        if (R.on1)
            *reinterpret_cast<float4 *>(act + (gtid + CL_GT) * CL_TS) =
                make_float4(b0.x * gtf[1][0], b0.y * gtf[1][1], b1.x * gtf[1][2], b1.y * gtf[1][3]);
    }
    CL_TMARK(mark0);
    CT_SYNC(g);
    if (kStore) {       // adjoint of hidden 1, kept for the weight gradient: one coalesced row segment per warp
        if (R.thin_store) *R.dl_thin = act[R.tcol * CL_TS + (gtid >> 5)];
        R.dl_thin -= R.thin_step;
    }
    // ---- wide: this CTA's columns of the adjoint of hidden 0, k-split 16 ways over the quarter-warps of the group ----
    ct_wide_accum(smem + n.s_ww, n.tW, n.hs, act, red, gtid);
    if (pingpong) CT_LSU_RELEASE(g);
    CL_TMARK(mark0 + 1);
    CT_SYNC(g);
    // ---- epilogue (warp = particle slot, lane = column) + partial sums of d(input) = delta_0 W_0 + exchange ----
    {
        const float v = ct_wide_reduce(red, gtid) * gwf;      // idle lanes: gwf = 0 and red holds zeros
        if (kStore) {
            if (R.wide_store) *R.dl_wide = v;
            R.dl_wide -= R.wide_step;
        }
        CL_TMARK(mark0 + 10);
        ct_narrow_send_any<C>(v, smem + n.s_nwt, R.send_ok, n.nN, mbox_saddr, R.slot_off, bar_saddr, wstride,
                              dbg_step ? prm.dbg + (mark0 + 8) * 8 + (threadIdx.x >> 5) : nullptr);
    }
    CL_TMARK(mark0 + 2);
}

template <int C, bool MM = false>
__global__ void __launch_bounds__(CL_NT, 1) cluster_bwd_kernel(const __grid_constant__ ClusterParams prm) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t xbar[2][2];       // [group][0 dynamics exchange, 1 policy exchange]
    const int tid = threadIdx.x;
    const int g = tid >> 7, gtid = tid & (CL_GT - 1);
    const int rank = (int)cl_rank();
    const int PG = prm.PG;
    const int n0 = (int)cl_id_x() * PG;
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H;
    const int nval = min(PG, N - n0);
    const int nv0 = (nval + 1) >> 1;
    const int n0g = n0 + (g ? nv0 : 0);
    const int nvg = g ? nval - nv0 : nv0;
    const CNet &pol = prm.pol;
    const CNet &dyn = prm.dyn;

    for (int i = tid; i < prm.smem_floats; i += CL_NT) smem[i] = 0.f;
    __syncthreads();
    const int tw_max = max(pol.tW, dyn.tW);
    float *xd = smem + prm.off_xa + g * (CL_NO * CL_TS);       // [2D][4]  adjoint of the dynamics net's raw outputs
    float *xp = smem + prm.off_xb + g * (CL_NO * CL_TS);       // [2U][4]  adjoint of the policy net's raw outputs
    float *act = smem + prm.off_act + g * (tw_max * CL_TS);
    float *red = smem + prm.off_red + g * (CL_KS * CL_TS * 32);
    float *gs = smem + prm.off_misc + g * (2 * CL_TS * SD), *gsp = gs + CL_TS * SD;   // [4][SD] carried / partial state adjoint
    const float *mbox_dyn = smem + prm.off_inbox + g * (2 * C * CL_MBOX);
    const float *mbox_pol = mbox_dyn + C * CL_MBOX;
    const uint32_t bytes_dyn = (uint32_t)(C * nvg * dyn.nNp) * 4u, bytes_pol = (uint32_t)(C * nvg * pol.nNp) * 4u;
    if (gtid == 0) {
        mbar_init(&xbar[g][0], 1);
        mbar_init(&xbar[g][1], 1);
        fence_mbar_init();
        mbar_expect_tx(&xbar[g][0], bytes_dyn);
        mbar_expect_tx(&xbar[g][1], bytes_pol);
    }
    cl_load_net(prm, dyn, smem, rank, n0, nv0, false);
    cl_load_net(prm, pol, smem, rank, n0, nv0, false);
    __syncthreads();
    BwdNetRegs Rd, Rp;
    Rd.init(prm, dyn, smem, rank, g, gtid, n0g, nvg, false);
    Rp.init(prm, pol, smem, rank, g, gtid, n0g, nvg, true);

    // ---- thread roles (fixed for the whole horizon) ----
    const bool roleB = gtid >= 64 && gtid - 64 < CL_TS * D;   // (particle slot, state dim)
    const int b_p = roleB ? (gtid - 64) / D : 0, b_d = roleB ? (gtid - 64) - b_p * D : 0;
    const int b_n = min(n0g + b_p, N - 1);
    const bool roleX = gtid < CL_TS * (D + U);                // (particle slot, dynamics-input dim)
    const int x_p = roleX ? gtid / (D + U) : 0, x_k = roleX ? gtid - x_p * (D + U) : 0;
    const bool x_own = roleX && x_p < nvg && ((g * CL_TS + x_p) % C) == rank;
    const float x_isx = roleX ? prm.iSx[x_k] : 0.f;
    const float b_sy = roleB ? prm.Sy[b_d] : 0.f;

    if (roleB) gs[b_p * SD + b_d] = prm.g_states ? __ldg(prm.g_states + ((size_t)H * N + b_n) * D + b_d) : 0.f;

    // step-local factors (cluster_bwd_pre_kernel) and gates, fetched one step ahead into registers
    const int PW = 2 * D + 3 * U;
    const int x_n = min(n0g + x_p, N - 1);
    const bool xact = roleX && x_k >= D;
    float nx_rs = 0.f, nx_fd = 0.f, nx_gs = 0.f;        // role B (slot, state dim)
    float nx_ra = 0.f, nx_tp = 0.f, nx_fp = 0.f;        // role X, action dims
    float pg_td[2][CL_TS], pg_tp[2][CL_TS];             // stored activations of the next step (thin columns)
    float pg_wd = 0.f, pg_wp = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int p = 0; p < CL_TS; ++p) pg_td[c][p] = pg_tp[c][p] = 0.f;
    const int last = max(nvg - 1, 0);                   // slots past the tile's last particle repeat it
    auto prefetch = [&](int tt) {
        if (roleB) {
            const float *q = prm.pre + ((size_t)tt * N + b_n) * PW;
            nx_rs = __ldg(q + b_d);
            nx_fd = __ldg(q + D + b_d);
            nx_gs = prm.g_states ? __ldg(prm.g_states + ((size_t)tt * N + b_n) * D + b_d) : 0.f;
        }
        if (xact) {
            const float *q = prm.pre + ((size_t)tt * N + x_n) * PW + 2 * D + (x_k - D);
            nx_ra = __ldg(q);
            nx_tp = __ldg(q + U);
            nx_fp = __ldg(q + 2 * U);
        }
        // stored activations that gate the adjoints (the running pointers stand on step tt)
#pragma unroll
        for (int p = 0; p < CL_TS; ++p) {
            const size_t row = (size_t)min(p, last);
            if (Rd.on0) pg_td[0][p] = __ldg(Rd.gt_ptr + row * Rd.tW);
            if (Rd.on1) pg_td[1][p] = __ldg(Rd.gt_ptr + row * Rd.tW + CL_GT);
            if (Rp.on0) pg_tp[0][p] = __ldg(Rp.gt_ptr + row * Rp.tW);
            if (Rp.on1) pg_tp[1][p] = __ldg(Rp.gt_ptr + row * Rp.tW + CL_GT);
        }
        if (Rd.wide_on) pg_wd = __ldg(Rd.gw_ptr);
        if (Rp.wide_on) pg_wp = __ldg(Rp.gw_ptr);
        Rd.gt_ptr -= Rd.thin_step;
        Rp.gt_ptr -= Rp.thin_step;
        Rd.gw_ptr -= Rd.wide_step;
        Rp.gw_ptr -= Rp.wide_step;
    };
    float gtf_dyn[2][CL_TS], gtf_pol[2][CL_TS];
    float gwf_dyn = 0.f, gwf_pol = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int p = 0; p < CL_TS; ++p) gtf_dyn[c][p] = gtf_pol[c][p] = 0.f;
    auto latch_gates = [&]() {
#pragma unroll
        for (int p = 0; p < CL_TS; ++p) {
            gtf_dyn[0][p] = (Rd.on0 && pg_td[0][p] != 0.f) ? Rd.tmk[p * Rd.tW] * Rd.tkinv : 0.f;
            gtf_dyn[1][p] = (Rd.on1 && pg_td[1][p] != 0.f) ? Rd.tmk[p * Rd.tW + CL_GT] * Rd.tkinv : 0.f;
            gtf_pol[0][p] = (Rp.on0 && pg_tp[0][p] != 0.f) ? Rp.tmk[p * Rp.tW] * Rp.tkinv : 0.f;
            gtf_pol[1][p] = (Rp.on1 && pg_tp[1][p] != 0.f) ? Rp.tmk[p * Rp.tW + CL_GT] * Rp.tkinv : 0.f;
        }
        gwf_dyn = pg_wd != 0.f ? Rd.wmk : 0.f;
        gwf_pol = pg_wp != 0.f ? Rp.wmk : 0.f;
    };

    const uint32_t mbox_dyn_saddr = smem_u32(mbox_dyn), mbox_pol_saddr = smem_u32(mbox_pol);
    const uint32_t bar_dyn = smem_u32(&xbar[g][0]), bar_pol = smem_u32(&xbar[g][1]);
    const uint32_t wstride = cl_window_stride(bar_dyn, C);
    float *odel_ptr = nullptr;          // role X (action dims): adjoint of the policy outputs of the current step
    if (xact) odel_ptr = prm.ws + pol.odel_off + ((size_t)(H - 1) * N + x_n) * pol.nraw + (x_k - D);
    const size_t odel_step = (size_t)N * pol.nraw;

    // moment matching of the states: scratch of this tile, arrivals per step, z statistics (constants of the launch)
    constexpr bool mm = MM;
    CMM M;
    M.carve(smem + prm.off_mm + g * CMM_FLOATS);
    const unsigned mm_ncl = mm ? cmm_arrivals(prm) : 0u;
    const unsigned mm_base = mm ? cmm_base(prm) : 0u;

    __syncthreads();
    cl_sync();                  // every CTA's barriers are initialised and armed before any peer may signal them

    const bool pingpong = prm.pingpong != 0 && nval - nv0 > 0;   // both tiles populated: alternate on the LSU phases
    if (nvg > 0) {
    if (mm) cmm_z_statistics(prm, M, g, gtid);
    if (pingpong && g == 1) CT_LSU_RELEASE(1);      // tile 0 goes first
    // ---- prologue: everything step H-1 needs ----
    float c_rs, c_fd, c_gs, c_ra, c_tp, c_fp;
    prefetch(H - 1);
    latch_gates();
    c_rs = nx_rs; c_fd = nx_fd; c_gs = nx_gs; c_ra = nx_ra; c_tp = nx_tp; c_fp = nx_fp;
    CT_SYNC(g);

#pragma unroll 1
    for (int t = H - 1, it = 0; t >= 0; --t, ++it) {
        const bool dbg_step = prm.dbg != nullptr && blockIdx.x == 0 && t == H / 2;
        const uint32_t par = (uint32_t)(it & 1);
        CL_TMARK(32);
        // ---- one step ahead: factors and gates of step t-1 ----
        if (t > 0) prefetch(t - 1);
        if (mm) {
            // ---- adjoint of the moment matching: cotangent of the matched s_{t+1} -> cotangent of the pre-matching
            //      particles (one exchange over all tiles of the grid, rollout.py:121-128) ----
            cmm_backward_prefetch(prm, M, g, gtid, t, roleB, b_p, b_d, b_n);
            cmm_backward(prm, M, g, gtid, rank, t, mm_base + (unsigned)(it + 1) * mm_ncl, nvg, gs, roleB, b_p, b_d);
        }
        // ---- total dL/ds_{t+1} (carried + reward) and the dynamics density adjoint:
        //      s' = s + mu*Sy + my + z*exp(lstd) ----
        if (roleB) {
            const float gg = gs[b_p * SD + b_d] + c_rs;
            gsp[b_p * SD + b_d] = gg;
            xd[b_d * CL_TS + b_p] = gg * b_sy;
            if (dyn.has_density) xd[(D + b_d) * CL_TS + b_p] = gg * c_fd;
        }
        CL_TMARK(33);
        CT_SYNC(g);
        // ================= dynamics net =================
        ct_net_backward<C, false>(prm, dyn, Rd, smem, xd, act, red, gtf_dyn, gwf_dyn, g, gtid, mbox_dyn_saddr, bar_dyn,
                                  wstride, dbg_step, 34, pingpong);
        if (roleX) {
            // ---- through the input scaler d[s;a] = dx * iSx, then (action dims) the tanh squash +
            //      policy density adjoint: a = scale*tanh(u)+bias, u = mu + z*exp(lstd) ----
            mbar_wait(&xbar[g][0], par);
            if (gtid == 0) mbar_expect_tx(&xbar[g][0], bytes_dyn);
            const float v = ct_gather<C>(mbox_dyn, x_p, x_k) * x_isx;
            if (x_k < D) {
                gsp[x_p * SD + x_k] += v;
            } else {
                const int u = x_k - D;
                const float du = (c_ra + v) * c_tp;
                if (x_own && prm.da_total) prm.da_total[((size_t)t * N + x_n) * U + u] = c_ra + v;
                xp[u * CL_TS + x_p] = du;
                float dls = 0.f;
                if (pol.has_density) {
                    dls = du * c_fp;
                    xp[(U + u) * CL_TS + x_p] = dls;
                }
                if (x_own) {
                    odel_ptr[0] = du;
                    if (pol.has_density) odel_ptr[U] = dls;
                }
            }
        }
        odel_ptr -= odel_step;
        CL_TMARK(37);
        CT_SYNC(g);
        // ================= policy net =================
        ct_net_backward<C, true>(prm, pol, Rp, smem, xp, act, red, gtf_pol, gwf_pol, g, gtid, mbox_pol_saddr, bar_pol,
                                 wstride, dbg_step, 38, pingpong);
        if (roleB) {
            // ---- dL/ds_t = carried + through dynamics input + through policy input + direct cotangent ----
            mbar_wait(&xbar[g][1], par);
            if (gtid == 64) mbar_expect_tx(&xbar[g][1], bytes_pol);
            gs[b_p * SD + b_d] = gsp[b_p * SD + b_d] + ct_gather<C>(mbox_pol, b_p, b_d) + c_gs;
        }
        if (t > 0) {
            latch_gates();
            c_rs = nx_rs; c_fd = nx_fd; c_gs = nx_gs; c_ra = nx_ra; c_tp = nx_tp; c_fp = nx_fp;
        }
        CL_TMARK(41);
        CT_SYNC(g);
    }
    if (prm.dx0 && roleB && b_p < nvg && ((g * CL_TS + b_p) % C) == rank)
        prm.dx0[(size_t)(n0g + b_p) * D + b_d] = gs[b_p * SD + b_d];
    } else if (mm) {
        // an idle tile still takes part in the per-step exchange (cluster barrier + CTA barrier)
        for (int it = 0; it < H; ++it) cmm_idle_step(prm, M, g, gtid, rank, H - 1 - it, mm_base + (unsigned)(it + 1) * mm_ncl);
    }
    if (mm && prm.mm_base_next && blockIdx.x == 0 && tid == 0) *prm.mm_base_next = (unsigned long long)(mm_base + (unsigned)H * mm_ncl);
    cl_sync();          // no CTA leaves while a peer could still address its shared memory
}

static cudaError_t cluster_launch_cfg_b(const void *fn, int C, int smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    if (C > 8) e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
}

cudaError_t launch_bwd_pre(const ClusterParams &prm, cudaStream_t stream) {
    const long long total = (long long)prm.H * prm.N;
    const int blocks = (int)min((total + 255) / 256, (long long)148 * 8);
    cluster_bwd_pre_kernel<<<blocks, 256, 0, stream>>>(prm);
    return cudaGetLastError();
}

cudaError_t launch_cluster_bwd(const ClusterParams &prm, int nclusters, cudaStream_t stream) {
    const int smem_bytes = prm.smem_floats * 4;
    {
        cudaError_t e0 = launch_bwd_pre(prm, stream);
        if (e0 != cudaSuccess) return e0;
    }
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
    switch (prm.C) {
        case 8:
            if (prm.mm_states) {
                if ((e = cluster_launch_cfg_b((const void *)cluster_bwd_kernel<8, true>, 8, smem_bytes)) != cudaSuccess) return e;
                return cudaLaunchKernelEx(&cfg, cluster_bwd_kernel<8, true>, prm);
            }
            if ((e = cluster_launch_cfg_b((const void *)cluster_bwd_kernel<8>, 8, smem_bytes)) != cudaSuccess) return e;
            return cudaLaunchKernelEx(&cfg, cluster_bwd_kernel<8>, prm);
        case 4:
            if (prm.mm_states) {
                if ((e = cluster_launch_cfg_b((const void *)cluster_bwd_kernel<4, true>, 4, smem_bytes)) != cudaSuccess) return e;
                return cudaLaunchKernelEx(&cfg, cluster_bwd_kernel<4, true>, prm);
            }
            if ((e = cluster_launch_cfg_b((const void *)cluster_bwd_kernel<4>, 4, smem_bytes)) != cudaSuccess) return e;
            return cudaLaunchKernelEx(&cfg, cluster_bwd_kernel<4>, prm);
        default:
            return cudaErrorInvalidValue;
    }
}

}  // namespace pmb
