// Forward sweep, cluster-resident variant (see pmb_cluster.cuh): H steps of
//   policy MLP -> Gaussian action sample -> tanh squash -> dynamics MLP -> Gaussian state sample -> reward
// for PG particles per cluster of C CTAs, all weights resident in the cluster's shared memory.
// Replaces the loop body of utils.rollout (reference utils/rollout.py:93-163) with Policy.forward
// (models/core.py:221-248), DynamicsModel.forward (models/core.py:265-303), B/CDropout masks
// (models/modules.py:61,160), DiagGaussianDensity (models/densities.py:87-121) and the env reward
// (envs/cartpole/env.py:41-86 et al.).  Same workspace layout as the streaming sweep (pmb_rollout_fwd.cu),
// so the reverse sweep and the weight-gradient kernels of either variant can follow.
#include "pmb_cluster.cuh"
#include "pmb_host.h"

namespace pmb {

// Per-thread constants of a net pass, fixed for the whole horizon (registers).
//   thin layer : thread = column j = tid of hidden 0: its TK weights, bias, the 8 particle slots' mask / keep
//   epilogue of the wide layer: thread = (particle slot = warp, column = lane of this CTA's slice)
template <int TK>
struct FwdNetRegs {
    float tw[TK];
    float tb;
    float tmk[CL_PS];
    float wb, wmk;
    bool thin_on, thin_store, wide_on, wide_store, send_ok;
    float *sv_thin, *sv_wide;       // running global pointers of the stored activations (advance per step)
    size_t thin_step, wide_step;
    int tK, tcol;
    __device__ __forceinline__ void init(const ClusterParams &prm, const CNet &n, const float *smem, int rank, int n0,
                                         int nval) {
        const int tid = threadIdx.x, lane = tid & 31, p = tid >> 5;
        thin_on = tid < n.tW;
#pragma unroll
        for (int k = 0; k < TK; ++k) tw[k] = (thin_on && k < n.tK) ? smem[n.s_tw + k * n.tW + tid] : 0.f;
        tb = thin_on ? smem[n.s_tb + tid] : 0.f;
#pragma unroll
        for (int q = 0; q < CL_PS; ++q) tmk[q] = thin_on ? smem[n.s_tm + q * n.tW + tid] * n.tkeep_inv : 0.f;
        tK = n.tK;
        // the thin output is stored by warp = particle slot, lane = column of this CTA's share of the columns
        tcol = rank * n.tsl + lane;
        thin_store = p < nval && lane < n.tsl && tcol < n.tW;
        sv_thin = prm.ws + n.tsav_off + (size_t)(n0 + p) * n.tW + tcol;
        thin_step = (size_t)prm.N * n.tW;
        const int gc = rank * n.hs + lane;
        wide_on = lane < n.hs;
        wb = wide_on ? smem[n.s_wb + lane] : 0.f;
        wmk = wide_on ? smem[n.s_wm + p * n.hs + lane] * n.wkeep_inv : 0.f;
        wide_store = wide_on && p < nval && gc < n.wN;
        sv_wide = prm.ws + n.wsav_off + (size_t)(n0 + p) * n.wN + gc;
        wide_step = (size_t)prm.N * n.wN;
        send_ok = p < prm.PG;
    }
};

// One net pass up to and including the send of the output partials.
//   x : [TK][8] input tile (rows >= tK are zero); on return the partial sums of the raw outputs are on their way
//   to every CTA of the cluster.  Two CTA barriers.
template <int C, int TK>
__device__ __forceinline__ void cl_net_forward(const ClusterParams &prm, const CNet &n, FwdNetRegs<TK> &R, float *smem,
                                               const float *x, int nval, int rank, uint32_t inbox_saddr,
                                               uint32_t bar_saddr, uint32_t wstride, bool dbg_step, int mark0) {
    float *act = smem + prm.off_act, *red = smem + prm.off_red;
    // ---- thin: hidden 0 = relu(x W0^T + b0) * mask0 / keep0 (full width, every CTA) ----
    if (R.thin_on) {
        float2 acc[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) acc[h] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            if (k < R.tK) {         // uniform: rows >= tK are zero padding
                const float4 x0 = *reinterpret_cast<const float4 *>(x + k * CL_PS);
                const float4 x1 = *reinterpret_cast<const float4 *>(x + k * CL_PS + 4);
                acc[0] = cl_fma2(R.tw[k], make_float2(x0.x, x0.y), acc[0]);
                acc[1] = cl_fma2(R.tw[k], make_float2(x0.z, x0.w), acc[1]);
                acc[2] = cl_fma2(R.tw[k], make_float2(x1.x, x1.y), acc[2]);
                acc[3] = cl_fma2(R.tw[k], make_float2(x1.z, x1.w), acc[3]);
            }
        }
        float v[CL_PS];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            v[2 * h] = acc[h].x + R.tb;
            v[2 * h + 1] = acc[h].y + R.tb;
        }
#pragma unroll
        for (int p = 0; p < CL_PS; ++p) v[p] = fmaxf(v[p], 0.f) * R.tmk[p];
        cl_store_act(act, threadIdx.x, v);
    }
    CL_TMARK(mark0);
    __syncthreads();
    // hidden 0 is kept for the reverse sweep: one coalesced row segment per warp (= particle slot), off the tile
    if (R.thin_store) *R.sv_thin = cl_act_at(act, R.tcol, threadIdx.x >> 5);
    R.sv_thin += R.thin_step;
    // ---- wide: this CTA's columns of hidden 1, k-split over the warps ----
    cl_wide_accum2(smem + n.s_ww, n.tW, n.hs, act, red);
    CL_TMARK(mark0 + 1);
    __syncthreads();
    // ---- epilogue (warp = particle slot, lane = column) + narrow partial sums + exchange ----
    {
        float v = cl_wide_reduce(red);
        v = fmaxf(v + R.wb, 0.f) * R.wmk;           // idle lanes: wb = wmk = 0 and red holds zeros
        if (R.wide_store) *R.sv_wide = v;
        R.sv_wide += R.wide_step;
        cl_narrow_send_any<C>(v, smem + n.s_nwt, threadIdx.x >> 5, R.send_ok, n.nN, inbox_saddr, bar_saddr, rank, wstride);
    }
    CL_TMARK(mark0 + 2);
}

template <int C, int TK>
__global__ void __launch_bounds__(CL_NT, 1) cluster_fwd_kernel(const __grid_constant__ ClusterParams prm) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t xbar[2];          // [0] policy exchange, [1] dynamics exchange
    const int tid = threadIdx.x;
    const int rank = (int)cl_rank();
    const int PG = prm.PG;
    const int n0 = (int)cl_id_x() * PG;
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H;
    const int nval = min(PG, N - n0);
    const CNet &pol = prm.pol;
    const CNet &dyn = prm.dyn;

    for (int i = tid; i < prm.smem_floats; i += CL_NT) smem[i] = 0.f;
    __syncthreads();
    float *cst = smem + prm.off_cst;
    float *xpol = smem + prm.off_xa;       // [TK][8] policy input (raw state), rows >= D stay zero
    float *xdyn = smem + prm.off_xb;       // [TK][8] dynamics input (scaled state, scaled action)
    const float *inbox_pol = smem + prm.off_inbox;
    const float *inbox_dyn = inbox_pol + C * CL_INBOX;
    const uint32_t bytes_pol = (uint32_t)(C * PG * pol.nNp) * 4u, bytes_dyn = (uint32_t)(C * PG * dyn.nNp) * 4u;
    if (tid == 0) {
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        fence_mbar_init();
        mbar_expect_tx(&xbar[0], bytes_pol);
        mbar_expect_tx(&xbar[1], bytes_dyn);
    }
    load_constants(prm, cst);
    cl_load_net(prm, pol, smem, rank, n0, true);
    cl_load_net(prm, dyn, smem, rank, n0, true);
    __syncthreads();
    FwdNetRegs<TK> Rp, Rd;
    Rp.init(prm, pol, smem, rank, n0, nval);
    Rd.init(prm, dyn, smem, rank, n0, nval);

    // ---- thread roles for the per-particle stages (fixed for the whole horizon) ----
    const bool roleA = tid < CL_PS * U;                        // one (particle slot, action dim)
    const int a_p = roleA ? tid / U : 0, a_u = roleA ? tid - a_p * U : 0;
    const int a_n = min(n0 + a_p, N - 1);
    const bool a_own = roleA && a_p < nval && (a_p % C) == rank;
    const bool roleB = tid >= 128 && tid - 128 < CL_PS * D;    // one (particle slot, state dim)
    const int b_p = roleB ? (tid - 128) / D : 0, b_d = roleB ? (tid - 128) - b_p * D : 0;
    const int b_n = min(n0 + b_p, N - 1);
    const bool b_own = roleB && b_p < nval && (b_p % C) == rank;

    float s_reg = 0.f;            // role B: this thread's element of the current state
    float b_mx = 0.f, b_isx = 0.f, b_sy = 0.f, b_my = 0.f, b_nbm = 0.f, b_nbl = 0.f;
    float a_mx = 0.f, a_isx = 0.f, a_sc = 0.f, a_bi = 0.f, a_nbm = 0.f, a_nbl = 0.f;
    if (roleB) {
        b_mx = prm.mx[b_d]; b_isx = prm.iSx[b_d]; b_sy = prm.Sy[b_d]; b_my = prm.my[b_d];
        b_nbm = smem[dyn.s_nb + b_d];
        if (dyn.has_density) b_nbl = smem[dyn.s_nb + D + b_d];
        s_reg = prm.x0[(size_t)b_n * D + b_d];
        xpol[b_d * CL_PS + b_p] = s_reg;
        xdyn[b_d * CL_PS + b_p] = (s_reg - b_mx) * b_isx;
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
    const uint32_t inbox_saddr = smem_u32(smem + prm.off_inbox);
    const uint32_t bar_pol = smem_u32(&xbar[0]), bar_dyn = smem_u32(&xbar[1]);
    const uint32_t wstride = cl_window_stride(bar_pol, C);
    // running global pointers of the role threads (advance per step)
    float *act_ptr = prm.actions + (size_t)a_n * U + a_u;
    float *rawp_ptr = prm.ws + pol.raw_off + (size_t)a_n * pol.nraw + a_u;
    float *st_ptr = prm.states + ((size_t)N + b_n) * D + b_d;
    float *rawd_ptr = prm.ws + dyn.raw_off + (size_t)b_n * dyn.nraw + b_d;
    const size_t act_step = (size_t)N * U, rawp_step = (size_t)N * pol.nraw, st_step = (size_t)N * D,
                 rawd_step = (size_t)N * dyn.nraw;

    __syncthreads();
    cl_sync();          // every CTA's barriers are initialised and armed before any peer may signal them

#pragma unroll 1
    for (int t = 0; t < H; ++t) {
        const bool dbg_step = prm.dbg != nullptr && blockIdx.x == 0 && t == H / 2;
        const uint32_t par = (uint32_t)(t & 1);
        CL_TMARK(0);
        // per-step noise (only when the caller pre-drew [H, N, .] tables): issue the loads early
        if (pol.zstride != 0 && roleA && pol.has_density) zA = __ldg(pol.z + (size_t)t * pol.zstride + (size_t)a_n * U + a_u);
        if (dyn.zstride != 0 && roleB && dyn.has_density) zB = __ldg(dyn.z + (size_t)t * dyn.zstride + (size_t)b_n * D + b_d);

        // ================= policy =================
        cl_net_forward<C, TK>(prm, pol, Rp, smem, xpol, nval, rank, inbox_saddr, bar_pol, wstride, dbg_step, 1);
        if (roleA) {
            // ---- Gaussian action sample + tanh squash (densities.py:95-119, core.py:243) ----
            mbar_wait(&xbar[0], par);
            if (tid == 0) mbar_expect_tx(&xbar[0], bytes_pol);        // arm the next phase
            const float mu = a_nbm + cl_gather2<C>(inbox_pol, a_p, a_u);
            float uu = mu, ls = 0.f;
            if (pol.has_density) {
                ls = a_nbl + cl_gather2<C>(inbox_pol, a_p, U + a_u);
                uu += zA * exp_clamped_logstd(ls, pol.lmax, elmax_pol);
            }
            const float a = a_sc * tanhf(uu) + a_bi;
            xdyn[(D + a_u) * CL_PS + a_p] = (a - a_mx) * a_isx;       // core.py:269,177
            if (a_own) {
                *act_ptr = a;
                rawp_ptr[0] = mu;
                if (pol.has_density) rawp_ptr[U] = ls;
            }
        }
        act_ptr += act_step;
        rawp_ptr += rawp_step;
        CL_TMARK(4);
        __syncthreads();

        // ================= dynamics =================
        cl_net_forward<C, TK>(prm, dyn, Rd, smem, xdyn, nval, rank, inbox_saddr + (uint32_t)(C * CL_INBOX) * 4u, bar_dyn,
                              wstride, dbg_step, 5);
        if (roleB) {
            // ---- Gaussian state sample, s' = s + delta (densities.py:100-119, core.py:293,298) ----
            mbar_wait(&xbar[1], par);
            if (tid == 128) mbar_expect_tx(&xbar[1], bytes_dyn);
            const float mu = b_nbm + cl_gather2<C>(inbox_dyn, b_p, b_d);
            float delta, ls = 0.f;
            if (dyn.has_density) {
                ls = b_nbl + cl_gather2<C>(inbox_dyn, b_p, D + b_d);
                // exp(clamped log-std + log Sy) = Sy * exp(clamped log-std)   (densities.py:105)
                delta = (mu * b_sy + b_my) + zB * (b_sy * exp_clamped_logstd(ls, dyn.lmax, elmax_dyn));
            } else {
                delta = mu * b_sy + b_my;
            }
            s_reg += delta;
            xpol[b_d * CL_PS + b_p] = s_reg;
            xdyn[b_d * CL_PS + b_p] = (s_reg - b_mx) * b_isx;
            if (b_own) {
                *st_ptr = s_reg;
                rawd_ptr[0] = mu;
                if (dyn.has_density) rawd_ptr[D] = ls;
            }
        }
        st_ptr += st_step;
        rawd_ptr += rawd_step;
        CL_TMARK(8);
        __syncthreads();
    }
    // ---- rewards r_t = scale*exp(-0.5*(d^T Q d + a^T R a)) + offset on (s_{t+1}, a_t) for every step
    //      (envs/cartpole/env.py:62-86).  Nothing in the recurrence consumes them: evaluated here, off the serial
    //      chain, for the particles whose trajectory THIS CTA wrote (slot p with p % C == rank). ----
    for (int i = tid; i < H * CL_PS; i += CL_NT) {
        const int tt = i / CL_PS, p = i - tt * CL_PS;
        if (p >= nval || (p % C) != rank) continue;
        const float *s1 = prm.states + ((size_t)(tt + 1) * N + n0 + p) * D;
        const float *a = prm.actions + ((size_t)tt * N + n0 + p) * U;
        float dl[PMB_MAX_REWARD_ROWS];
        for (int r = 0; r < prm.KR; ++r) {
            float acc = cst[C_C0 + r];
            for (int d = 0; d < D; ++d) acc = fmaf(cst[C_C + r * SD + d], s1[d], acc);
            dl[r] = acc;
        }
        float cost = 0.f;
        for (int r = 0; r < prm.KR; ++r) {
            float q = 0.f;
            for (int j = 0; j < prm.KR; ++j) q = fmaf(dl[j], cst[C_Q + j * 4 + r], q);
            cost = fmaf(q, dl[r], cost);
        }
        for (int u = 0; u < U; ++u) {
            float q = 0.f;
            for (int v = 0; v < U; ++v) q = fmaf(a[v], cst[C_R + v * SD + u], q);
            cost = fmaf(q, a[u], cost);
        }
        prm.rewards[(size_t)tt * N + n0 + p] = prm.rew_scale * expf(-0.5f * cost) + prm.rew_offset;
    }
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
    cudaLaunchAttribute attr[1];
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
    cudaError_t e;
    const int tk = max(prm.pol.tK, prm.dyn.tK) <= 8 ? 8 : 16;
#define PMB_CL_FWD(CC, TT)                                                                                      \
    if (prm.C == CC && tk == TT) {                                                                              \
        if ((e = cluster_launch_cfg((const void *)cluster_fwd_kernel<CC, TT>, CC, smem_bytes)) != cudaSuccess)  \
            return e;                                                                                           \
        return cudaLaunchKernelEx(&cfg, cluster_fwd_kernel<CC, TT>, prm);                                       \
    }
    PMB_CL_FWD(8, 8)
    PMB_CL_FWD(8, 16)
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
