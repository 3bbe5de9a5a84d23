// Forward sweep, wide cluster-resident variant (see pmb_cw.cuh): H steps of
//   policy MLP -> Gaussian action sample -> tanh squash -> dynamics MLP -> Gaussian state sample -> reward
// for up to 36 particles per cluster of 16 CTAs, all weights of two-hidden-layer nets up to 512 wide resident in the
// cluster's shared memory.  Replaces the loop body of utils.rollout (reference utils/rollout.py:93-163) with
// Policy.forward (models/core.py:221-248), DynamicsModel.forward (models/core.py:265-303), B/CDropout masks
// (models/modules.py:61,160), DiagGaussianDensity (models/densities.py:87-121) and the env reward
// (envs/cartpole/env.py:41-86 et al.).  Stores what pmb_cw_bwd.cu and the weight-gradient kernels consume: the
// policy's hidden activations and raw outputs, and the ReLU/dropout gates of both nets as bit words.
#include "pmb_cw.cuh"
#include "pmb_host.h"

namespace pmb {

// thin layer of one net for all 36 slots: hidden 0 = relu(x W0^T + b0) * mask0 / keep0, column j = tid, written to
// act[j][slot]; returns the gate bits (hidden 0 != 0) of column j and optionally stores the column to global memory
template <int TK>
__device__ __forceinline__ void cw_thin_forward(const CwThin<TK> &T, const float *__restrict__ x, float *__restrict__ act, int j,
                                                uint32_t &g0, uint32_t &g1, float *sv, size_t sv_stride, int nval) {
    g0 = g1 = 0u;
    float *row = act + j * CW_PS;
#pragma unroll
    for (int q = 0; q < CW_PS / 4; ++q) {
        float2 a01 = make_float2(0.f, 0.f), a23 = a01;
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 xv = *reinterpret_cast<const float4 *>(x + k * CW_PS + 4 * q);
            a01 = cl_fma2(T.w[k], make_float2(xv.x, xv.y), a01);
            a23 = cl_fma2(T.w[k], make_float2(xv.z, xv.w), a23);
        }
        float h[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int p = 4 * q + e;
            const uint32_t bit = p < 32 ? (T.m0 >> p) & 1u : (T.m1 >> (p - 32)) & 1u;
            const float v = bit ? fmaxf(h[e] + T.b, 0.f) * T.kinv : 0.f;
            h[e] = v;
            if (v != 0.f) {
                if (p < 32) g0 |= 1u << p;
                else g1 |= 1u << (p - 32);
            }
            if (sv != nullptr && p < nval) sv[(size_t)p * sv_stride] = v;
        }
        *reinterpret_cast<float4 *>(row + 4 * q) = make_float4(h[0], h[1], h[2], h[3]);
    }
}

// per-thread constants of the wide layer's epilogue: thread = (particle p = warp + 16 r for r < 3, column = lane)
struct CwWideFwd {
    float wb;
    float wmk[CW_OWN];              // mask / keep of (particle w + 16 r, column)
    bool store[CW_OWN], send[CW_OWN];
    float *sv[CW_OWN];              // policy: running global pointers of the stored hidden 1
    size_t sv_step;
    __device__ __forceinline__ void init(const ClusterParams &prm, const CNet &n, const float *smem, int rank, int n0,
                                         int nval, bool keep_act) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const int gc = rank * n.hs + lane;
        const bool on = lane < n.hs && gc < n.wN;
        wb = on ? smem[n.s_wb + lane] : 0.f;
        sv_step = (size_t)prm.N * n.wN;
#pragma unroll
        for (int r = 0; r < CW_OWN; ++r) {
            const int p = w + 16 * r;
            const int nn = min(n0 + p, prm.N - 1);
            wmk[r] = (on && p < CW_PS) ? (n.wm_off >= 0 ? __ldg(prm.ws + n.wm_off + (long long)nn * n.wN + gc) : 1.f) * n.wkeep_inv : 0.f;
            send[r] = p < nval;
            store[r] = keep_act && on && p < nval;
            sv[r] = prm.ws + n.wsav_off + (size_t)nn * n.wN + gc;
        }
    }
};

// One net pass up to and including the send of the output partials to the owners.  Three CTA barriers.
template <int TK>
__device__ __forceinline__ void cw_net_forward(const ClusterParams &prm, const CNet &n, const CwThin<TK> &T, CwWideFwd &W,
                                               float *smem, const float *x, float *act, int rank, int nval, int n0, int t,
                                               int net, bool keep_act, uint32_t mbox_saddr, uint32_t bar_saddr,
                                               uint32_t wstride, bool dbg_step, int mark0) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // ---- thin (full width, every CTA); the warp whose 32 columns carry this CTA's rank stores them ----
    if (T.on) {
        uint32_t g0, g1;
        const bool mine = w == rank;
        float *sv = (keep_act && mine) ? prm.ws + n.tsav_off + ((size_t)t * prm.N + n0) * n.tW + tid : nullptr;
        cw_thin_forward<TK>(T, x, act, tid, g0, g1, sv, (size_t)n.tW, nval);
        if (mine) {
            unsigned *g = prm.g1 + ((((size_t)t * prm.ncl + blockIdx.x / CW_C) * 2 + net) * 2) * CW_TW + tid;
            g[0] = g0;
            g[CW_TW] = g1;
        }
    }
    CW_MARK(mark0);
    __syncthreads();
    // ---- wide: this CTA's 32 columns of hidden 1, k-split over the 16 warps ----
    {
        float2 acc[9][2];
        cw_wide_accum(smem + n.s_ww, n.tW, act, acc);
        CW_MARK(mark0 + 1);
        __syncthreads();                 // every warp is done reading the activation tile
        cw_wide_park(act, acc);
    }
    CW_MARK(mark0 + 2);
    __syncthreads();
    // ---- epilogue (warp = particle, lane = column) + narrow partial sums -> owner ----
#pragma unroll
    for (int r = 0; r < CW_OWN; ++r) {
        const int p = w + 16 * r;
        if (p >= CW_PS) break;
        float v = cw_wide_reduce(act, p);
        v = fmaxf(v + W.wb, 0.f) * W.wmk[r];
        if (W.store[r]) *W.sv[r] = v;
        W.sv[r] += W.sv_step;
        const unsigned gate = __ballot_sync(0xffffffffu, v != 0.f);
        if (lane == 0 && W.send[r]) prm.g2[(((size_t)t * prm.N + n0 + p) * 2 + net) * CW_C + rank] = gate;
        cw_narrow_send(v, smem + n.s_nwt, n.nN, W.send[r], p, rank, mbox_saddr, bar_saddr, wstride);
    }
    CW_MARK(mark0 + 3);
}

template <int TKP, int TKD>
__global__ void __launch_bounds__(CW_NT, 1) cw_fwd_kernel(const __grid_constant__ ClusterParams prm) {
    extern __shared__ __align__(128) float smem[];
    // 0: policy partials at the owner, 1: actions everywhere, 2: dynamics partials at the owner, 3: states everywhere
    __shared__ __align__(8) uint64_t xbar[4];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int rank = (int)cl_rank();
    const int PG = prm.PG;
    const int n0 = (int)cl_id_x() * PG;
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H;
    const int nval = min(PG, N - n0);                  // particles of this cluster
    const CNet &pol = prm.pol;
    const CNet &dyn = prm.dyn;

    for (int i = tid; i < prm.smem_floats; i += CW_NT) smem[i] = 0.f;
    __syncthreads();
    float *cst = smem + prm.off_cst;
    float *xpol = smem + prm.off_xa;           // [16][36] policy input (raw state), rows >= D stay zero
    float *xdyn = smem + prm.off_xb;           // [16][36] dynamics input (scaled state, scaled action)
    float *act = smem + prm.off_act;           // [tW][36] hidden 0 / [16][36][32] partial sums of the wide layer
    const float *mb_pol = smem + prm.off_inbox;
    const float *mb_dyn = mb_pol + CW_MB;
    // owned particles: p = rank + 16 lp
    int nown = 0;
    for (int lp = 0; lp < CW_OWN; ++lp) nown += (rank + 16 * lp < nval) ? 1 : 0;
    const uint32_t bytes_own_pol = (uint32_t)(CW_C * nown * pol.nNp) * 4u, bytes_own_dyn = (uint32_t)(CW_C * nown * dyn.nNp) * 4u;
    const uint32_t bytes_act = (uint32_t)(nval * U) * 4u, bytes_st = (uint32_t)(nval * 2 * D) * 4u;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&xbar[i], 1);
        fence_mbar_init();
        if (nown) mbar_expect_tx(&xbar[0], bytes_own_pol);
        mbar_expect_tx(&xbar[1], bytes_act);
        if (nown) mbar_expect_tx(&xbar[2], bytes_own_dyn);
        mbar_expect_tx(&xbar[3], bytes_st);
    }
    load_constants(prm, cst);
    cw_load_net(prm, pol, smem, rank, true);
    cw_load_net(prm, dyn, smem, rank, true);
    __syncthreads();
    CwThin<TKP> Tp;
    CwThin<TKD> Td;
    Tp.init(prm, pol, tid, n0, true);
    Td.init(prm, dyn, tid, n0, true);
    CwWideFwd Wp, Wd;
    Wp.init(prm, pol, smem, rank, n0, nval, true);
    Wd.init(prm, dyn, smem, rank, n0, nval, false);

    // ---- owner role: warp lp < 3 runs the per-particle stages of particle p = rank + 16 lp; lane = (half, e):
    //      e = element (action / state dim) for the arithmetic, = destination rank for the broadcast ----
    const int e = lane & 15, half = lane >> 4;
    const int op = rank + 16 * w;                       // owned particle slot (warps 0..2)
    const bool owner = w < CW_OWN && op < nval;
    const int on_ = min(n0 + op, N - 1);
    const bool oa = owner && e < U, os = owner && e < D;
    float s_reg = 0.f;
    float b_mx = 0.f, b_isx = 0.f, b_sy = 0.f, b_my = 0.f, b_nbm = 0.f, b_nbl = 0.f;
    float a_mx = 0.f, a_isx = 0.f, a_sc = 0.f, a_bi = 0.f, a_nbm = 0.f, a_nbl = 0.f;
    if (os) {
        b_mx = prm.mx[e]; b_isx = prm.iSx[e]; b_sy = prm.Sy[e]; b_my = prm.my[e];
        b_nbm = smem[dyn.s_nb + e];
        if (dyn.has_density) b_nbl = smem[dyn.s_nb + D + e];
        s_reg = prm.x0[(size_t)on_ * D + e];
        if (half == 0) prm.states[(size_t)on_ * D + e] = s_reg;
    }
    if (oa) {
        a_mx = prm.mx[D + e]; a_isx = prm.iSx[D + e]; a_sc = prm.act_scale[e]; a_bi = prm.act_bias[e];
        a_nbm = smem[pol.s_nb + e];
        if (pol.has_density) a_nbl = smem[pol.s_nb + U + e];
    }
    const float elmax_pol = expf(pol.lmax), elmax_dyn = expf(dyn.lmax);
    float zA = 0.f, zB = 0.f;
    if (oa && pol.has_density) zA = __ldg(pol.z + (size_t)on_ * U + e);
    if (os && dyn.has_density) zB = __ldg(dyn.z + (size_t)on_ * D + e);
    // every CTA fills its own copy of the initial x tiles
    for (int i = tid; i < CW_PS * D; i += CW_NT) {
        const int p = i / D, d = i - p * D;
        const float s = prm.x0[(size_t)min(n0 + p, N - 1) * D + d];
        xpol[d * CW_PS + p] = s;
        xdyn[d * CW_PS + p] = (s - prm.mx[d]) * prm.iSx[d];
    }
    const uint32_t mbp_saddr = smem_u32(mb_pol), mbd_saddr = smem_u32(mb_dyn);
    const uint32_t bar0 = smem_u32(&xbar[0]), bar1 = smem_u32(&xbar[1]), bar2 = smem_u32(&xbar[2]), bar3 = smem_u32(&xbar[3]);
    const uint32_t wstride = cl_window_stride(bar0, CW_C);
    const uint32_t xpol_saddr = smem_u32(xpol), xdyn_saddr = smem_u32(xdyn);
    const uint32_t dst_off = (uint32_t)e * wstride;     // this lane's destination CTA for the broadcasts
    const uint32_t xpol0 = cl_mapa(xpol_saddr, 0), xdyn0 = cl_mapa(xdyn_saddr, 0);
    const uint32_t bar1_0 = cl_mapa(bar1, 0), bar3_0 = cl_mapa(bar3, 0);

    __syncthreads();
    cl_sync();          // every CTA's barriers are initialised and armed before any peer may signal them

#pragma unroll 1
    for (int t = 0; t < H; ++t) {
        const uint32_t par = (uint32_t)(t & 1);
        const bool dbg_step = prm.dbg != nullptr && blockIdx.x == 0 && t == H / 2;
        CW_MARK(0);
        if (pol.zstride != 0 && oa && pol.has_density) zA = __ldg(pol.z + (size_t)t * pol.zstride + (size_t)on_ * U + e);
        if (dyn.zstride != 0 && os && dyn.has_density) zB = __ldg(dyn.z + (size_t)t * dyn.zstride + (size_t)on_ * D + e);

        // ================= policy =================
        cw_net_forward<TKP>(prm, pol, Tp, Wp, smem, xpol, act, rank, nval, n0, t, 0, true, mbp_saddr, bar0, wstride, dbg_step, 1);
        if (owner) {
            // ---- Gaussian action sample + tanh squash (densities.py:95-119, core.py:243), once per particle ----
            mbar_wait(&xbar[0], par);
            CW_MARK(5);
            float xs = 0.f;
            if (e < U) {
                const float mu = a_nbm + cw_gather(mb_pol, w, e);
                float uu = mu, ls = 0.f;
                if (pol.has_density) {
                    ls = a_nbl + cw_gather(mb_pol, w, U + e);
                    uu += zA * ct_exp_clamped_logstd(ls, pol.lmax, elmax_pol);
                }
                const float a = a_sc * tanhf(uu) + a_bi;
                xs = (a - a_mx) * a_isx;                                  // core.py:269,177
                if (half == 0) {
                    prm.actions[((size_t)t * N + on_) * U + e] = a;
                    float *rp = prm.ws + pol.raw_off + ((size_t)t * N + on_) * pol.nraw;
                    rp[e] = mu;
                    if (pol.has_density) rp[U + e] = ls;
                }
            }
            __syncwarp();
            if (w == 0 && lane == 0) mbar_expect_tx(&xbar[0], bytes_own_pol);       // arm the next phase (warp 0 owns whenever any warp does)
            for (int u = 0; u < U; ++u) {
                const float v = __shfl_sync(0xffffffffu, xs, u);
                if ((u & 1) == half)
                    cw_st_async_f32(xdyn0 + dst_off + (uint32_t)(((D + u) * CW_PS + op) * 4), v, bar1_0 + dst_off);
            }
        }
        CW_MARK(6);
        mbar_wait(&xbar[1], par);
        CW_MARK(7);
        if (tid == 0) mbar_expect_tx(&xbar[1], bytes_act);
        __syncthreads();

        // ================= dynamics =================
        cw_net_forward<TKD>(prm, dyn, Td, Wd, smem, xdyn, act, rank, nval, n0, t, 1, false, mbd_saddr, bar2, wstride, dbg_step, 8);
        if (owner) {
            // ---- Gaussian state sample, s' = s + delta (densities.py:100-119, core.py:293,298) ----
            mbar_wait(&xbar[2], par);
            CW_MARK(12);
            float xs = 0.f;
            if (e < D) {
                const float mu = b_nbm + cw_gather(mb_dyn, w, e);
                float delta, ls = 0.f;
                if (dyn.has_density) {
                    ls = b_nbl + cw_gather(mb_dyn, w, D + e);
                    // exp(clamped log-std + log Sy) = Sy * exp(clamped log-std)   (densities.py:105)
                    delta = (mu * b_sy + b_my) + zB * (b_sy * ct_exp_clamped_logstd(ls, dyn.lmax, elmax_dyn));
                } else {
                    delta = mu * b_sy + b_my;
                }
                s_reg += delta;
                xs = (s_reg - b_mx) * b_isx;
                if (half == 0) {
                    prm.states[((size_t)(t + 1) * N + on_) * D + e] = s_reg;
                    float *rp = prm.ws + dyn.raw_off + ((size_t)t * N + on_) * dyn.nraw;
                    rp[e] = mu;
                    if (dyn.has_density) rp[D + e] = ls;
                }
            }
            __syncwarp();
            if (w == 0 && lane == 0) mbar_expect_tx(&xbar[2], bytes_own_dyn);
            for (int d = 0; d < D; ++d) {
                const float vr = __shfl_sync(0xffffffffu, s_reg, d), vs = __shfl_sync(0xffffffffu, xs, d);
                if (half == 0) cw_st_async_f32(xpol0 + dst_off + (uint32_t)((d * CW_PS + op) * 4), vr, bar3_0 + dst_off);
                else cw_st_async_f32(xdyn0 + dst_off + (uint32_t)((d * CW_PS + op) * 4), vs, bar3_0 + dst_off);
            }
        }
        CW_MARK(13);
        mbar_wait(&xbar[3], par);
        CW_MARK(14);
        if (tid == 0) mbar_expect_tx(&xbar[3], bytes_st);
        __syncthreads();
    }
    // ---- rewards r_t = scale*exp(-0.5*(d^T Q d + a^T R a)) + offset on (s_{t+1}, a_t) for every step
    //      (envs/cartpole/env.py:62-86).  Nothing in the recurrence consumes them: evaluated here, off the serial
    //      chain, for the particles whose trajectory THIS CTA wrote. ----
    for (int i = tid; i < H * CW_OWN; i += CW_NT) {
        const int tt = i / CW_OWN, lp = i - tt * CW_OWN;
        const int p = rank + 16 * lp;
        if (p >= nval) continue;
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
            for (int j = 0; j < prm.KR; ++j) q = fmaf(dl[j], cst[C_Q + j * SD + r], q);
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

static void cw_cfg(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, int nclusters, int smem_bytes, cudaStream_t stream) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CW_C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(nclusters * CW_C);
    cfg.blockDim = dim3(CW_NT);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}
static cudaError_t cw_attr(const void *fn, int smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
}

cudaError_t launch_cw_fwd(const ClusterParams &prm, int nclusters, cudaStream_t stream) {
    const int smem_bytes = prm.smem_floats * 4;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cw_cfg(cfg, attr, nclusters, smem_bytes, stream);
    cudaError_t e;
    const int tkp = prm.pol.tK, tkd = prm.dyn.tK;
#define PMB_CW_FWD(PP, DD)                                                                         \
    if (tkp <= PP && tkd <= DD) {                                                                  \
        if ((e = cw_attr((const void *)cw_fwd_kernel<PP, DD>, smem_bytes)) != cudaSuccess) return e; \
        return cudaLaunchKernelEx(&cfg, cw_fwd_kernel<PP, DD>, prm);                               \
    }
    PMB_CW_FWD(4, 6)
    PMB_CW_FWD(6, 6)
    PMB_CW_FWD(8, 8)
    PMB_CW_FWD(8, 12)
    PMB_CW_FWD(16, 16)
#undef PMB_CW_FWD
    return cudaErrorInvalidValue;
}

// co-resident 16-CTA clusters of the sweeps (0 when the query is unavailable, e.g. no device)
int cw_max_active() {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    const int smem_bytes = 220 * 1024;
    cw_cfg(cfg, attr, 64, smem_bytes, nullptr);
    const void *fn = (const void *)cw_fwd_kernel<4, 6>;
    int n = 0;
    if (cw_attr(fn, smem_bytes) != cudaSuccess || cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

}  // namespace pmb
