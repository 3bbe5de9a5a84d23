// Reverse sweep, wide cluster-resident variant (see pmb_cw.cuh): back-propagation through time without recompute for
// up to 36 particles per cluster of 16 CTAs, all (transposed) weights resident in the cluster's shared memory.
// Consumes what pmb_cw_fwd.cu stored (gate bit words, raw outputs, the policy's activations), walks t = H-1 .. 0 and
// produces dL/dx0 plus the per-layer output adjoints of the POLICY net for every (t, particle), which pmb_wgrad.cu
// contracts over the (H*N) axis afterwards.  Replaces loss.backward() through utils.rollout (reference
// algorithms/mc_pilco.py:197); the adjoint formulas are those of oracle/rollout_oracle.py::manual_backward.  The
// step-local factors come from cluster_bwd_pre_kernel (pmb_cluster_bwd.cu).
#include "pmb_cw.cuh"
#include "pmb_host.h"

namespace pmb {

// thin layer of one net's adjoint pass for all 36 slots: adjoint of hidden 1 = (dout W2) * gate, column c = tid,
// written to act[c][slot]; gate bit (slot, c) = bit (c % hs) of word g2s[slot][c / hs]
template <int TK>
__device__ __forceinline__ void cw_thin_backward(const CwThin<TK> &T, const float *__restrict__ x, float *__restrict__ act, int c,
                                                 const unsigned *__restrict__ gw, int gsh, float *dl, size_t dl_stride, int nval) {
    float *row = act + c * CW_PS;
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
            const float v = ((gw[p * CW_C] >> gsh) & 1u) ? h[e] * T.kinv : 0.f;
            h[e] = v;
            if (dl != nullptr && p < nval) dl[(size_t)p * dl_stride] = v;
        }
        *reinterpret_cast<float4 *>(row + 4 * q) = make_float4(h[0], h[1], h[2], h[3]);
    }
}

// per-thread constants of the wide layer's epilogue: thread = (particle p = warp + 16 r for r < 3, column = lane)
struct CwWideBwd {
    float wmk[CW_OWN];              // mask / keep of (particle, column) of hidden 0
    bool on, send[CW_OWN];
    int gc;
    __device__ __forceinline__ void init(const ClusterParams &prm, const CNet &n, int rank, int n0, int nval) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        gc = rank * n.hs + lane;
        on = lane < n.hs && gc < n.wN;
#pragma unroll
        for (int r = 0; r < CW_OWN; ++r) {
            const int p = w + 16 * r;
            const int nn = min(n0 + p, prm.N - 1);
            wmk[r] = (on && p < CW_PS) ? (n.wm_off >= 0 ? __ldg(prm.ws + n.wm_off + (long long)nn * n.wN + gc) : 1.f) * n.wkeep_inv : 0.f;
            send[r] = p < nval;
        }
    }
};

// Adjoint pass of one net up to and including the send of the input-adjoint partials to the owners.
//   x  : [TK][36] adjoint of the net's raw outputs
//   g2s: [36][16] gate words of hidden 1 (this step), gh0/gh1: gate words of hidden 0 for this thread's column
template <int TK, bool kStore>
__device__ __forceinline__ void cw_net_backward(const ClusterParams &prm, const CNet &n, const CwThin<TK> &T, const CwWideBwd &W,
                                                float *smem, const float *x, float *act, const unsigned *g2s, int gcr, int gsh,
                                                unsigned gh0, unsigned gh1, int rank, int nval, int n0, int t,
                                                uint32_t mbox_saddr, uint32_t bar_saddr, uint32_t wstride) {
    const int tid = threadIdx.x, w = tid >> 5;
    // ---- thin: adjoint of hidden 1 (full width, every CTA) ----
    if (T.on) {
        float *dl = (kStore && w == rank) ? prm.ws + n.tdel_off + ((size_t)t * prm.N + n0) * n.tW + tid : nullptr;
        cw_thin_backward<TK>(T, x, act, tid, g2s + gcr, gsh, dl, (size_t)n.tW, nval);
    }
    __syncthreads();
    // ---- wide: this CTA's 32 columns of the adjoint of hidden 0 ----
    {
        float2 acc[9][2];
        cw_wide_accum(smem + n.s_ww, n.tW, act, acc);
        __syncthreads();
        cw_wide_park(act, acc);
    }
    __syncthreads();
    // ---- epilogue + partial sums of d(input) = delta_0 W_0 -> owner ----
#pragma unroll
    for (int r = 0; r < CW_OWN; ++r) {
        const int p = w + 16 * r;
        if (p >= CW_PS) break;
        const unsigned gb = r < 2 ? (gh0 >> p) & 1u : (gh1 >> (p - 32)) & 1u;
        const float v = gb ? cw_wide_reduce(act, p) * W.wmk[r] : 0.f;
        if (kStore && W.on && W.send[r]) prm.ws[n.wdel_off + ((size_t)t * prm.N + n0 + p) * n.wN + W.gc] = v;
        cw_narrow_send(v, smem + n.s_nwt, n.nN, W.send[r], p, rank, mbox_saddr, bar_saddr, wstride);
    }
}

template <int TKP, int TKD>
__global__ void __launch_bounds__(CW_NT, 1) cw_bwd_kernel(const __grid_constant__ ClusterParams prm) {
    extern __shared__ __align__(128) float smem[];
    // 0: dynamics output adjoints everywhere, 1: dynamics input partials at the owner,
    // 2: policy output adjoints everywhere, 3: policy input partials at the owner
    __shared__ __align__(8) uint64_t xbar[4];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int rank = (int)cl_rank();
    const int cid = (int)cl_id_x();
    const int PG = prm.PG;
    const int n0 = cid * PG;
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H;
    const int nval = min(PG, N - n0);
    const CNet &pol = prm.pol;
    const CNet &dyn = prm.dyn;

    for (int i = tid; i < prm.smem_floats; i += CW_NT) smem[i] = 0.f;
    __syncthreads();
    float *xd = smem + prm.off_xa;            // [2D][36] adjoint of the dynamics net's raw outputs
    float *xp = smem + prm.off_xb;            // [2U][36] adjoint of the policy net's raw outputs
    float *act = smem + prm.off_act;
    const float *mb_dyn = smem + prm.off_inbox;
    const float *mb_pol = mb_dyn + CW_MB;
    unsigned *g2d = reinterpret_cast<unsigned *>(smem + prm.off_misc);     // [36][16] gate words of hidden 1, dynamics
    unsigned *g2p = g2d + CW_G2;                                           // ... policy
    int nown = 0;
    for (int lp = 0; lp < CW_OWN; ++lp) nown += (rank + 16 * lp < nval) ? 1 : 0;
    const int nod = dyn.tK, nop = pol.tK;     // raw outputs of the nets (rows of the x tiles that are sent)
    const uint32_t bytes_xd = (uint32_t)(nval * nod) * 4u, bytes_xp = (uint32_t)(nval * nop) * 4u;
    const uint32_t bytes_own_dyn = (uint32_t)(CW_C * nown * dyn.nNp) * 4u, bytes_own_pol = (uint32_t)(CW_C * nown * pol.nNp) * 4u;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&xbar[i], 1);
        fence_mbar_init();
        mbar_expect_tx(&xbar[0], bytes_xd);
        if (nown) mbar_expect_tx(&xbar[1], bytes_own_dyn);
        mbar_expect_tx(&xbar[2], bytes_xp);
        if (nown) mbar_expect_tx(&xbar[3], bytes_own_pol);
    }
    cw_load_net(prm, dyn, smem, rank, false);
    cw_load_net(prm, pol, smem, rank, false);
    __syncthreads();
    CwThin<TKP> Tp;
    CwThin<TKD> Td;
    Tp.init(prm, pol, tid, n0, false);
    Td.init(prm, dyn, tid, n0, false);
    CwWideBwd Wd, Wp;
    Wd.init(prm, dyn, rank, n0, nval);
    Wp.init(prm, pol, rank, n0, nval);
    // gate word of hidden 1 for column c = tid: word (slot, c / hs), bit c % hs
    const int gcr_d = min(tid / dyn.tsl, CW_C - 1), gsh_d = tid - gcr_d * dyn.tsl;
    const int gcr_p = min(tid / pol.tsl, CW_C - 1), gsh_p = tid - gcr_p * pol.tsl;

    // ---- owner role (see pmb_cw_fwd.cu): warp lp < 3, lane = (half, e) ----
    const int e = lane & 15, half = lane >> 4;
    const int op = rank + 16 * w;
    const bool owner = w < CW_OWN && op < nval;
    const int on_ = min(n0 + op, N - 1);
    const bool os = owner && e < D, oa = owner && e >= D && e < D + U, ox = owner && e < D + U;
    const int ua = e - D;
    const float x_isx = ox ? prm.iSx[e] : 0.f;
    const float b_sy = os ? prm.Sy[e] : 0.f;
    float gs = 0.f, gsp = 0.f;                  // carried / partial dL/ds of (owned particle, state dim e)
    if (os) gs = prm.g_states ? __ldg(prm.g_states + ((size_t)H * N + on_) * D + e) : 0.f;

    // step-local factors (cluster_bwd_pre_kernel) and gate words, fetched one step ahead into registers
    const int PW = 2 * D + 3 * U;
    float nx_rs = 0.f, nx_fd = 0.f, nx_gs = 0.f, nx_ra = 0.f, nx_tp = 0.f, nx_fp = 0.f;
    unsigned nx_g2[2][2] = {{0u, 0u}, {0u, 0u}};        // [net][tid, tid + 512]: words of the [36][16] gate tiles of hidden 1
    unsigned nx_h[2][2] = {{0u, 0u}, {0u, 0u}};         // [net][half]: gate words of hidden 0 for this thread's column
    const bool g2b = tid + CW_NT < CW_G2;
    const int g2n0 = min(n0 + (tid >> 4), N - 1), g2n1 = min(n0 + ((tid + CW_NT) >> 4), N - 1);
    auto prefetch = [&](int tt) {
        if (os) {
            const float *q = prm.pre + ((size_t)tt * N + on_) * PW;
            nx_rs = __ldg(q + e);
            nx_fd = __ldg(q + D + e);
            nx_gs = prm.g_states ? __ldg(prm.g_states + ((size_t)tt * N + on_) * D + e) : 0.f;
        }
        if (oa) {
            const float *q = prm.pre + ((size_t)tt * N + on_) * PW + 2 * D + ua;
            nx_ra = __ldg(q);
            nx_tp = __ldg(q + U);
            nx_fp = __ldg(q + 2 * U);
        }
#pragma unroll
        for (int net = 0; net < 2; ++net) {     // 0 policy, 1 dynamics, as the forward sweep wrote them
            nx_g2[net][0] = __ldg(prm.g2 + (((size_t)tt * N + g2n0) * 2 + net) * CW_C + (tid & 15));
            if (g2b) nx_g2[net][1] = __ldg(prm.g2 + (((size_t)tt * N + g2n1) * 2 + net) * CW_C + (tid & 15));
            const CwWideBwd &W = net ? Wd : Wp;
            if (W.on) {
                const unsigned *g = prm.g1 + ((((size_t)tt * prm.ncl + cid) * 2 + net) * 2) * CW_TW + W.gc;
                nx_h[net][0] = __ldg(g);
                nx_h[net][1] = __ldg(g + CW_TW);
            }
        }
    };
    float c_rs = 0.f, c_fd = 0.f, c_gs = 0.f, c_ra = 0.f, c_tp = 0.f, c_fp = 0.f;
    unsigned gh[2][2] = {{0u, 0u}, {0u, 0u}};
    auto latch = [&]() {
        c_rs = nx_rs; c_fd = nx_fd; c_gs = nx_gs; c_ra = nx_ra; c_tp = nx_tp; c_fp = nx_fp;
        g2p[tid] = nx_g2[0][0];
        g2d[tid] = nx_g2[1][0];
        if (g2b) {
            g2p[tid + CW_NT] = nx_g2[0][1];
            g2d[tid + CW_NT] = nx_g2[1][1];
        }
        gh[0][0] = nx_h[0][0]; gh[0][1] = nx_h[0][1]; gh[1][0] = nx_h[1][0]; gh[1][1] = nx_h[1][1];
    };

    const uint32_t mbd_saddr = smem_u32(mb_dyn), mbp_saddr = smem_u32(mb_pol);
    const uint32_t bar0 = smem_u32(&xbar[0]), bar1 = smem_u32(&xbar[1]), bar2 = smem_u32(&xbar[2]), bar3 = smem_u32(&xbar[3]);
    const uint32_t wstride = cl_window_stride(bar0, CW_C);
    const uint32_t dst_off = (uint32_t)e * wstride;     // this lane's destination CTA for the broadcasts
    const uint32_t xd0 = cl_mapa(smem_u32(xd), 0), xp0 = cl_mapa(smem_u32(xp), 0);
    const uint32_t bar0_0 = cl_mapa(bar0, 0), bar2_0 = cl_mapa(bar2, 0);

    // total dL/ds_{t+1} (carried + reward) and the dynamics density adjoint of step t, s' = s + mu*Sy + my + z*exp(lstd):
    // the owner forms the adjoint of the dynamics net's raw outputs and broadcasts the column to every CTA
    auto emit_xd = [&]() {
        const float gg = gs + c_rs;
        gsp = gg;
        const float va = gg * b_sy, vb = gg * c_fd;
        for (int d = 0; d < D; ++d) {
            const float a = __shfl_sync(0xffffffffu, va, d), b = __shfl_sync(0xffffffffu, vb, d);
            if (half == 0) cw_st_async_f32(xd0 + dst_off + (uint32_t)((d * CW_PS + op) * 4), a, bar0_0 + dst_off);
            else if (dyn.has_density) cw_st_async_f32(xd0 + dst_off + (uint32_t)(((D + d) * CW_PS + op) * 4), b, bar0_0 + dst_off);
        }
    };

    __syncthreads();
    cl_sync();                  // every CTA's barriers are initialised and armed before any peer may signal them

    prefetch(H - 1);
    latch();
    if (owner) emit_xd();

#pragma unroll 1
    for (int t = H - 1, it = 0; t >= 0; --t, ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        if (t > 0) prefetch(t - 1);
        mbar_wait(&xbar[0], par);
        if (tid == 0) mbar_expect_tx(&xbar[0], bytes_xd);
        __syncthreads();
        // ================= dynamics net =================
        cw_net_backward<TKD, false>(prm, dyn, Td, Wd, smem, xd, act, g2d, gcr_d, gsh_d, gh[1][0], gh[1][1], rank, nval, n0, t,
                                    mbd_saddr, bar1, wstride);
        if (owner) {
            // ---- through the input scaler d[s;a] = dx * iSx, then (action dims) the tanh squash + policy density
            //      adjoint: a = scale*tanh(u)+bias, u = mu + z*exp(lstd) ----
            mbar_wait(&xbar[1], par);
            const float v = ox ? cw_gather(mb_dyn, w, e) * x_isx : 0.f;
            float du = 0.f, dls = 0.f;
            if (os) gsp += v;
            if (oa) {
                const float tot = c_ra + v;
                du = tot * c_tp;
                if (pol.has_density) dls = du * c_fp;
                if (half == 0) {
                    if (prm.da_total) prm.da_total[((size_t)t * N + on_) * U + ua] = tot;
                    float *od = prm.ws + pol.odel_off + ((size_t)t * N + on_) * pol.nraw;
                    od[ua] = du;
                    if (pol.has_density) od[U + ua] = dls;
                }
            }
            __syncwarp();
            if (w == 0 && lane == 0) mbar_expect_tx(&xbar[1], bytes_own_dyn);
            for (int u = 0; u < U; ++u) {
                const float a = __shfl_sync(0xffffffffu, du, D + u), b = __shfl_sync(0xffffffffu, dls, D + u);
                if (half == 0) cw_st_async_f32(xp0 + dst_off + (uint32_t)((u * CW_PS + op) * 4), a, bar2_0 + dst_off);
                else if (pol.has_density) cw_st_async_f32(xp0 + dst_off + (uint32_t)(((U + u) * CW_PS + op) * 4), b, bar2_0 + dst_off);
            }
        }
        mbar_wait(&xbar[2], par);
        if (tid == 0) mbar_expect_tx(&xbar[2], bytes_xp);
        __syncthreads();
        // ================= policy net =================
        cw_net_backward<TKP, true>(prm, pol, Tp, Wp, smem, xp, act, g2p, gcr_p, gsh_p, gh[0][0], gh[0][1], rank, nval, n0, t,
                                   mbp_saddr, bar3, wstride);
        const float gs_direct = c_gs;      // direct cotangent of s_t (before the factors of step t-1 are latched)
        if (t > 0) latch();                // gates and factors of step t-1 (the tiles are read again only after the next CTA barrier)
        if (owner) {
            // ---- dL/ds_t = carried + through dynamics input + through policy input + direct cotangent ----
            mbar_wait(&xbar[3], par);
            const float v = os ? cw_gather(mb_pol, w, e) : 0.f;
            __syncwarp();
            if (w == 0 && lane == 0) mbar_expect_tx(&xbar[3], bytes_own_pol);
            gs = gsp + v + gs_direct;
            if (t > 0) emit_xd();
        }
    }
    if (prm.dx0 && os && half == 0) prm.dx0[(size_t)on_ * D + e] = gs;
    cl_sync();          // no CTA leaves while a peer could still address its shared memory
}

cudaError_t launch_cw_bwd(const ClusterParams &prm, int nclusters, cudaStream_t stream) {
    const int smem_bytes = prm.smem_floats * 4;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
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
    cudaError_t e;
    const int tkp = prm.pol.tK, tkd = prm.dyn.tK;
#define PMB_CW_BWD(PP, DD)                                                                                              \
    if (tkp <= PP && tkd <= DD) {                                                                                       \
        const void *fn = (const void *)cw_bwd_kernel<PP, DD>;                                                           \
        if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)) != cudaSuccess) return e; \
        if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess) return e;  \
        return cudaLaunchKernelEx(&cfg, cw_bwd_kernel<PP, DD>, prm);                                                    \
    }
    PMB_CW_BWD(2, 8)
    PMB_CW_BWD(2, 12)
    PMB_CW_BWD(4, 16)
    PMB_CW_BWD(8, 16)
    PMB_CW_BWD(16, 16)
#undef PMB_CW_BWD
    return cudaErrorInvalidValue;
}

}  // namespace pmb
