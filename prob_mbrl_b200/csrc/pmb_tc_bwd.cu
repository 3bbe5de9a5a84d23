// Reverse sweep, tensor-core cluster variant (see pmb_tc.cuh): back-propagation through time without recompute
// for a tile of up to 128 particles per cluster of 16 CTAs.  Consumes what any forward variant stored, walks
// t = H-1 .. 0 and produces dL/dx0 plus the per-layer output adjoints of the POLICY net for every (t, particle),
// which pmb_wgrad.cu contracts over the (H*N) axis afterwards.  The hidden x hidden adjoints
//   d(hidden l-1)[128 x W] = d(hidden l)[128 x K] . W_l[K x W]
// run on tcgen05 (3xTF32 split, fp32 accumulation in TMEM) exactly like the forward layers, with the transposed
// weight slices; the skinny first / last layers and every per-particle factor stay on the FP32 pipe.  Replaces
// loss.backward() through utils.rollout (reference algorithms/mc_pilco.py:197); the adjoint formulas are those of
// oracle/rollout_oracle.py::manual_backward / mm_backward.  Step-local factors that need transcendental or reward
// arithmetic come from the fully parallel pre-pass cluster_bwd_pre_kernel (pmb_cluster_bwd.cu).
#include "pmb_tc.cuh"
#include "pmb_tc_mm.cuh"
#include "pmb_host.h"

namespace pmb {

constexpr int TC_XO = TC_NOUT + 1;      // row stride of the output-adjoint tile (thread-per-row, conflict-free)

// gate of (particle n, columns c0 .. c0+HW) of hidden layer l at step t:
//   y = relu(pre) * mask / keep  =>  dpre = dy * (mask / keep) * [pre > 0];   y != 0 <=> pre > 0 and mask != 0
// The stored activation and the mask are ISSUED early (tc_gate_issue) and turned into the gate factor where the
// adjoint is ready (tc_gate_apply), so that their L2 latency hides behind the layer in between.
template <int HW>
__device__ __forceinline__ void tc_gate_issue(const TcParams &prm, const TcNet &n, int l, int t, int nld, int c0,
                                              float (&sv)[HW], float (&mk)[HW]) {
    const int npad = n.npad[l];
    const float *ps = prm.ws + n.saved_off[l] + ((size_t)t * prm.N + nld) * npad + c0;
    const float *pm = n.mask_off[l] >= 0 ? prm.ws + n.mask_off[l] + (long long)nld * npad + c0 : nullptr;
#pragma unroll
    for (int j = 0; j < HW; j += 4) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), m = make_float4(1.f, 1.f, 1.f, 1.f);
        if (c0 + j < npad) {
            s = __ldcg(reinterpret_cast<const float4 *>(ps + j));
            if (pm) m = __ldg(reinterpret_cast<const float4 *>(pm + j));
        }
        sv[j] = s.x; sv[j + 1] = s.y; sv[j + 2] = s.z; sv[j + 3] = s.w;
        mk[j] = m.x; mk[j + 1] = m.y; mk[j + 2] = m.z; mk[j + 3] = m.w;
    }
}
template <int HW>
__device__ __forceinline__ void tc_gate_apply(float (&h)[HW], const float (&sv)[HW], const float (&mk)[HW], float ki) {
#pragma unroll
    for (int j = 0; j < HW; ++j) h[j] = sv[j] != 0.f ? h[j] * (mk[j] * ki) : 0.f;
}

template <int HW>
__global__ void __launch_bounds__(TC_NTL, 1) tc_bwd_kernel(const __grid_constant__ TcParams prm) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) TcBars bars;
    __shared__ __align__(16) TcWItem sched[TC_MAXITEMS];
    __shared__ uint32_t tmem_base_s, sched_n;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool compute = tid < TC_NT;
    const int rank = (int)tc_rank(), tile = (int)tc_cluster_id();
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H, ns = prm.ns;
    constexpr int C = TC_C;
    const int n0 = tile * prm.TP;
    const int nval = min(prm.TP, N - n0);
    const int p = 32 * (warp & 3) + lane, half = (warp >> 2) & 1;
    const bool valid = compute && p < nval;
    const int n = n0 + min(p, nval - 1);
    const int c0 = rank * ns + half * HW;
    const bool owner = valid && (p % C) == rank;
    const TcNet &pol = prm.pol;
    const TcNet &dyn = prm.dyn;

    for (int i = tid; i < prm.smem_floats; i += TC_NTL) smem[i] = 0.f;
    if (tid == 0) {
        for (int s = 0; s < TC_NSW; ++s) {
            mbar_init(&bars.w_full[s], 1);
            mbar_init(&bars.w_empty[s], TC_NDRV);
        }
        for (int s = 0; s < TC_NSX; ++s) {
            mbar_init(&bars.x_full[s], 1);
            mbar_init(&bars.x_empty[s], TC_NT / 32);
        }
        for (int s = 0; s < TC_NSA; ++s) {
            mbar_init(&bars.a_full[s], TC_NT / 32);
            mbar_init(&bars.a_empty[s], TC_NDRV);
        }
        mbar_init(&bars.done, TC_NDRV);
        fence_mbar_init();
    }
    if (tid == TC_NT) {
        const TcNet *const order[2] = {&dyn, &pol};
        sched_n = tc_build_schedule(prm, order, true, rank, sched);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t tmem_rd = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(half * HW);

    float *cst = smem + prm.off_cst;
    float *xo = smem + prm.off_xin;           // [128][TC_XO]  adjoint of the raw outputs of the net being walked
    float *gs = smem + prm.off_st;            // [128][TC_SDP] carried dL/ds_{t+1}; second tile: partial dL/ds_t
    float *gsp = gs + TC_M * TC_SDP;
    float *aux = smem + prm.off_aux;          // [2][nop][128]
    float *pre = smem + prm.off_z;            // [128][PW] step-local adjoint factors of the tile (cluster_bwd_pre_kernel)
    float *mmscr = smem + prm.off_mm;
    float *ring = smem + prm.off_ring;        // weight stages
    float *xring = smem + prm.off_xring;      // image stages
    const int PW = 2 * D + 3 * U;
    if (compute) {
        load_constants(prm, cst);
        // same resident slices as the forward sweep: the output projection feeds the first adjoint, the first layer the last
        const TcNet *nets[2] = {&pol, &dyn};
        for (int w = 0; w < 2; ++w) {
            const TcNet &nt = *nets[w];
            const int wl = nt.width[nt.L - 1];
            for (int i = tid; i < 16 * ns; i += TC_NT) {
                const int k = i / ns, j = i - k * ns;
                float v = 0.f;
                if (k < nt.nin && rank * ns + j < nt.width[0]) v = __ldg(nt.W_first + (long long)(rank * ns + j) * nt.nin + k);
                smem[nt.s_wfirst + i] = v;
            }
            for (int i = tid; i < TC_NOUT * ns; i += TC_NT) {
                const int o = i / ns, j = i - o * ns;
                float v = 0.f;
                if (o < nt.nout && rank * ns + j < wl) v = __ldg(nt.W_last + (long long)o * wl + rank * ns + j);
                smem[nt.s_wlast + i] = v;
            }
        }
        if (tid < TC_M)
            for (int d = 0; d < D; ++d)
                gs[p * TC_SDP + d] = (valid && prm.g_states) ? __ldg(prm.g_states + ((size_t)H * N + n) * D + d) : 0.f;
        if (prm.mm_states) tc_mm_load_table(prm, mmscr);
    }
    const long long img_floats = (long long)prm.kbmax * 1024;
    float *ximg = prm.xbuf + (size_t)tile * 2 * img_floats;
    float *opart2 = prm.opart + (size_t)tile * 2 * C * prm.nop * TC_M;   // [pass parity][rank][o][128]
    const int nop = prm.nop;
    int pass = 0;
    __syncthreads();
    TcPipe pp;
    pp.init(sched_n, sched_n * (uint32_t)H);
    tc_cluster_sync();

    float sv[HW], mk[HW];          // stored activation + mask behind the next gate (prefetched)
    if (compute) tc_gate_issue<HW>(prm, dyn, dyn.L - 1, H - 1, n, c0, sv, mk);

#pragma unroll 1
    for (int t = H - 1; t >= 0; --t) {
        int buf = 0;
        if (compute) {
            // ---- step-local adjoint factors of the tile: one coalesced block [nval][PW] ----
            const float *src = prm.pre + ((size_t)t * N + n0) * PW;
            for (int i = tid; i < nval * PW; i += TC_NT) pre[i] = __ldcg(src + i);
            // ---- moment matching adjoint: cotangent of x' = m + zhat chol(S)^T  ->  cotangent of x ----
            if (prm.mm_states) tc_mm_backward(prm, gs, mmscr, t, n0, nval);
            else CTA_SYNC();
            // ---- total dL/ds_{t+1} (carried + reward) and the dynamics density adjoint:
            //      s' = s + mu*Sy + my + z*exp(lstd) ----
            const float *q = pre + min(p, nval - 1) * PW;
            for (int d = half; d < D; d += 2) {
                const float g = valid ? gs[p * TC_SDP + d] + q[d] : 0.f;
                gsp[p * TC_SDP + d] = g;
                xo[p * TC_XO + d] = g * cst[C_SY + d];
                if (dyn.has_density) xo[p * TC_XO + D + d] = g * q[D + d];
            }
            CTA_SYNC();
        }
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const TcNet &net = which ? pol : dyn;
            const bool store = which == 1;
            const int L = net.L;
            float h[HW];
            // ---------------- adjoint of the output projection (K = nout <= 32), own columns of the last hidden ----
            if (compute) {
                const float *xr = xo + p * TC_XO;
                const float *wl = smem + net.s_wlast + half * HW;
#pragma unroll
                for (int j = 0; j < HW; ++j) h[j] = 0.f;
#pragma unroll 2
                for (int o = 0; o < net.nout; ++o) {
                    const float x = xr[o];
#pragma unroll
                    for (int j = 0; j < HW; j += 4) {
                        const float4 w = *reinterpret_cast<const float4 *>(wl + o * ns + j);
                        h[j] = fmaf(x, w.x, h[j]); h[j + 1] = fmaf(x, w.y, h[j + 1]);
                        h[j + 2] = fmaf(x, w.z, h[j + 2]); h[j + 3] = fmaf(x, w.w, h[j + 3]);
                    }
                }
            }
#pragma unroll 1
            for (int l = L - 1; l >= 0; --l) {
                if (compute) tc_gate_apply<HW>(h, sv, mk, net.keep_inv[l]);
                if (l == 0) break;
                // ---------------- hidden x hidden adjoint on the tensor cores ----------------
                if (compute) {
                    float *img = ximg + (size_t)buf * img_floats;
#pragma unroll
                    for (int j = 0; j < HW; j += 4)
                        tc_store_img(img, (c0 + j) >> 3, ((c0 + j) >> 2) & 1, p, make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]));
                }
                tc_fence_before();
                tc_cluster_sync();
                tc_fence_after();
                if (compute) {
                    if (store && valid) {        // policy: adjoint of hidden l kept for the weight gradient
                        const int npad = net.npad[l];
                        float *dl = prm.ws + net.delta_off[l] + ((size_t)t * N + n) * npad + c0;
#pragma unroll
                        for (int j = 0; j < HW; j += 4)
                            if (c0 + j < npad) *reinterpret_cast<float4 *>(dl + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
                    }
                    tc_gate_issue<HW>(prm, net, l - 1, t, n, c0, sv, mk);      // in flight while the layer runs
                }
                tc_wide_layer(prm, ring, xring, &bars, pp, sched, ximg + (size_t)buf * img_floats, net.kb[l], tmem_d, p, half);
                buf ^= 1;
                if (compute) tc_ld_acc_sum<HW>(tmem_rd, ns, min(TC_NDRV, net.kb[l]), h);
            }
            // ---------------- adjoint of the first layer: partial sums of d(input) over my columns ----------------
            if (compute) {
                const float *wf = smem + net.s_wfirst + half * HW;
                float *mine = aux + half * (nop * TC_M) + p;
#pragma unroll 1
                for (int i = 0; i < net.nin; ++i) {
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int j = 0; j < HW; j += 4) {
                        const float4 w = *reinterpret_cast<const float4 *>(wf + i * ns + j);
                        s0 = fmaf(h[j], w.x, s0); s1 = fmaf(h[j + 1], w.y, s1);
                        s0 = fmaf(h[j + 2], w.z, s0); s1 = fmaf(h[j + 3], w.w, s1);
                    }
                    mine[i * TC_M] = s0 + s1;
                }
                CTA_SYNC();
            }
            float *opart = opart2 + (size_t)pass * C * nop * TC_M;      // double-buffered by pass parity
            pass ^= 1;
            if (compute)
                for (int i = tid; i < net.nin * TC_M; i += TC_NT)
                    opart[(size_t)rank * nop * TC_M + i] = aux[i] + aux[nop * TC_M + i];
            tc_fence_before();
            tc_cluster_sync();
            tc_fence_after();
            if (compute) {
                if (store && valid) {            // policy: adjoint of hidden 0 kept for the weight gradient
                    const int npad = net.npad[0];
                    float *dl = prm.ws + net.delta_off[0] + ((size_t)t * N + n) * npad + c0;
#pragma unroll
                    for (int j = 0; j < HW; j += 4)
                        if (c0 + j < npad) *reinterpret_cast<float4 *>(dl + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
                }
                // gate behind the first adjoint of the next pass: policy at this step, dynamics at the step before
                if (which == 0) tc_gate_issue<HW>(prm, pol, pol.L - 1, t, n, c0, sv, mk);
                else if (t > 0) tc_gate_issue<HW>(prm, dyn, dyn.L - 1, t - 1, n, c0, sv, mk);
                tc_reduce_partials(opart, net.nin * TC_M, nop, nullptr, aux);
                CTA_SYNC();
                const float *q = pre + min(p, nval - 1) * PW + 2 * D;
                if (which == 0) {
                    // ---- through the input scaler d[s;a] = dx * iSx, then (action dims) the tanh squash +
                    //      policy density adjoint: a = scale*tanh(u)+bias, u = mu + z*exp(lstd) ----
                    for (int k = half; k < D; k += 2) gsp[p * TC_SDP + k] += aux[k * TC_M + p] * cst[C_ISX + k];
                    for (int u = half; u < U; u += 2) {
                        const float v = aux[(D + u) * TC_M + p] * cst[C_ISX + D + u];
                        const float ga = valid ? q[u] + v : 0.f;
                        const float du = ga * q[U + u];
                        xo[p * TC_XO + u] = du;
                        float dls = 0.f;
                        if (pol.has_density) {
                            dls = du * q[2 * U + u];
                            xo[p * TC_XO + U + u] = dls;
                        }
                        if (owner) {
                            if (prm.da_total) prm.da_total[((size_t)t * N + n) * U + u] = ga;
                            float *dd = prm.ws + pol.delta_off[pol.L] + ((size_t)t * N + n) * pol.nout;
                            dd[u] = du;
                            if (pol.has_density) dd[U + u] = dls;
                        }
                    }
                } else {
                    // ---- dL/ds_t = carried + through dynamics input + through policy input + direct cotangent ----
                    for (int d = half; d < D; d += 2) {
                        const float g0 = (valid && prm.g_states) ? __ldg(prm.g_states + ((size_t)t * N + n) * D + d) : 0.f;
                        gs[p * TC_SDP + d] = valid ? gsp[p * TC_SDP + d] + aux[d * TC_M + p] + g0 : 0.f;
                    }
                }
                CTA_SYNC();
            }
        }
    }
    if (prm.dx0 && owner && half == 0)
        for (int d = 0; d < D; ++d) prm.dx0[(size_t)n * D + d] = gs[p * TC_SDP + d];
    tc_fence_before();
    __syncthreads();
    tc_cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_d) : "memory");
}

static cudaError_t tc_launch_cfg_b(const void *fn, int C, int smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    if (C > 8) e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
}

cudaError_t launch_tc_bwd(const TcParams &prm, cudaStream_t stream) {
    const int smem_bytes = prm.smem_floats * 4;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = prm.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(prm.ntiles * prm.C);
    cfg.blockDim = dim3(TC_NTL);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
#define PMB_TC_BWD(HH)                                                                                 \
    if (prm.ns == 2 * HH) {                                                                            \
        if ((e = tc_launch_cfg_b((const void *)tc_bwd_kernel<HH>, prm.C, smem_bytes)) != cudaSuccess)  \
            return e;                                                                                  \
        return cudaLaunchKernelEx(&cfg, tc_bwd_kernel<HH>, prm);                                       \
    }
    PMB_TC_BWD(8)
    PMB_TC_BWD(16)
    PMB_TC_BWD(32)
#undef PMB_TC_BWD
    return cudaErrorInvalidValue;
}

}  // namespace pmb
