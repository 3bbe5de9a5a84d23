// Forward sweep, tensor-core cluster variant (see pmb_tc.cuh): H steps of
//   policy MLP -> Gaussian action sample -> tanh squash -> dynamics MLP -> Gaussian state sample
//   [-> moment matching of the states] -> reward
// for a tile of up to 128 particles per cluster of 16 CTAs; hidden x hidden layers on tcgen05 (3xTF32 split, fp32
// accumulation in TMEM), operands fed by TMA.  Replaces the loop body of utils.rollout (reference
// utils/rollout.py:93-163) with Policy.forward (models/core.py:221-248), DynamicsModel.forward
// (models/core.py:265-303), B/CDropout masks (models/modules.py:61,160), DiagGaussianDensity
// (models/densities.py:87-121), mm_resample_ (utils/rollout.py:20-29) and the env reward
// (envs/cartpole/env.py:41-86 et al.).  Same workspace layout as the other sweep variants, so the reverse sweeps
// and the weight-gradient kernels of any variant can follow.
#include "pmb_tc.cuh"
#include "pmb_tc_mm.cuh"
#include "pmb_host.h"

namespace pmb {

// ---------------------------------------------------------------------------------------------
// weight split: W -> per-CTA hi/lo slices in the UMMA canonical layout (once per call: the policy weights change
// every iteration).  B[n][k] (n = output column of the layer in this sweep direction, k = reduction index):
// element (n, k) of rank r = n / ns sits at  r * 2*KB*ns*8 + [lo: KB*ns*8] + (k/8)*ns*8 + ((n%ns)/8)*64 +
// ((k%8)/4)*32 + (n%8)*4 + (k%4).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_pack_kernel(const __grid_constant__ TcPackJobs jobs, float *__restrict__ wpack) {
    const TcPackJob &j = jobs.job[blockIdx.y];
    const int ns = jobs.ns, C = jobs.C;
    const int Kp = j.kb * 8;
    const long long total = (long long)C * ns * Kp;
    float *dst = wpack + j.dst_off;
    const int nN = j.transpose ? j.in : j.out;      // extent of the n axis
    const int nK = j.transpose ? j.out : j.in;      // extent of the k axis
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // consecutive threads walk k fastest for the forward orientation (coalesced reads of W rows)
        const int k = (int)(i % Kp);
        const int n = (int)(i / Kp);
        float v = 0.f;
        if (n < nN && k < nK) v = j.transpose ? __ldg(j.W + (long long)k * j.in + n) : __ldg(j.W + (long long)n * j.in + k);
        const float h = tc_tf32_hi(v);
        const int r = n / ns, nl = n - r * ns;
        const long long o = (long long)r * 2 * j.kb * ns * 8 + (long long)(k >> 3) * ns * 8 + (nl >> 3) * 64 + ((k & 7) >> 2) * 32 +
                            (nl & 7) * 4 + (k & 3);
        dst[o] = h;
        dst[o + (long long)j.kb * ns * 8] = v - h;
    }
}

cudaError_t launch_tc_pack(const TcPackJobs &jobs, float *wpack, cudaStream_t stream) {
    if (jobs.n == 0) return cudaSuccess;
    dim3 grid(96, jobs.n);
    tc_pack_kernel<<<grid, 256, 0, stream>>>(jobs, wpack);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// resident fp32 operands of this CTA's columns (first layer, output projection, biases)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_load_resident(const TcParams &prm, const TcNet &n, float *smem, int rank, bool fwd) {
    const int tid = threadIdx.x, ns = prm.ns;
    const int c0 = rank * ns;
    const int wl = n.width[n.L - 1];
    for (int i = tid; i < 16 * ns; i += TC_NT) {
        const int k = i / ns, j = i - k * ns;
        float v = 0.f;
        if (k < n.nin && c0 + j < n.width[0]) v = __ldg(n.W_first + (long long)(c0 + j) * n.nin + k);
        smem[n.s_wfirst + i] = v;
    }
    for (int i = tid; i < TC_NOUT * ns; i += TC_NT) {
        const int o = i / ns, j = i - o * ns;
        float v = 0.f;
        if (o < n.nout && c0 + j < wl) v = __ldg(n.W_last + (long long)o * wl + c0 + j);
        smem[n.s_wlast + i] = v;
    }
    if (fwd) {
        for (int i = tid; i < (n.L + 1) * TC_MAXNS; i += TC_NT) {
            const int l = i / TC_MAXNS, j = i - l * TC_MAXNS;
            float v = 0.f;
            if (n.bias_off[l] >= 0) {
                if (l < n.L) {
                    if (j < ns && c0 + j < n.npad[l]) v = __ldg(prm.ws + n.bias_off[l] + c0 + j);
                } else if (j < n.nout) {
                    v = __ldg(prm.ws + n.bias_off[l] + j);
                }
            }
            smem[n.s_bias + i] = v;
        }
    }
}

// dropout mask values of (particle n, columns c0 .. c0+HW) of hidden layer l, times 1/keep; 1/keep where the layer
// has no mask.  Columns past the padded width read as 0 (their activations are 0 anyway: zero weights and biases).
template <int HW>
__device__ __forceinline__ void tc_load_mask(const TcParams &prm, const TcNet &n, int l, int nld, int c0, float (&mk)[HW]) {
    const int npad = n.npad[l];
    const float ki = n.keep_inv[l];
    if (n.mask_off[l] >= 0) {
        const float *src = prm.ws + n.mask_off[l] + (long long)nld * npad + c0;
#pragma unroll
        for (int j = 0; j < HW; j += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + j < npad) v = __ldg(reinterpret_cast<const float4 *>(src + j));
            mk[j] = v.x * ki; mk[j + 1] = v.y * ki; mk[j + 2] = v.z * ki; mk[j + 3] = v.w * ki;
        }
    } else {
#pragma unroll
        for (int j = 0; j < HW; ++j) mk[j] = ki;
    }
}

// raw dropout-mask values of (particle n, columns c0 .. c0+HW) of hidden layer l: issued early, consumed by the
// layer's epilogue (tc_mask_apply) so that the L2 latency hides behind the layer
template <int HW>
__device__ __forceinline__ void tc_mask_issue(const TcParams &prm, const TcNet &n, int l, int nld, int c0, float (&mk)[HW]) {
    const int npad = n.npad[l];
    if (n.mask_off[l] >= 0) {
        const float *src = prm.ws + n.mask_off[l] + (long long)nld * npad + c0;
#pragma unroll
        for (int j = 0; j < HW; j += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + j < npad) v = __ldg(reinterpret_cast<const float4 *>(src + j));
            mk[j] = v.x; mk[j + 1] = v.y; mk[j + 2] = v.z; mk[j + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < HW; ++j) mk[j] = 1.f;
    }
}

template <int HW>
__global__ void __launch_bounds__(TC_NTL, 1) tc_fwd_kernel(const __grid_constant__ TcParams prm) {
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
    const int p = 32 * (warp & 3) + lane, half = (warp >> 2) & 1;   // particle row of the tile, column half of the slice
    const bool valid = compute && p < nval;
    const int n = n0 + min(p, nval - 1);                          // clamped: safe address for loads
    const int c0 = rank * ns + half * HW;                         // first column this thread owns
    const bool owner = valid && (p % C) == rank;                  // this CTA writes particle p's trajectory

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
        const TcNet *const order[2] = {&prm.pol, &prm.dyn};
        sched_n = tc_build_schedule(prm, order, false, rank, sched);
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
    float *xin = smem + prm.off_xin;          // [128][TC_SDP] input rows of the net being evaluated
    float *st = smem + prm.off_st;            // [128][TC_SDP] current state
    float *aux = smem + prm.off_aux;          // [2][nop][128] partial sums of the column halves / reduced outputs
    float *zt = smem + prm.off_z;             // [2][128][TC_SDP] density noise of the tile: policy, dynamics
    float *mmscr = smem + prm.off_mm;
    float *ring = smem + prm.off_ring;        // weight stages
    float *xring = smem + prm.off_xring;      // image stages
    if (compute) {
        load_constants(prm, cst);
        tc_load_resident(prm, prm.pol, smem, rank, true);
        tc_load_resident(prm, prm.dyn, smem, rank, true);
        if (tid < TC_M) {
            for (int d = 0; d < D; ++d) {
                const float v = valid ? prm.x0[(size_t)n * D + d] : 0.f;
                st[p * TC_SDP + d] = v;
                xin[p * TC_SDP + d] = v;
                if (owner) prm.states[(size_t)n * D + d] = v;
            }
            // PEGASUS: the density noise is constant over the horizon (per-step tables are re-read every step)
            if (prm.pol.has_density) for (int u = 0; u < U; ++u) zt[p * TC_SDP + u] = __ldg(prm.pol.z + (size_t)n * U + u);
            if (prm.dyn.has_density) for (int d = 0; d < D; ++d) zt[(TC_M + p) * TC_SDP + d] = __ldg(prm.dyn.z + (size_t)n * D + d);
        }
        if (prm.mm_states) tc_mm_load_table(prm, mmscr);
    }
    const float elmax_pol = expf(prm.pol.lmax), elmax_dyn = expf(prm.dyn.lmax);
    const long long img_floats = (long long)prm.kbmax * 1024;
    float *ximg = prm.xbuf + (size_t)tile * 2 * img_floats;        // two fp32 images
    float *opart2 = prm.opart + (size_t)tile * 2 * C * prm.nop * TC_M;   // [pass parity][rank][o][128]
    const int nop = prm.nop;
    int pass = 0;
    __syncthreads();
    TcPipe pp;
    pp.init(sched_n, sched_n * (uint32_t)H);
    tc_cluster_sync();

    float mk[HW];                  // raw dropout mask of the next layer to finish (prefetched)
    if (compute) tc_mask_issue<HW>(prm, prm.pol, 0, n, c0, mk);

#pragma unroll 1
    for (int t = 0; t < H; ++t) {
        const bool dbg_step = prm.dbg != nullptr && blockIdx.x == 0 && t == H / 2;
        int buf = 0;
        if (compute && tid < TC_M) {     // per-step noise tables (only when the caller pre-drew [H, N, .] noise)
            if (prm.pol.has_density && prm.pol.zstride != 0)
                for (int u = 0; u < U; ++u) zt[p * TC_SDP + u] = __ldg(prm.pol.z + (size_t)t * prm.pol.zstride + (size_t)n * U + u);
            if (prm.dyn.has_density && prm.dyn.zstride != 0)
                for (int d = 0; d < D; ++d)
                    zt[(TC_M + p) * TC_SDP + d] = __ldg(prm.dyn.z + (size_t)t * prm.dyn.zstride + (size_t)n * D + d);
        }
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const TcNet &net = which ? prm.dyn : prm.pol;
            const int L = net.L;
            float h[HW];
            TC_MARK(0 + 16 * which);
            // ---------------- first layer (K = nin <= 16) on the FP32 pipe, own columns ----------------
            if (compute) {
                const float *xr = xin + p * TC_SDP;
                const float *wf = smem + net.s_wfirst + half * HW;
#pragma unroll
                for (int j = 0; j < HW; ++j) h[j] = 0.f;
#pragma unroll 2
                for (int k = 0; k < net.nin; ++k) {
                    const float x = xr[k];
#pragma unroll
                    for (int j = 0; j < HW; j += 4) {
                        const float4 w = *reinterpret_cast<const float4 *>(wf + k * ns + j);
                        h[j] = fmaf(x, w.x, h[j]); h[j + 1] = fmaf(x, w.y, h[j + 1]);
                        h[j + 2] = fmaf(x, w.z, h[j + 2]); h[j + 3] = fmaf(x, w.w, h[j + 3]);
                    }
                }
            }
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                // ---------------- epilogue: bias, ReLU, dropout mask / keep (modules.py:61,160) ----------------
                if (compute) {
                    const float *bs = smem + net.s_bias + l * TC_MAXNS + half * HW;
                    const float ki = net.keep_inv[l];
#pragma unroll
                    for (int j = 0; j < HW; j += 4) {
                        const float4 b = *reinterpret_cast<const float4 *>(bs + j);
                        h[j] = fmaxf(h[j] + b.x, 0.f) * mk[j] * ki;
                        h[j + 1] = fmaxf(h[j + 1] + b.y, 0.f) * mk[j + 1] * ki;
                        h[j + 2] = fmaxf(h[j + 2] + b.z, 0.f) * mk[j + 2] * ki;
                        h[j + 3] = fmaxf(h[j + 3] + b.w, 0.f) * mk[j + 3] * ki;
                    }
                }
                if (l + 1 == L) break;
                // next layer's A operand: my columns, plain fp32, into the exchange image
                if (compute) {
                    float *img = ximg + (size_t)buf * img_floats;
#pragma unroll
                    for (int j = 0; j < HW; j += 4)
                        tc_store_img(img, (c0 + j) >> 3, ((c0 + j) >> 2) & 1, p, make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]));
                }
                TC_MARK(1 + 16 * which);
                tc_fence_before();
                tc_cluster_sync();
                tc_fence_after();
                TC_MARK(3 + 16 * which);
                if (compute) {
                    // kept for the reverse sweep / weight gradient (after the barrier: off the exchange's critical path)
                    const int npad = net.npad[l];
                    float *sv = prm.ws + net.saved_off[l] + ((size_t)t * N + n) * npad + c0;
#pragma unroll
                    for (int j = 0; j < HW; j += 4)
                        if (valid && c0 + j < npad) *reinterpret_cast<float4 *>(sv + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
                    tc_mask_issue<HW>(prm, net, l + 1, n, c0, mk);      // in flight while the layer runs
                }
                // ---------------- hidden x hidden layer on the tensor cores ----------------
                tc_wide_layer(prm, ring, xring, &bars, pp, sched, ximg + (size_t)buf * img_floats, net.kb[l], tmem_d, p, half,
                              (dbg_step && which == 1) ? prm.dbg + 512 : nullptr);
                buf ^= 1;
                TC_MARK(4 + 16 * which);
                if (compute) tc_ld_acc_sum<HW>(tmem_rd, ns, min(TC_NDRV, net.kb[l]), h);
            }
            TC_MARK(5 + 16 * which);
            // ---------------- output projection: partial sums over my columns ----------------
            if (compute) {
                const float *wl = smem + net.s_wlast + half * HW;
                float *mine = aux + half * (nop * TC_M) + p;
#pragma unroll 1
                for (int o = 0; o < net.nout; ++o) {
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int j = 0; j < HW; j += 4) {
                        const float4 w = *reinterpret_cast<const float4 *>(wl + o * ns + j);
                        s0 = fmaf(h[j], w.x, s0); s1 = fmaf(h[j + 1], w.y, s1);
                        s0 = fmaf(h[j + 2], w.z, s0); s1 = fmaf(h[j + 3], w.w, s1);
                    }
                    mine[o * TC_M] = s0 + s1;
                }
                CTA_SYNC();
            }
            float *opart = opart2 + (size_t)pass * C * nop * TC_M;      // double-buffered by pass parity
            pass ^= 1;
            if (compute) {
                // the two column halves meet; [rank][o][particle] rows of 128 floats, coalesced
                for (int i = tid; i < net.nout * TC_M; i += TC_NT)
                    opart[(size_t)rank * nop * TC_M + i] = aux[i] + aux[nop * TC_M + i];
            }
            TC_MARK(6 + 16 * which);
            tc_fence_before();
            tc_cluster_sync();
            tc_fence_after();
            TC_MARK(8 + 16 * which);
            if (compute) {
                {   // last hidden layer, kept for the reverse sweep / weight gradient
                    const int npad = net.npad[L - 1];
                    float *sv = prm.ws + net.saved_off[L - 1] + ((size_t)t * N + n) * npad + c0;
#pragma unroll
                    for (int j = 0; j < HW; j += 4)
                        if (valid && c0 + j < npad) *reinterpret_cast<float4 *>(sv + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
                }
                tc_mask_issue<HW>(prm, which ? prm.pol : prm.dyn, 0, n, c0, mk);     // first layer of the next pass
                // every CTA adds the C partials in rank order: bit-identical outputs everywhere
                tc_reduce_partials(opart, net.nout * TC_M, nop, smem + net.s_bias + L * TC_MAXNS, aux);
                CTA_SYNC();
            }
            TC_MARK(9 + 16 * which);
            if (compute) {
                if (which == 0) {
                    // ---- Gaussian action sample + tanh squash (densities.py:95-119, core.py:243);
                    //      dynamics input (core.py:269,177): thread (particle, half) takes every second dim ----
                    for (int u = half; u < U; u += 2) {
                        const float mu = aux[u * TC_M + p];
                        float uu = mu, ls = 0.f;
                        if (net.has_density) {
                            ls = aux[(U + u) * TC_M + p];
                            uu += zt[p * TC_SDP + u] * exp_clamped_logstd(ls, net.lmax, elmax_pol);
                        }
                        const float a = cst[C_SCALE + u] * tanhf(uu) + cst[C_BIAS + u];
                        xin[p * TC_SDP + D + u] = (a - cst[C_MX + D + u]) * cst[C_ISX + D + u];
                        if (owner) {
                            prm.actions[((size_t)t * N + n) * U + u] = a;
                            float *raw = prm.ws + net.raw_off + ((size_t)t * N + n) * net.nout;
                            raw[u] = mu;
                            if (net.has_density) raw[U + u] = ls;
                        }
                    }
                    for (int d = half; d < D; d += 2) xin[p * TC_SDP + d] = (st[p * TC_SDP + d] - cst[C_MX + d]) * cst[C_ISX + d];
                } else {
                    // ---- Gaussian state sample, s' = s + delta (densities.py:100-119, core.py:293,298) ----
                    for (int d = half; d < D; d += 2) {
                        const float sy = cst[C_SY + d], my = cst[C_MY + d];
                        const float mu = aux[d * TC_M + p];
                        float delta, ls = 0.f;
                        if (net.has_density) {
                            ls = aux[(D + d) * TC_M + p];
                            // exp(clamped log-std + log Sy) = Sy * exp(clamped log-std)   (densities.py:105)
                            delta = (mu * sy + my) + zt[(TC_M + p) * TC_SDP + d] * (sy * exp_clamped_logstd(ls, net.lmax, elmax_dyn));
                        } else {
                            delta = mu * sy + my;
                        }
                        const float s1 = st[p * TC_SDP + d] + delta;
                        st[p * TC_SDP + d] = valid ? s1 : 0.f;
                        if (owner) {
                            float *raw = prm.ws + net.raw_off + ((size_t)t * N + n) * net.nout;
                            raw[d] = mu;
                            if (net.has_density) raw[D + d] = ls;
                            if (prm.mm_states) prm.s1pre[((size_t)t * N + n) * D + d] = s1;
                        }
                    }
                    if (prm.mm_states) {
                        CTA_SYNC();
                        TC_MARK(11 + 16 * which);
                        tc_mm_forward(prm, st, mmscr, t, n0, nval, rank);      // rollout.py:121-132
                    }
                    for (int d = half; d < D; d += 2) {
                        const float s1 = st[p * TC_SDP + d];
                        xin[p * TC_SDP + d] = s1;
                        if (owner) prm.states[((size_t)(t + 1) * N + n) * D + d] = s1;
                    }
                }
                CTA_SYNC();
            }
            TC_MARK(10 + 16 * which);
        }
    }
    // ---- rewards r_t = scale*exp(-0.5*(d^T Q d + a^T R a)) + offset on (s_{t+1}, a_t) for every step
    //      (envs/cartpole/env.py:62-86).  Nothing in the recurrence consumes them: evaluated here, off the serial
    //      chain, for the particles whose trajectory THIS CTA wrote. ----
    if (compute) {
        for (int i = tid; i < H * TC_M; i += TC_NT) {
            const int tt = i / TC_M, q = i - tt * TC_M;
            if (q >= nval || (q % C) != rank) continue;
            // the reward sees the next state BEFORE moment matching (models/core.py:293 runs inside dynamics())
            const float *s1 = prm.mm_states ? prm.s1pre + ((size_t)tt * N + n0 + q) * D
                                            : prm.states + ((size_t)(tt + 1) * N + n0 + q) * D;
            const float *a = prm.actions + ((size_t)tt * N + n0 + q) * U;
            float dl[PMB_MAX_REWARD_ROWS];
            for (int r = 0; r < prm.KR; ++r) {
                float acc = cst[C_C0 + r];
                for (int d = 0; d < D; ++d) acc = fmaf(cst[C_C + r * SD + d], s1[d], acc);
                dl[r] = acc;
            }
            float cost = 0.f;
            for (int r = 0; r < prm.KR; ++r) {
                float qq = 0.f;
                for (int j = 0; j < prm.KR; ++j) qq = fmaf(dl[j], cst[C_Q + j * SD + r], qq);
                cost = fmaf(qq, dl[r], cost);
            }
            for (int u = 0; u < U; ++u) {
                float qq = 0.f;
                for (int v = 0; v < U; ++v) qq = fmaf(a[v], cst[C_R + v * SD + u], qq);
                cost = fmaf(qq, a[u], cost);
            }
            prm.rewards[(size_t)tt * N + n0 + q] = prm.rew_scale * expf(-0.5f * cost) + prm.rew_offset;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_cluster_sync();          // peers may still read my partial sums / image until here
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_d) : "memory");
}

static cudaError_t tc_launch_cfg(const void *fn, int C, int smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    if (C > 8) e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
}

cudaError_t launch_tc_fwd(const TcParams &prm, cudaStream_t stream) {
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
#define PMB_TC_FWD(HH)                                                                                \
    if (prm.ns == 2 * HH) {                                                                           \
        if ((e = tc_launch_cfg((const void *)tc_fwd_kernel<HH>, prm.C, smem_bytes)) != cudaSuccess)   \
            return e;                                                                                 \
        return cudaLaunchKernelEx(&cfg, tc_fwd_kernel<HH>, prm);                                      \
    }
    PMB_TC_FWD(8)
    PMB_TC_FWD(16)
    PMB_TC_FWD(32)
#undef PMB_TC_FWD
    return cudaErrorInvalidValue;
}

// co-resident clusters of the forward kernel (0 when the query is unavailable, e.g. no device)
int tc_max_active_clusters(int C, int smem_bytes) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(C * 16);
    cfg.blockDim = dim3(TC_NTL);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nmax = 0;
    const void *fn = (const void *)tc_fwd_kernel<16>;
    if (tc_launch_cfg(fn, C, smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (cudaOccupancyMaxActiveClusters(&nmax, fn, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nmax;
}

}  // namespace pmb
