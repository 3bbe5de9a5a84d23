// Reverse sweep of the imagined rollout (back-propagation through time), no recompute:
// consumes the activations the forward sweep stored, walks the steps t = H-1 .. 0 and produces
//   * dL/dx0,
//   * the per-layer output adjoints ("deltas") of the POLICY net for every (t, particle), which the
//     batched weight-gradient GEMMs (pmb_wgrad.cu) contract afterwards over the (H*N) axis.
// Dynamics weight gradients are never formed (the reference computes and discards them,
// SURVEY.md App. D-4).  Replaces loss.backward() through utils.rollout (reference
// algorithms/mc_pilco.py:197); adjoint formulas are those of oracle/rollout_oracle.py::manual_backward,
// which tests/test_oracle_backward.py checks against autograd.
//
// Everything that depends only on forward values (reward adjoint, density / tanh derivative factors,
// direct cotangents) is computed ONE STEP AHEAD into a double-buffered "pre" block in shared memory,
// from global loads issued at the top of the previous step, and the stored hidden activations of the
// step arrive by TMA bulk copies issued one step ahead: the serial chain only touches shared memory.
#include "pmb_internal.cuh"
#include "pmb_mm.cuh"

namespace pmb {

// float offsets inside the per-particle scratch block (`misc`), P-scaled
constexpr int M_GS = 0, M_GSP = 1, M_OBUF = 2, M_STG_S1 = 3, M_STG_A = 4, M_PRE = 5;   // x P*SD
constexpr int PRE_RS = 0, PRE_RA = 1, PRE_FD = 2, PRE_TP = 3, PRE_FP = 4, PRE_GS0 = 5, PRE_N = 6;

// Backward through one net.  `in` holds the adjoint of the net's raw outputs as a [nout][P] tile.
// Wide layers l = nlin-1 .. 1 (weights W_l as stored, [out][in]) produce the adjoint of hidden l-1,
// gated by the stored activation; the final narrow layer (W_0^T) leaves d(input)[p][nin] in obuf.
// With kStoreDelta every hidden adjoint is also written to global for the weight gradient.
template <int P>
__device__ __forceinline__ void net_backward(const SweepParams &prm, const NetSweep &net, bool kStoreDelta,
                                             int &sched_i, float *&in, float *&out, float *obuf, float *smem,
                                             float *red, const float *sav, Stream &S, const NarrowMap &nm, int t,
                                             int n0, float *part, int &wpg) {
    const int N = prm.N;
    const Lin &L0 = net.lin[0];
#pragma unroll 1
    for (int l = net.nlin - 1; l >= 1; --l) {
        const Lin &L = net.lin[l];      // wide: K = outputs of linear l (padded), Npad = width of hidden l-1
        const int h = l - 1;            // hidden layer whose adjoint we produce
        const int npad = L.Npad;
        const float *mask_s = net.mask_soff[h] >= 0 ? smem + net.mask_soff[h] : nullptr;
        const float *mask_g = (!mask_s && net.mask_off[h] >= 0) ? prm.ws + net.mask_off[h] : nullptr;
        const float *sv = sav + net.sav_soff[h];
        // y = relu(pre) * mask / keep  =>  dpre = (dy / keep) * mask * [pre > 0];  y != 0 <=> pre > 0, mask != 0
        const float inv_keep = 1.f / net.keep[h];
        float *dl = prm.ws + net.delta_off[h] + ((size_t)t * N + n0) * npad;
        if (!L.streamed) {
            // adjoint of the output projection: K = 2D / 2U rows
            thin_layer<P>(L, smem, in, [&](int j, float (&acc)[P]) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float mk = 1.f;
                    if (mask_s) mk = mask_s[p * npad + j];
                    else if (mask_g) mk = __ldg(mask_g + (size_t)min(n0 + p, N - 1) * npad + j);
                    const float hh = sv[p * npad + j];
                    const float x = hh != 0.f ? acc[p] * inv_keep * mk : 0.f;
                    out[j * P + p] = x;
                    if (kStoreDelta && n0 + p < N) dl[(size_t)p * npad + j] = x;
                }
            });
            for (int j = L.Nout + threadIdx.x; j < npad; j += NT) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    out[j * P + p] = 0.f;
                    if (kStoreDelta && n0 + p < N) dl[(size_t)p * npad + j] = 0.f;
                }
            }
        } else {
            WideMap m;
            m.set(npad);
            const int col = 4 * m.cq;
            float *dst = out + col * P;
            // the adjoint of hidden layer 0 also forms its share of d(input) = delta_0 W_0 (narrow_fused)
            const bool fuse = (l == 1);
            const int cqn = npad >> 2;
            wpg = fuse ? ((cqn <= 32 ? 32 : cqn <= 64 ? 64 : cqn <= 128 ? 128 : 256) >> 5) : 1;
            const int wig = (threadIdx.x >> 5) & (wpg - 1);
            wide_layer<P>(L, &prm.sched[sched_i], smem, in, red, S, m, [&](int p, float4 v, bool act) {
              if (act) {
                float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
                if (mask_s) mk = *reinterpret_cast<const float4 *>(mask_s + p * npad + col);
                else if (mask_g) mk = __ldg(reinterpret_cast<const float4 *>(mask_g + (size_t)min(n0 + p, N - 1) * npad + col));
                const float4 hh = *reinterpret_cast<const float4 *>(sv + p * npad + col);
                v.x = hh.x != 0.f ? v.x * inv_keep * mk.x : 0.f;
                v.y = hh.y != 0.f ? v.y * inv_keep * mk.y : 0.f;
                v.z = hh.z != 0.f ? v.z * inv_keep * mk.z : 0.f;
                v.w = hh.w != 0.f ? v.w * inv_keep * mk.w : 0.f;
                dst[p] = v.x;
                dst[P + p] = v.y;
                dst[2 * P + p] = v.z;
                dst[3 * P + p] = v.w;
                if (kStoreDelta && n0 + p < N) *reinterpret_cast<float4 *>(dl + (size_t)p * npad + col) = v;
              }
              if (fuse) narrow_fused(v, act, p, col, smem + L0.soff, L0.K, L0.Nout, part, wpg, wig);
            });
            ++sched_i;
            if (fuse) return;               // d(input) sits in `part` as wpg partials per value
        }
        float *tmp = in; in = out; out = tmp;
    }
    wpg = 0;
    narrow_layer<P>(L0, nm, smem, in, obuf, nullptr, red);
}

template <int P>
__global__ void __launch_bounds__(NT_LAUNCH, 1) rollout_bwd_kernel(const __grid_constant__ SweepParams prm) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) RingBars ring;
    __shared__ __align__(8) uint64_t sav_full[2], sav_empty[2];
    __shared__ __align__(16) ChunkDesc chunk_tab[MAXCHUNKS];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * P;
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H, KR = prm.KR;
    const NetSweep &pol = prm.pol;
    const NetSweep &dyn = prm.dyn;
    const bool has_sav = (pol.nlin > 1) || (dyn.nlin > 1);
    // tiles and scratch start finite; done by all threads BEFORE the producer may issue any TMA write
    for (int i = tid; i < prm.off_stage; i += NT_LAUNCH) smem[i] = 0.f;
    // ---- barriers, then the producer warp peels off: it feeds the weight ring and, one step ahead, the
    //      stored hidden activations of each step (TMA bulk copies, one per hidden layer) ----
    if (tid == 0) {
        for (int s = 0; s < prm.nstages; ++s) {
            mbar_init(&ring.full[s], 1);
            mbar_init(&ring.empty[s], NWARP);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sav_full[b], 1);
            mbar_init(&sav_empty[b], NWARP);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();                       // the only barrier all 288 threads take
    if (tid >= NT) {
        if (tid == NT) {
            fence_proxy_async();
            float *savb_p = smem + prm.off_sav;
            ring_fill_table(prm, chunk_tab);
            RingProducer rp;
            rp.init();
            uint32_t sav_bytes = 0;
            for (int n = 0; n < 2; ++n) {
                const NetSweep &net = n ? pol : dyn;
                for (int h = 0; h + 1 < net.nlin; ++h) sav_bytes += (uint32_t)(P * net.lin[h + 1].Npad) * 4u;
            }
            for (int s = 0; s < H; ++s) {
                const int tt = H - 1 - s, b = s & 1;
                if (has_sav) {
                    if (s >= 2) {
                        mbar_wait(&sav_empty[b], ((s >> 1) - 1) & 1);
                        fence_proxy_async();
                    }
                    mbar_expect_tx(&sav_full[b], sav_bytes);
                    for (int n = 0; n < 2; ++n) {
                        const NetSweep &net = n ? pol : dyn;
                        for (int h = 0; h + 1 < net.nlin; ++h) {
                            const int npad = net.lin[h + 1].Npad;
                            tma_bulk_g2s(savb_p + (size_t)b * prm.sav_floats + net.sav_soff[h],
                                         prm.ws + net.saved_off[h] + ((size_t)tt * N + n0) * npad,
                                         (uint32_t)(P * npad) * 4u, &sav_full[b]);
                        }
                    }
                }
                if (prm.chunks_per_step > 0) rp.issue(prm, smem + prm.off_stage, &ring, chunk_tab, prm.chunks_per_step);
            }
        }
        return;
    }
    float *cst = smem + prm.off_cst;
    float *act0 = smem + prm.off_act0;
    float *act1 = smem + prm.off_act1;
    float *red = smem + prm.off_red;
    float *misc = smem + prm.off_misc;
    float *savb = smem + prm.off_sav;
    constexpr int BL = P * SD;   // one [P][SD] block
    float *gs = misc + M_GS * BL, *gsp = misc + M_GSP * BL, *obuf = misc + M_OBUF * BL;
    float *stg_s1 = misc + M_STG_S1 * BL, *stg_a = misc + M_STG_A * BL;
    float *pre0 = misc + M_PRE * BL;                    // two buffers of PRE_N blocks
    float *stg_w = pre0 + 2 * PRE_N * BL;               // [P]
    float *part = misc + 320 * P;                       // partials of the fused input projection

    // ---- thread roles (fixed for the whole horizon) ----
    const bool roleA = tid < P * U;                       // (particle, action dim)
    const int a_p = roleA ? tid / U : 0, a_u = roleA ? tid - a_p * U : 0;
    const int a_n = min(n0 + a_p, N - 1);
    const bool roleB = tid >= 128 && tid - 128 < P * D;   // (particle, state dim)
    const int b_p = roleB ? (tid - 128) / D : 0, b_d = roleB ? (tid - 128) - b_p * D : 0;
    const int b_n = min(n0 + b_p, N - 1);
    const bool roleX = tid < P * (D + U);                 // (particle, dynamics-input dim)
    const int x_p = roleX ? tid / (D + U) : 0, x_k = roleX ? tid - x_p * (D + U) : 0;
    const bool roleR = tid >= 224 && tid - 224 < P;       // particle (reward weight)
    const int r_p = roleR ? tid - 224 : 0;
    const int r_n = min(n0 + r_p, N - 1);

    load_constants(prm, cst);
    load_resident(prm, smem, n0);
    if (roleB) gs[b_p * SD + b_d] = prm.g_states ? __ldg(prm.g_states + ((size_t)H * N + b_n) * D + b_d) : 0.f;
    NarrowMap nm_pol, nm_dyn;
    nm_pol.set<P>(pol.lin[0]);
    nm_dyn.set<P>(dyn.lin[0]);
    Stream S;
    S.init(&prm, smem, &ring);
    CTA_SYNC();

    // prefetch registers of the one-step-ahead precompute
    float pf_s1 = 0.f, pf_ls = 0.f, pf_zd = 0.f, pf_gs = 0.f;                 // role B
    float pf_a = 0.f, pf_mu = 0.f, pf_lsp = 0.f, pf_zp = 0.f, pf_ga = 0.f;    // role A
    float pf_r = 0.f, pf_gr = 0.f;                                            // role R
    auto prefetch = [&](int tt) {
        if (roleB) {
            // the reward (and the mm adjoint) act on the next state BEFORE moment matching
            pf_s1 = prm.mm_states ? __ldg(prm.s1pre + ((size_t)tt * N + b_n) * D + b_d)
                                  : __ldg(prm.states + ((size_t)(tt + 1) * N + b_n) * D + b_d);
            if (dyn.has_density) {
                pf_ls = __ldg(prm.ws + dyn.outsaved_off + ((size_t)tt * N + b_n) * dyn.nout + D + b_d);
                pf_zd = __ldg(dyn.z + (size_t)tt * dyn.zstride + (size_t)b_n * D + b_d);
            }
            pf_gs = prm.g_states ? __ldg(prm.g_states + ((size_t)tt * N + b_n) * D + b_d) : 0.f;
        }
        if (roleA) {
            const float *op = prm.ws + pol.outsaved_off + ((size_t)tt * N + a_n) * pol.nout;
            pf_a = __ldg(prm.actions + ((size_t)tt * N + a_n) * U + a_u);
            pf_mu = __ldg(op + a_u);
            if (pol.has_density) {
                pf_lsp = __ldg(op + U + a_u);
                pf_zp = __ldg(pol.z + (size_t)tt * pol.zstride + (size_t)a_n * U + a_u);
            }
            pf_ga = prm.g_actions ? __ldg(prm.g_actions + ((size_t)tt * N + a_n) * U + a_u) : 0.f;
        }
        if (roleR) {
            pf_r = __ldg(prm.rewards + (size_t)tt * N + r_n);
            pf_gr = prm.g_rewards ? __ldg(prm.g_rewards + (size_t)tt * N + r_n) : 0.f;
        }
    };
    // first half of the precompute: factors that need no cross-thread data (+ staging of s', a, w)
    auto precompute_a = [&](float *pre) {
        if (roleB) {
            stg_s1[b_p * SD + b_d] = pf_s1;
            float fd = 0.f;
            if (dyn.has_density) {
                const float lst = clamp_logstd(pf_ls, dyn.lmax) + cst[C_LSY + b_d];
                fd = pf_zd * expf(lst) * sigmoid_f(dyn.lmax - pf_ls);     // d s' / d log_std (raw)
            }
            pre[PRE_FD * BL + b_p * SD + b_d] = fd;
            pre[PRE_GS0 * BL + b_p * SD + b_d] = pf_gs;
        }
        if (roleA) {
            stg_a[a_p * SD + a_u] = pf_a;
            const float sc = cst[C_SCALE + a_u];
            float tp, fp = 0.f;
            if (pol.has_density) {
                const float el = expf(clamp_logstd(pf_lsp, pol.lmax));
                const float th = tanhf(pf_mu + pf_zp * el);
                tp = sc * (1.f - th * th);                                  // d a / d u
                fp = pf_zp * el * sigmoid_f(pol.lmax - pf_lsp);             // d u / d log_std (raw)
            } else {
                const float th = tanhf(pf_mu);
                tp = sc * (1.f - th * th);
            }
            pre[PRE_TP * BL + a_p * SD + a_u] = tp;
            pre[PRE_FP * BL + a_p * SD + a_u] = fp;
        }
        if (roleR) stg_w[r_p] = -0.5f * pf_gr * (pf_r - prm.rew_offset);   // g_r * d r / d cost, r - off = scale*exp(-cost)
    };
    // second half: reward adjoint on (s', a):  w * C^T (Q+Q^T) delta  and  g_a + w * (R+R^T) a
    auto precompute_b = [&](float *pre) {
        if (roleB) {
            float dl[PMB_MAX_REWARD_ROWS];
            for (int i = 0; i < KR; ++i) {
                float s = cst[C_C0 + i];
                for (int d = 0; d < D; ++d) s = fmaf(cst[C_C + i * SD + d], stg_s1[b_p * SD + d], s);
                dl[i] = s;
            }
            float acc = 0.f;
            for (int i = 0; i < KR; ++i) {
                float qd = 0.f;
                for (int j = 0; j < KR; ++j) qd = fmaf(cst[C_QS + i * SD + j], dl[j], qd);
                acc = fmaf(qd, cst[C_C + i * SD + b_d], acc);
            }
            pre[PRE_RS * BL + b_p * SD + b_d] = stg_w[b_p] * acc;
        }
        if (roleA) {
            float s = 0.f;
            for (int v = 0; v < U; ++v) s = fmaf(cst[C_RS + a_u * SD + v], stg_a[a_p * SD + v], s);
            pre[PRE_RA * BL + a_p * SD + a_u] = pf_ga + stg_w[a_p] * s;
        }
    };
    MMSmem mmS;
    MMGroup grp;
    unsigned epoch = 0;
    if (prm.mm_states) {
        mmS.carve<P>(smem + prm.off_mm);
        mmS.xs = stg_s1;            // the staged pre-mm particles of the current step
        grp.set<P>(prm, n0);
    }

    // ---- prologue: everything step H-1 needs ----
    int cur = 0;
    uint32_t spar[2] = {0u, 0u};
    prefetch(H - 1);
    precompute_a(pre0 + cur * PRE_N * BL);
    CTA_SYNC();
    precompute_b(pre0 + cur * PRE_N * BL);
    CTA_SYNC();

#pragma unroll 1
    for (int t = H - 1; t >= 0; --t) {
        const bool dbg_on = prm.dbg != nullptr && blockIdx.x == 0 && tid == 0 && t == H / 2;
        PMB_MARK(32);
        const int nxt = cur ^ 1;
        float *pre = pre0 + cur * PRE_N * BL;
        float *pren = pre0 + nxt * PRE_N * BL;
        int sched_i = 0;
        float *in = act0, *out = act1;
        // ---- one step ahead: stored activations + scalars of step t-1 ----
        if (t > 0) prefetch(t - 1);
        // ---- moment matching adjoint: cotangent of x' = m + zhat chol(S)^T  ->  cotangent of x ----
        if (prm.mm_states) {
            if (roleB) mmS.zs[b_p * SD + b_d] = __ldg(prm.z_mm + (size_t)((t + n0 + b_p) % N) * D + b_d);
            mm_states_backward<P>(prm, mmS, grp, t, gs, roleB, b_p, b_d, epoch);
        }
        // ---- total dL/ds_{t+1} (carried + reward) and the dynamics density adjoint:
        //      s' = s + mu*Sy + my + z*exp(lstd) ----
        if (roleB) {
            const float g = gs[b_p * SD + b_d] + pre[PRE_RS * BL + b_p * SD + b_d];
            gsp[b_p * SD + b_d] = g;
            in[b_d * P + b_p] = g * cst[C_SY + b_d];
            if (dyn.has_density) in[(D + b_d) * P + b_p] = g * pre[PRE_FD * BL + b_p * SD + b_d];
        }
        if (has_sav) {
            mbar_wait(&sav_full[cur], spar[cur]);
            spar[cur] ^= 1u;
        }
        const float *sav = savb + (size_t)cur * prm.sav_floats;
        PMB_MARK(33);
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const NetSweep &net = which ? pol : dyn;
            int wpg = 0;
            net_backward<P>(prm, net, which == 1, sched_i, in, out, obuf, smem, red, sav, S, which ? nm_pol : nm_dyn, t,
                            n0, part, wpg);
            const float *ob = wpg ? part : obuf;
            const int owpg = wpg ? wpg : 1;
            PMB_MARK(34 + 3 * which);
            CTA_SYNC();
            PMB_MARK(35 + 3 * which);
            if (which == 0) {
                // ---- through the input scaler d[s;a] = dx * iSx, then (action dims) the tanh squash +
                //      policy density adjoint: a = scale*tanh(u)+bias, u = mu + z*exp(lstd) ----
                if (roleX) {
                    const float v = read_out(ob, nullptr, x_p, x_k, dyn.nin, owpg) * cst[C_ISX + x_k];
                    if (x_k < D) {
                        gsp[x_p * SD + x_k] += v;
                    } else {
                        const int u = x_k - D;
                        const float ga = pre[PRE_RA * BL + x_p * SD + u] + v;
                        const float du = ga * pre[PRE_TP * BL + x_p * SD + u];
                        out[u * P + x_p] = du;
                        float dls = 0.f;
                        if (pol.has_density) {
                            dls = du * pre[PRE_FP * BL + x_p * SD + u];
                            out[(U + u) * P + x_p] = dls;
                        }
                        if (n0 + x_p < N) {
                            if (prm.da_total) prm.da_total[((size_t)t * N + n0 + x_p) * U + u] = ga;
                            float *dd = prm.ws + pol.delta_off[pol.nlin - 1] + ((size_t)t * N + n0 + x_p) * pol.nout;
                            dd[u] = du;
                            if (pol.has_density) dd[U + u] = dls;
                        }
                    }
                }
                if (t > 0) precompute_a(pren);
                float *tmp = in; in = out; out = tmp;
            } else {
                // ---- dL/ds_t = carried + through dynamics input + through policy input + direct cotangent ----
                if (roleB)
                    gs[b_p * SD + b_d] = gsp[b_p * SD + b_d] + read_out(ob, nullptr, b_p, b_d, pol.nin, owpg) +
                                         pre[PRE_GS0 * BL + b_p * SD + b_d];
                if (t > 0) precompute_b(pren);
            }
            PMB_MARK(36 + 3 * which);
        }
        if (has_sav) {      // this step's stored activations are consumed: the buffer may be refilled
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&sav_empty[cur]);
        }
        CTA_SYNC();
        PMB_MARK(40);
        cur = nxt;
    }
    if (prm.dx0 && roleB && n0 + b_p < N) prm.dx0[(size_t)(n0 + b_p) * D + b_d] = gs[b_p * SD + b_d];
}

cudaError_t launch_rollout_bwd(const SweepParams &prm, int P, int smem_bytes, cudaStream_t stream) {
    const int grid = (prm.N + P - 1) / P;
    cudaError_t e;
#define PMB_LAUNCH_BWD(PP)                                                                                   \
    case PP:                                                                                                 \
        e = cudaFuncSetAttribute(rollout_bwd_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                 smem_bytes);                                                                \
        if (e != cudaSuccess) return e;                                                                      \
        if (prm.mm_states) {                                                                                 \
            void *args[] = {(void *)&prm};                                                                   \
            e = cudaLaunchCooperativeKernel((void *)rollout_bwd_kernel<PP>, dim3(grid), dim3(NT_LAUNCH), args,      \
                                            smem_bytes, stream);                                             \
            if (e != cudaSuccess) return e;                                                                  \
        } else {                                                                                             \
            rollout_bwd_kernel<PP><<<grid, NT_LAUNCH, smem_bytes, stream>>>(prm);                                   \
        }                                                                                                    \
        break;
    switch (P) {
        PMB_LAUNCH_BWD(1)
        PMB_LAUNCH_BWD(2)
        PMB_LAUNCH_BWD(4)
        PMB_LAUNCH_BWD(8)
        default:
            return cudaErrorInvalidValue;
    }
#undef PMB_LAUNCH_BWD
    return cudaGetLastError();
}

}  // namespace pmb
