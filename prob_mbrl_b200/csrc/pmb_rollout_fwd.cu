// Forward sweep of the imagined rollout: H steps of
//   policy MLP -> Gaussian action sample -> tanh squash -> dynamics MLP -> Gaussian state
//   sample -> reward
// for P particles per CTA, state tile resident in shared memory across the horizon.
// Replaces the loop body of utils.rollout (reference utils/rollout.py:93-163) with
// Policy.forward (models/core.py:221-248), DynamicsModel.forward (models/core.py:265-303),
// B/CDropout masks (models/modules.py:61,160), DiagGaussianDensity (models/densities.py:87-121)
// and the env reward (envs/cartpole/env.py:41-86 et al.).
#include "pmb_internal.cuh"
#include "pmb_mm.cuh"

namespace pmb {

// Hidden layers (wide) then output projection (narrow) of one net.  `in` holds the [K][P] input
// tile; on return `in` is the last hidden tile, obuf[p][nout] the raw outputs (visible after the
// caller's next barrier).
template <int P>
__device__ __forceinline__ void net_forward(const SweepParams &prm, const NetSweep &net, int &sched_i,
                                            float *&in, float *&out, float *obuf, float *smem, float *red,
                                            Stream &S, const NarrowMap &nm, int t, int n0, bool dbg_on, int mark0,
                                            float *part, int &wpg) {
    const int N = prm.N;
    const Lin &Lo = net.lin[net.nlin - 1];
#pragma unroll 1
    for (int l = 0; l + 1 < net.nlin; ++l) {
        const Lin &L = net.lin[l];
        const int npad = L.Npad;
        const float *bias_s = L.bias_soff >= 0 ? smem + L.bias_soff : nullptr;
        const float *mask_s = net.mask_soff[l] >= 0 ? smem + net.mask_soff[l] : nullptr;
        const float *mask_g = (!mask_s && net.mask_off[l] >= 0) ? prm.ws + net.mask_off[l] : nullptr;
        // relu, x * noise[:N] (modules.py:61,160), / p (BDropout only; multiplied by the fp32 reciprocal,
        // <= 1 ulp from the reference's division)
        const float inv_keep = 1.f / net.keep[l];
        float *sv = prm.ws + net.saved_off[l] + ((size_t)t * N + n0) * npad;
        if (!L.streamed) {
            // first layer: K = D or D+U
            thin_layer<P>(L, smem, in, [&](int j, float (&acc)[P]) {
                const float b = bias_s ? bias_s[j] : 0.f;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float x = acc[p] + b;
                    float mk = 1.f;
                    if (mask_s) mk = mask_s[p * npad + j];
                    else if (mask_g) mk = __ldg(mask_g + (size_t)min(n0 + p, N - 1) * npad + j);
                    x = (x < 0.f ? 0.f : x) * mk * inv_keep;
                    out[j * P + p] = x;
                    if (n0 + p < N) sv[(size_t)p * npad + j] = x;
                }
            });
            for (int j = L.Nout + threadIdx.x; j < npad; j += NT) {
#pragma unroll
                for (int p = 0; p < P; ++p) out[j * P + p] = 0.f;
            }
        } else {
            WideMap m;
            m.set(npad);
            const int col = 4 * m.cq;
            float *dst = out + col * P;
            // the last hidden layer also forms its share of the output projection (narrow_fused)
            const bool fuse = (l + 2 == net.nlin);
            const int cqn = npad >> 2;
            wpg = fuse ? ((cqn <= 32 ? 32 : cqn <= 64 ? 64 : cqn <= 128 ? 128 : 256) >> 5) : 1;
            const int wig = (threadIdx.x >> 5) & (wpg - 1);
            wide_layer<P>(L, &prm.sched[sched_i], smem, in, red, S, m, [&](int p, float4 v, bool act) {
              if (act) {
                if (bias_s) {
                    const float4 b = *reinterpret_cast<const float4 *>(bias_s + col);
                    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                }
                float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
                if (mask_s) mk = *reinterpret_cast<const float4 *>(mask_s + p * npad + col);
                else if (mask_g) mk = __ldg(reinterpret_cast<const float4 *>(mask_g + (size_t)min(n0 + p, N - 1) * npad + col));
                v.x = (v.x < 0.f ? 0.f : v.x) * mk.x * inv_keep;
                v.y = (v.y < 0.f ? 0.f : v.y) * mk.y * inv_keep;
                v.z = (v.z < 0.f ? 0.f : v.z) * mk.z * inv_keep;
                v.w = (v.w < 0.f ? 0.f : v.w) * mk.w * inv_keep;
                dst[p] = v.x;
                dst[P + p] = v.y;
                dst[2 * P + p] = v.z;
                dst[3 * P + p] = v.w;
                if (n0 + p < N) *reinterpret_cast<float4 *>(sv + (size_t)p * npad + col) = v;
              }
              if (fuse) narrow_fused(v, act, p, col, smem + Lo.soff, Lo.K, Lo.Nout, part, wpg, wig);
            }, dbg_on ? prm.dbg + 64 + 12 * (mark0 + l) : nullptr);
            ++sched_i;
            if (fuse) {
                float *tmp = in; in = out; out = tmp;
                PMB_MARK(mark0 + l);
                return;                     // outputs are in `part` (wpg partials each), bias not yet added
            }
        }
        float *tmp = in; in = out; out = tmp;
        PMB_MARK(mark0 + l);
    }
    wpg = 0;                            // plain buffer: narrow_layer adds the bias itself
    narrow_layer<P>(Lo, nm, smem, in, obuf, Lo.bias_soff >= 0 ? smem + Lo.bias_soff : nullptr, red);
    PMB_MARK(mark0 + net.nlin - 1);
}

template <int P>
__global__ void __launch_bounds__(NT_LAUNCH, 1) rollout_fwd_kernel(const __grid_constant__ SweepParams prm) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) RingBars ring;
    __shared__ __align__(16) ChunkDesc chunk_tab[MAXCHUNKS];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * P;
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H;
    // tiles and scratch start finite; done by all threads BEFORE the producer may issue any TMA write
    for (int i = tid; i < prm.off_stage; i += NT_LAUNCH) smem[i] = 0.f;
    // ---- barriers, then the producer warp peels off: it only feeds the weight ring ----
    if (tid == 0) {
        for (int s = 0; s < prm.nstages; ++s) {
            mbar_init(&ring.full[s], 1);
            mbar_init(&ring.empty[s], NWARP);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();                       // the only barrier all 288 threads take
    if (tid >= NT) {
        if (tid == NT && prm.chunks_per_step > 0) {
            fence_proxy_async();
            ring_fill_table(prm, chunk_tab);
            RingProducer rp;
            rp.init();
            rp.issue(prm, smem + prm.off_stage, &ring, chunk_tab, H * prm.chunks_per_step);
        }
        return;
    }
    float *cst = smem + prm.off_cst;
    float *act0 = smem + prm.off_act0;
    float *act1 = smem + prm.off_act1;
    float *red = smem + prm.off_red;
    float *obuf = smem + prm.off_misc;
    float *part = obuf + 320 * P;         // partials of the fused output projection

    // ---- thread roles for the per-particle stages (fixed for the whole horizon) ----
    const bool roleA = tid < P * U;                       // one (particle, action dim)
    const int a_p = roleA ? tid / U : 0, a_u = roleA ? tid - a_p * U : 0;
    const int a_n = min(n0 + a_p, N - 1);
    const bool roleB = tid >= 128 && tid - 128 < P * D;   // one (particle, state dim)
    const int b_p = roleB ? (tid - 128) / D : 0, b_d = roleB ? (tid - 128) - b_p * D : 0;
    const int b_n = min(n0 + b_p, N - 1);
    // raw outputs kept for the reverse sweep: thread -> (particle, output) of the policy / dynamics net
    const int op_p = tid / prm.pol.nout, op_j = tid - op_p * prm.pol.nout;
    const int od_p = tid / prm.dyn.nout, od_j = tid - od_p * prm.dyn.nout;

    load_constants(prm, cst);
    load_resident(prm, smem, n0);
    float s_reg = 0.f;            // role B: this thread's element of the current state (never leaves registers)
    if (roleB) {
        s_reg = prm.x0[(size_t)b_n * D + b_d];
        act0[b_d * P + b_p] = s_reg;
        if (n0 + b_p < N) prm.states[(size_t)b_n * D + b_d] = s_reg;
    }
    const float elmax_pol = expf(prm.pol.lmax), elmax_dyn = expf(prm.dyn.lmax);
    float zA = 0.f, zB = 0.f;
    if (roleA && prm.pol.has_density) zA = __ldg(prm.pol.z + (size_t)a_n * U + a_u);
    if (roleB && prm.dyn.has_density) zB = __ldg(prm.dyn.z + (size_t)b_n * D + b_d);
    NarrowMap nm_pol, nm_dyn;
    nm_pol.set<P>(prm.pol.lin[prm.pol.nlin - 1]);
    nm_dyn.set<P>(prm.dyn.lin[prm.dyn.nlin - 1]);
    Stream S;
    S.init(&prm, smem, &ring);
    MMSmem mmS;
    MMGroup grp;
    unsigned epoch = 0;
    if (prm.mm_states) {
        mmS.carve<P>(smem + prm.off_mm);
        grp.set<P>(prm, n0);
    }
    CTA_SYNC();

#pragma unroll 1
    for (int t = 0; t < H; ++t) {
        const bool dbg_on = prm.dbg != nullptr && blockIdx.x == 0 && tid == 0 && t == H / 2;
        PMB_MARK(0);
        int sched_i = 0;
        float *in = act0, *out = act1;
        // per-step noise (only when the caller pre-drew [H, N, .] tables): issue the loads early
        if (prm.pol.zstride != 0 && roleA && prm.pol.has_density)
            zA = __ldg(prm.pol.z + (size_t)t * prm.pol.zstride + (size_t)a_n * U + a_u);
        if (prm.dyn.zstride != 0 && roleB && prm.dyn.has_density)
            zB = __ldg(prm.dyn.z + (size_t)t * prm.dyn.zstride + (size_t)b_n * D + b_d);
        float zrow = 0.f;      // moment matching: this particle's row of z_mm, rotated by the step index
        if (prm.mm_states && roleB) zrow = __ldg(prm.z_mm + (size_t)((t + n0 + b_p) % N) * D + b_d);
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const NetSweep &net = which ? prm.dyn : prm.pol;
            int wpg = 0;
            net_forward<P>(prm, net, sched_i, in, out, obuf, smem, red, S, which ? nm_dyn : nm_pol, t, n0, dbg_on,
                           1 + 8 * which, part, wpg);
            CTA_SYNC();
            PMB_MARK(7 + 8 * which);
            // raw outputs of the net: fused path = partials + bias, fallback = finished values in obuf
            const Lin &Lout = net.lin[net.nlin - 1];
            const float *ob = wpg ? part : obuf;
            const float *obias = (wpg && Lout.bias_soff >= 0) ? smem + Lout.bias_soff : nullptr;
            const int owpg = wpg ? wpg : 1;
            if (which == 0) {
                // ---- Gaussian action sample + tanh squash (densities.py:95-119, core.py:243);
                //      dynamics input (core.py:269,177) ----
                if (roleA) {
                    float uu = read_out(ob, obias, a_p, a_u, net.nout, owpg);
                    if (net.has_density)
                        uu += zA * exp_clamped_logstd(read_out(ob, obias, a_p, U + a_u, net.nout, owpg), net.lmax, elmax_pol);
                    const float a = cst[C_SCALE + a_u] * tanhf(uu) + cst[C_BIAS + a_u];
                    if (n0 + a_p < N) prm.actions[((size_t)t * N + a_n) * U + a_u] = a;
                    out[(D + a_u) * P + a_p] = (a - cst[C_MX + D + a_u]) * cst[C_ISX + D + a_u];
                }
                if (roleB) out[b_d * P + b_p] = (s_reg - cst[C_MX + b_d]) * cst[C_ISX + b_d];
                if (tid < P * net.nout && n0 + op_p < N)
                    prm.ws[net.outsaved_off + ((size_t)t * N + n0 + op_p) * net.nout + op_j] =
                        read_out(ob, obias, op_p, op_j, net.nout, owpg);
                float *tmp = in; in = out; out = tmp;
            } else {
                // ---- Gaussian state sample, s' = s + delta (densities.py:100-119, core.py:293,298);
                //      it is also the next step's policy input ----
                if (roleB) {
                    const float sy = cst[C_SY + b_d], my = cst[C_MY + b_d];
                    float delta;
                    if (net.has_density) {
                        const float mu = read_out(ob, obias, b_p, b_d, net.nout, owpg);
                        // exp(clamped log-std + log Sy) = Sy * exp(clamped log-std)   (densities.py:105)
                        const float sd = sy * exp_clamped_logstd(read_out(ob, obias, b_p, D + b_d, net.nout, owpg),
                                                                 net.lmax, elmax_dyn);
                        delta = (mu * sy + my) + zB * sd;
                    } else {
                        delta = read_out(ob, obias, b_p, b_d, net.nout, owpg) * sy + my;
                    }
                    s_reg += delta;
                }
                if (prm.mm_states)      // rollout.py:121-132
                    mm_states_forward<P>(prm, mmS, grp, t, n0, roleB, b_p, b_d, b_n, s_reg, zrow, epoch);
                if (roleB) {
                    act0[b_d * P + b_p] = s_reg;
                    if (n0 + b_p < N) prm.states[((size_t)(t + 1) * N + b_n) * D + b_d] = s_reg;
                }
                if (tid < P * net.nout && n0 + od_p < N)
                    prm.ws[net.outsaved_off + ((size_t)t * N + n0 + od_p) * net.nout + od_j] =
                        read_out(ob, obias, od_p, od_j, net.nout, owpg);
            }
            PMB_MARK(8 + 8 * which);
        }
    }
    // ---- rewards r_t = scale*exp(-0.5*(d^T Q d + a^T R a)) + offset on (s_{t+1}, a_t) for every step
    //      (envs/cartpole/env.py:62-86).  Nothing in the recurrence consumes them, so they are
    //      evaluated here, off the serial chain, from the trajectory this CTA just wrote. ----
    CTA_SYNC();
    for (int i = tid; i < H * P; i += NT) {
        const int tt = i / P, p = i - tt * P;
        if (n0 + p >= N) continue;
        // the reward sees the next state BEFORE moment matching (models/core.py:293 runs inside dynamics())
        const float *s1 = prm.mm_states ? prm.s1pre + ((size_t)tt * N + n0 + p) * D
                                        : prm.states + ((size_t)(tt + 1) * N + n0 + p) * D;
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
}

cudaError_t launch_rollout_fwd(const SweepParams &prm, int P, int smem_bytes, cudaStream_t stream) {
    const int grid = (prm.N + P - 1) / P;
    cudaError_t e;
    // moment matching synchronises the whole grid every step: cooperative launch guarantees co-residency
#define PMB_LAUNCH_FWD(PP)                                                                                   \
    case PP:                                                                                                 \
        e = cudaFuncSetAttribute(rollout_fwd_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                 smem_bytes);                                                                \
        if (e != cudaSuccess) return e;                                                                      \
        if (prm.mm_states) {                                                                                 \
            void *args[] = {(void *)&prm};                                                                   \
            e = cudaLaunchCooperativeKernel((void *)rollout_fwd_kernel<PP>, dim3(grid), dim3(NT_LAUNCH), args,      \
                                            smem_bytes, stream);                                             \
            if (e != cudaSuccess) return e;                                                                  \
        } else {                                                                                             \
            rollout_fwd_kernel<PP><<<grid, NT_LAUNCH, smem_bytes, stream>>>(prm);                                   \
        }                                                                                                    \
        break;
    switch (P) {
        PMB_LAUNCH_FWD(1)
        PMB_LAUNCH_FWD(2)
        PMB_LAUNCH_FWD(4)
        PMB_LAUNCH_FWD(8)
        default:
            return cudaErrorInvalidValue;
    }
#undef PMB_LAUNCH_FWD
    return cudaGetLastError();
}

}  // namespace pmb
