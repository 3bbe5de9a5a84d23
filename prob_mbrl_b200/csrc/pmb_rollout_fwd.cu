// Forward sweep of the imagined rollout: H steps of
//   policy MLP -> Gaussian action sample -> tanh squash -> dynamics MLP -> Gaussian state
//   sample -> reward
// for P particles per CTA, state tile resident in shared memory across the horizon.
// Replaces the loop body of utils.rollout (reference utils/rollout.py:93-163) with
// Policy.forward (models/core.py:221-248), DynamicsModel.forward (models/core.py:265-303),
// B/CDropout masks (models/modules.py:61,160), DiagGaussianDensity (models/densities.py:87-121)
// and the env reward (envs/cartpole/env.py:41-86 et al.).
#include "pmb_internal.cuh"

namespace pmb {

// Hidden layers (wide) then output projection (narrow) of one net.  `in` holds the [K][P] input
// tile; on return `in` is the last hidden tile, obuf[p][nout] the raw outputs (visible after the
// caller's next barrier).
template <int P>
__device__ __forceinline__ void net_forward(const SweepParams &prm, const NetSweep &net, int &sched_i,
                                            float *&in, float *&out, float *obuf, float *smem, float *red,
                                            Stream &S, int t, int n0) {
    const int N = prm.N;
    for (int l = 0; l + 1 < net.nlin; ++l) {
        const Lin &L = net.lin[l];
        WideMap m;
        m.set(L.Npad);
        const int col = 4 * m.cq;
        const float *bias_s = L.bias_soff >= 0 ? smem + L.bias_soff + col : nullptr;
        const float *mask_s = net.mask_soff[l] >= 0 ? smem + net.mask_soff[l] + col : nullptr;
        const float *mask_g = (!mask_s && net.mask_off[l] >= 0) ? prm.ws + net.mask_off[l] + col : nullptr;
        const float keep = net.keep[l];
        float *sv = prm.ws + net.saved_off[l] + ((size_t)t * N + n0) * L.Npad + col;
        float *dst = out + col * P;
        const int npad = L.Npad;
        wide_layer<P>(L, L.streamed ? &prm.sched[sched_i] : nullptr, smem, in, red, S, m, [&](int p, float4 v) {
            if (bias_s) {
                const float4 b = *reinterpret_cast<const float4 *>(bias_s);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);
            if (mask_s) mk = *reinterpret_cast<const float4 *>(mask_s + p * npad);
            else if (mask_g) mk = __ldg(reinterpret_cast<const float4 *>(mask_g + (size_t)min(n0 + p, N - 1) * npad));
            // relu (NaN propagates like torch), x * noise[:N] (modules.py:61,160), ... / p (BDropout only)
            v.x = (v.x < 0.f ? 0.f : v.x) * mk.x;
            v.y = (v.y < 0.f ? 0.f : v.y) * mk.y;
            v.z = (v.z < 0.f ? 0.f : v.z) * mk.z;
            v.w = (v.w < 0.f ? 0.f : v.w) * mk.w;
            if (keep != 1.f) {
                v.x = v.x / keep; v.y = v.y / keep; v.z = v.z / keep; v.w = v.w / keep;
            }
            dst[p] = v.x;
            dst[P + p] = v.y;
            dst[2 * P + p] = v.z;
            dst[3 * P + p] = v.w;
            if (n0 + p < N) *reinterpret_cast<float4 *>(sv + (size_t)p * npad) = v;
        });
        if (L.streamed) ++sched_i;
        float *tmp = in; in = out; out = tmp;
    }
    const Lin &Lo = net.lin[net.nlin - 1];
    narrow_layer<P>(Lo, smem, in, obuf, Lo.bias_soff >= 0 ? smem + Lo.bias_soff : nullptr);
}

template <int P>
__global__ void __launch_bounds__(NT, 1) rollout_fwd_kernel(const __grid_constant__ SweepParams prm) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bars[MAXS];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * P;
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H;
    const NetSweep &pol = prm.pol;
    const NetSweep &dyn = prm.dyn;
    float *cst = smem + prm.off_cst;
    float *act0 = smem + prm.off_act0;
    float *act1 = smem + prm.off_act1;
    float *red = smem + prm.off_red;
    float *misc = smem + prm.off_misc;
    float *s_cur = misc, *s_nxt = misc + P * SD, *abuf = misc + 2 * P * SD, *obuf = misc + 3 * P * SD;

    // ---- thread roles for the per-particle stages (fixed for the whole horizon) ----
    const bool roleA = tid < P * U;                       // one (particle, action dim)
    const int a_p = roleA ? tid / U : 0, a_u = roleA ? tid - a_p * U : 0;
    const int a_n = min(n0 + a_p, N - 1);
    const bool roleB = tid >= 128 && tid - 128 < P * D;   // one (particle, state dim)
    const int b_p = roleB ? (tid - 128) / D : 0, b_d = roleB ? (tid - 128) - b_p * D : 0;
    const int b_n = min(n0 + b_p, N - 1);
    const bool roleOp = tid < P * pol.nout;               // raw policy outputs to keep for the reverse sweep
    const int op_p = roleOp ? tid / pol.nout : 0, op_j = roleOp ? tid - op_p * pol.nout : 0;
    const bool roleOd = tid < P * dyn.nout;
    const int od_p = roleOd ? tid / dyn.nout : 0, od_j = roleOd ? tid - od_p * dyn.nout : 0;

    load_constants(prm, cst);
    load_resident(prm, smem, n0);
    if (roleB) {
        const float v = prm.x0[(size_t)b_n * D + b_d];
        s_cur[b_p * SD + b_d] = v;
        if (n0 + b_p < N) prm.states[(size_t)b_n * D + b_d] = v;
    }
    float zA = 0.f, zB = 0.f;
    if (roleA && pol.has_density) zA = __ldg(pol.z + (size_t)a_n * U + a_u);
    if (roleB && dyn.has_density) zB = __ldg(dyn.z + (size_t)b_n * D + b_d);
    Stream S;
    S.init(&prm, smem, bars);
    __syncthreads();

    for (int t = 0; t < H; ++t) {
        int sched_i = 0;
        float *in = act0, *out = act1;
        // per-step noise (only when the caller pre-drew [H, N, .] tables): issue the loads early
        if (pol.zstride != 0 && roleA && pol.has_density) zA = __ldg(pol.z + (size_t)t * pol.zstride + (size_t)a_n * U + a_u);
        if (dyn.zstride != 0 && roleB && dyn.has_density) zB = __ldg(dyn.z + (size_t)t * dyn.zstride + (size_t)b_n * D + b_d);
        // ---- policy input tile ----
        if (roleB) in[b_d * P + b_p] = s_cur[b_p * SD + b_d];
        net_forward<P>(prm, pol, sched_i, in, out, obuf, smem, red, S, t, n0);
        __syncthreads();
        // ---- Gaussian action sample + tanh squash (densities.py:95-119, core.py:243);
        //      dynamics input (core.py:269,177) ----
        if (roleA) {
            float uu;
            if (pol.has_density) {
                const float mu = obuf[a_p * pol.nout + a_u];
                const float ls = clamp_logstd(obuf[a_p * pol.nout + U + a_u], pol.lmax);
                uu = mu + zA * expf(ls);
            } else {
                uu = obuf[a_p * pol.nout + a_u];
            }
            const float a = cst[C_SCALE + a_u] * tanhf(uu) + cst[C_BIAS + a_u];
            abuf[a_p * SD + a_u] = a;
            if (n0 + a_p < N) prm.actions[((size_t)t * N + a_n) * U + a_u] = a;
            out[(D + a_u) * P + a_p] = (a - cst[C_MX + D + a_u]) * cst[C_ISX + D + a_u];
        }
        if (roleB) out[b_d * P + b_p] = (s_cur[b_p * SD + b_d] - cst[C_MX + b_d]) * cst[C_ISX + b_d];
        if (roleOp && n0 + op_p < N)
            prm.ws[pol.outsaved_off + ((size_t)t * N + n0 + op_p) * pol.nout + op_j] = obuf[tid];
        {
            float *tmp = in; in = out; out = tmp;
        }
        net_forward<P>(prm, dyn, sched_i, in, out, obuf, smem, red, S, t, n0);
        __syncthreads();
        // ---- Gaussian state sample, s' = s + delta (densities.py:100-119, core.py:293,298) ----
        if (roleB) {
            const float sy = cst[C_SY + b_d], my = cst[C_MY + b_d];
            float delta;
            if (dyn.has_density) {
                const float mu = obuf[b_p * dyn.nout + b_d];
                const float ls = clamp_logstd(obuf[b_p * dyn.nout + D + b_d], dyn.lmax) + cst[C_LSY + b_d];
                delta = (mu * sy + my) + zB * expf(ls);
            } else {
                delta = obuf[b_p * dyn.nout + b_d] * sy + my;
            }
            const float s1 = s_cur[b_p * SD + b_d] + delta;
            s_nxt[b_p * SD + b_d] = s1;
            if (n0 + b_p < N) prm.states[((size_t)(t + 1) * N + b_n) * D + b_d] = s1;
        }
        if (roleOd && n0 + od_p < N)
            prm.ws[dyn.outsaved_off + ((size_t)t * N + n0 + od_p) * dyn.nout + od_j] = obuf[tid];
        __syncthreads();
        // ---- reward on (s', a) (envs/cartpole/env.py:62-86) ----
        if (tid < P && n0 + tid < N) {
            const int p = tid;
            float dl[PMB_MAX_REWARD_ROWS];
            for (int i = 0; i < prm.KR; ++i) {
                float s = cst[C_C0 + i];
                for (int d = 0; d < D; ++d) s = fmaf(cst[C_C + i * SD + d], s_nxt[p * SD + d], s);
                dl[i] = s;
            }
            float cost = 0.f;
            for (int i = 0; i < prm.KR; ++i) {
                float q = 0.f;
                for (int j = 0; j < prm.KR; ++j) q = fmaf(dl[j], cst[C_Q + j * 4 + i], q);
                cost = fmaf(q, dl[i], cost);
            }
            for (int u = 0; u < U; ++u) {
                float q = 0.f;
                for (int v = 0; v < U; ++v) q = fmaf(abuf[p * SD + v], cst[C_R + v * SD + u], q);
                cost = fmaf(q, abuf[p * SD + u], cost);
            }
            prm.rewards[(size_t)t * N + n0 + p] = prm.rew_scale * expf(-0.5f * cost) + prm.rew_offset;
        }
        {
            float *tmp = s_cur; s_cur = s_nxt; s_nxt = tmp;
        }
    }
}

cudaError_t launch_rollout_fwd(const SweepParams &prm, int P, int smem_bytes, cudaStream_t stream) {
    const int grid = (prm.N + P - 1) / P;
    cudaError_t e;
#define PMB_LAUNCH_FWD(PP)                                                                                   \
    case PP:                                                                                                 \
        e = cudaFuncSetAttribute(rollout_fwd_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                 smem_bytes);                                                                \
        if (e != cudaSuccess) return e;                                                                      \
        rollout_fwd_kernel<PP><<<grid, NT, smem_bytes, stream>>>(prm);                                       \
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
