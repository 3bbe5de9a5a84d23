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

constexpr int SD = PMB_MAX_STATE;   // row stride of the small per-particle buffers

// Hidden layers (wide) then output projection (narrow) of one net.  `in` holds the [K][P] input
// tile; on return `in` is the last hidden tile, obuf[p][nout] the raw outputs (visible after the
// caller's next barrier).
template <int P>
__device__ __forceinline__ void net_forward(const SweepParams &prm, const NetSweep &net, int &sched_i,
                                            float *&in, float *&out, float *obuf, const float *res,
                                            float *red, Stream &S, int t, int n0) {
    const int N = prm.N;
    for (int l = 0; l + 1 < net.nlin; ++l) {
        const Lin &L = net.lin[l];
        WideMap m;
        m.set(L.Npad);
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 mk[P];
        const bool epi = m.active && m.g == 0;
        if (epi) {
            if (L.boff >= 0) bv = __ldg(reinterpret_cast<const float4 *>(prm.ws + L.boff) + m.cq);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                int n = min(n0 + p, N - 1);
                mk[p] = net.mask_off[l] >= 0
                            ? __ldg(reinterpret_cast<const float4 *>(prm.ws + net.mask_off[l] + (size_t)n * L.Npad) + m.cq)
                            : make_float4(1.f, 1.f, 1.f, 1.f);
            }
        }
        float acc[P][4];
        wide_layer<P>(acc, L, L.streamed ? &prm.sched[sched_i] : nullptr, res, in, red, S, m);
        if (L.streamed) ++sched_i;
        if (epi) {
            const float keep = net.keep[l];
            float *sv = prm.ws + net.saved_off[l];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float v[4] = {acc[p][0] + bv.x, acc[p][1] + bv.y, acc[p][2] + bv.z, acc[p][3] + bv.w};
                const float mm[4] = {mk[p].x, mk[p].y, mk[p].z, mk[p].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x = v[j] < 0.f ? 0.f : v[j];   // relu (NaN propagates like torch)
                    x = x * mm[j];                        // x * noise[:N]         (modules.py:61,160)
                    if (keep != 1.f) x = x / keep;        // ... / p  (BDropout only)
                    v[j] = x;
                    out[(4 * m.cq + j) * P + p] = x;
                }
                if (n0 + p < N)
                    *reinterpret_cast<float4 *>(sv + ((size_t)t * N + n0 + p) * L.Npad + 4 * m.cq) =
                        make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        float *tmp = in; in = out; out = tmp;
    }
    const Lin &Lo = net.lin[net.nlin - 1];
    narrow_layer<P>(Lo, res, in, obuf, Lo.boff >= 0 ? prm.ws + Lo.boff : nullptr);
}

template <int P>
__global__ void __launch_bounds__(NT, 1) rollout_fwd_kernel(const __grid_constant__ SweepParams prm) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bars[MAXS];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * P;
    const int N = prm.N, D = prm.D, U = prm.U, H = prm.H;
    float *res = smem;
    float *act0 = smem + prm.off_act0;
    float *act1 = smem + prm.off_act1;
    float *red = smem + prm.off_red;
    float *misc = smem + prm.off_misc;
    float *s_cur = misc, *s_nxt = misc + P * SD, *abuf = misc + 2 * P * SD, *obuf = misc + 3 * P * SD;

    load_resident(prm, res);
    for (int i = tid; i < P * D; i += NT) {
        int p = i / D, d = i - p * D;
        int n = min(n0 + p, N - 1);
        float v = prm.x0[(size_t)n * D + d];
        s_cur[p * SD + d] = v;
        if (n0 + p < N) prm.states[(size_t)n * D + d] = v;
    }
    Stream S;
    S.init(&prm, smem, bars);
    __syncthreads();

    const NetSweep &pol = prm.pol;
    const NetSweep &dyn = prm.dyn;
    for (int t = 0; t < H; ++t) {
        int sched_i = 0;
        float *in = act0, *out = act1;
        // ---- policy input tile ----
        for (int i = tid; i < P * D; i += NT) {
            int p = i / D, d = i - p * D;
            in[d * P + p] = s_cur[p * SD + d];
        }
        net_forward<P>(prm, pol, sched_i, in, out, obuf, res, red, S, t, n0);
        __syncthreads();
        // ---- Gaussian action sample + tanh squash (densities.py:95-119, core.py:243);
        //      dynamics input (core.py:269,177) ----
        for (int i = tid; i < P * U; i += NT) {
            int p = i / U, u = i - p * U;
            int n = min(n0 + p, N - 1);
            float uu;
            if (pol.has_density) {
                float mu = obuf[p * pol.nout + u];
                float ls = clamp_logstd(obuf[p * pol.nout + U + u], pol.lmax);
                float z = __ldg(pol.z + (size_t)t * pol.zstride + (size_t)n * U + u);
                uu = mu + z * expf(ls);
            } else {
                uu = obuf[p * pol.nout + u];
            }
            float a = __ldg(prm.act_scale + u) * tanhf(uu) + __ldg(prm.act_bias + u);
            abuf[p * SD + u] = a;
            if (n0 + p < N) prm.actions[((size_t)t * N + n) * U + u] = a;
            out[(D + u) * P + p] = (a - __ldg(prm.mx + D + u)) * __ldg(prm.iSx + D + u);
        }
        for (int i = tid; i < P * D; i += NT) {
            int p = i / D, d = i - p * D;
            out[d * P + p] = (s_cur[p * SD + d] - __ldg(prm.mx + d)) * __ldg(prm.iSx + d);
        }
        for (int i = tid; i < P * pol.nout; i += NT) {
            int p = i / pol.nout, j = i - p * pol.nout;
            if (n0 + p < N) prm.ws[pol.outsaved_off + ((size_t)t * N + n0 + p) * pol.nout + j] = obuf[i];
        }
        {
            float *tmp = in; in = out; out = tmp;
        }
        net_forward<P>(prm, dyn, sched_i, in, out, obuf, res, red, S, t, n0);
        __syncthreads();
        // ---- Gaussian state sample, s' = s + delta (densities.py:100-119, core.py:293,298) ----
        for (int i = tid; i < P * D; i += NT) {
            int p = i / D, d = i - p * D;
            int n = min(n0 + p, N - 1);
            float sy = __ldg(prm.Sy + d), my = __ldg(prm.my + d);
            float delta;
            if (dyn.has_density) {
                float mu = obuf[p * dyn.nout + d];
                float ls = clamp_logstd(obuf[p * dyn.nout + D + d], dyn.lmax) + logf(sy);
                float z = __ldg(dyn.z + (size_t)t * dyn.zstride + (size_t)n * D + d);
                delta = (mu * sy + my) + z * expf(ls);
            } else {
                delta = obuf[p * dyn.nout + d] * sy + my;
            }
            float s1 = s_cur[p * SD + d] + delta;
            s_nxt[p * SD + d] = s1;
            if (n0 + p < N) prm.states[((size_t)(t + 1) * N + n) * D + d] = s1;
        }
        for (int i = tid; i < P * dyn.nout; i += NT) {
            int p = i / dyn.nout, j = i - p * dyn.nout;
            if (n0 + p < N) prm.ws[dyn.outsaved_off + ((size_t)t * N + n0 + p) * dyn.nout + j] = obuf[i];
        }
        __syncthreads();
        // ---- reward on (s', a) (envs/cartpole/env.py:62-86) ----
        if (tid < P && n0 + tid < N) {
            const int p = tid;
            float dl[PMB_MAX_REWARD_ROWS];
            for (int i = 0; i < prm.KR; ++i) {
                float s = __ldg(prm.rew_c0 + i);
                for (int d = 0; d < D; ++d) s = fmaf(__ldg(prm.rew_C + i * D + d), s_nxt[p * SD + d], s);
                dl[i] = s;
            }
            float cost = 0.f;
            for (int i = 0; i < prm.KR; ++i) {
                float q = 0.f;
                for (int j = 0; j < prm.KR; ++j) q = fmaf(dl[j], __ldg(prm.rew_Q + j * prm.KR + i), q);
                cost = fmaf(q, dl[i], cost);
            }
            for (int u = 0; u < U; ++u) {
                float q = 0.f;
                for (int v = 0; v < U; ++v) q = fmaf(abuf[p * SD + v], __ldg(prm.rew_R + v * U + u), q);
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
