// Reverse sweep of the imagined rollout (back-propagation through time), no recompute:
// consumes the activations the forward sweep stored, walks the steps t = H-1 .. 0 and produces
//   * dL/dx0,
//   * the per-layer output adjoints ("deltas") of the POLICY net for every (t, particle), which the
//     batched weight-gradient GEMMs (pmb_wgrad.cu) contract afterwards over the (H*N) axis.
// Dynamics weight gradients are never formed (the reference computes and discards them,
// SURVEY.md App. D-4).  Replaces loss.backward() through utils.rollout (reference
// algorithms/mc_pilco.py:197); adjoint formulas are those of oracle/rollout_oracle.py::manual_backward,
// which tests/test_oracle_backward.py checks against autograd.
#include "pmb_internal.cuh"

namespace pmb {

constexpr int SD = PMB_MAX_STATE;

// Backward through one net.  `in` holds the adjoint of the net's raw outputs as a [nout][P] tile.
// Wide layers l = nlin-1 .. 1 (weights W_l as stored, [out][in]) produce the adjoint of hidden l-1,
// gated by the stored activation; the final narrow layer (W_0^T) leaves d(input)[p][nin] in obuf.
// If `store_delta`, every linear layer's output adjoint is written to global for the weight gradient.
template <int P, bool kStoreDelta>
__device__ __forceinline__ void net_backward(const SweepParams &prm, const NetSweep &net, int &sched_i,
                                             float *&in, float *&out, float *obuf, const float *res,
                                             float *red, Stream &S, int t, int n0) {
    const int N = prm.N;
    for (int l = net.nlin - 1; l >= 1; --l) {
        const Lin &L = net.lin[l];      // wide: K = outputs of linear l (padded), Npad = width of hidden l-1
        const int h = l - 1;            // hidden layer whose adjoint we produce
        WideMap m;
        m.set(L.Npad);
        const bool epi = m.active && m.g == 0;
        float4 mk[P], sv[P];
        if (epi) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                int n = min(n0 + p, N - 1);
                mk[p] = net.mask_off[h] >= 0
                            ? __ldg(reinterpret_cast<const float4 *>(prm.ws + net.mask_off[h] + (size_t)n * L.Npad) + m.cq)
                            : make_float4(1.f, 1.f, 1.f, 1.f);
                sv[p] = __ldg(reinterpret_cast<const float4 *>(prm.ws + net.saved_off[h] +
                                                               ((size_t)t * N + n) * L.Npad) + m.cq);
            }
        }
        float acc[P][4];
        wide_layer<P>(acc, L, L.streamed ? &prm.sched[sched_i] : nullptr, res, in, red, S, m);
        if (L.streamed) ++sched_i;
        if (epi) {
            const float keep = net.keep[h];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float mm[4] = {mk[p].x, mk[p].y, mk[p].z, mk[p].w};
                const float hh[4] = {sv[p].x, sv[p].y, sv[p].z, sv[p].w};
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // y = relu(pre) * mask / keep  =>  dpre = (dy / keep) * mask * [pre > 0];
                    // y != 0 <=> pre > 0 and mask != 0
                    float x = acc[p][j];
                    if (keep != 1.f) x = x / keep;
                    x = hh[j] != 0.f ? x * mm[j] : 0.f;
                    v[j] = x;
                    out[(4 * m.cq + j) * P + p] = x;
                }
                if (kStoreDelta && n0 + p < N)
                    *reinterpret_cast<float4 *>(prm.ws + net.delta_off[h] + ((size_t)t * N + n0 + p) * L.Npad +
                                                4 * m.cq) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        float *tmp = in; in = out; out = tmp;
    }
    narrow_layer<P>(net.lin[0], res, in, obuf, nullptr);
}

template <int P>
__global__ void __launch_bounds__(NT, 1) rollout_bwd_kernel(const __grid_constant__ SweepParams prm) {
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
    float *gs = misc;                 // [P][SD] dL/ds_{t+1}
    float *gsp = misc + P * SD;       // [P][SD] dL/ds_t under construction
    float *ga = misc + 2 * P * SD;    // [P][SD] dL/da_t
    float *obuf = misc + 3 * P * SD;  // [P][<=SD] narrow outputs

    load_resident(prm, res);
    for (int i = tid; i < P * D; i += NT) {
        int p = i / D, d = i - p * D;
        int n = min(n0 + p, N - 1);
        gs[p * SD + d] = prm.g_states ? __ldg(prm.g_states + ((size_t)H * N + n) * D + d) : 0.f;
    }
    Stream S;
    S.init(&prm, smem, bars);
    __syncthreads();

    const NetSweep &pol = prm.pol;
    const NetSweep &dyn = prm.dyn;
    for (int t = H - 1; t >= 0; --t) {
        int sched_i = 0;
        float *in = act0, *out = act1;
        // ---- reward adjoint: r = scale*exp(-0.5*(d^T Q d + a^T R a)) + offset ----
        if (tid < P) {
            const int p = tid;
            const int n = min(n0 + p, N - 1);
            float gr = prm.g_rewards ? __ldg(prm.g_rewards + (size_t)t * N + n) : 0.f;
            float e = __ldg(prm.rewards + (size_t)t * N + n) - prm.rew_offset;
            float w = -0.5f * gr * e;
            const float *s1 = prm.states + ((size_t)(t + 1) * N + n) * D;
            const float *a = prm.actions + ((size_t)t * N + n) * U;
            float dl[PMB_MAX_REWARD_ROWS], qd[PMB_MAX_REWARD_ROWS];
            for (int i = 0; i < prm.KR; ++i) {
                float s = __ldg(prm.rew_c0 + i);
                for (int d = 0; d < D; ++d) s = fmaf(__ldg(prm.rew_C + i * D + d), __ldg(s1 + d), s);
                dl[i] = s;
            }
            for (int i = 0; i < prm.KR; ++i) {   // (Q + Q^T) d
                float s = 0.f;
                for (int j = 0; j < prm.KR; ++j)
                    s = fmaf(__ldg(prm.rew_Q + i * prm.KR + j) + __ldg(prm.rew_Q + j * prm.KR + i), dl[j], s);
                qd[i] = s;
            }
            for (int d = 0; d < D; ++d) {
                float s = 0.f;
                for (int i = 0; i < prm.KR; ++i) s = fmaf(qd[i], __ldg(prm.rew_C + i * D + d), s);
                gs[p * SD + d] += w * s;
            }
            for (int u = 0; u < U; ++u) {
                float s = 0.f;
                for (int v = 0; v < U; ++v)
                    s = fmaf(__ldg(prm.rew_R + u * U + v) + __ldg(prm.rew_R + v * U + u), __ldg(a + v), s);
                float g0 = prm.g_actions ? __ldg(prm.g_actions + ((size_t)t * N + n) * U + u) : 0.f;
                ga[p * SD + u] = g0 + w * s;
            }
        }
        __syncthreads();
        // ---- dynamics density adjoint: s' = s + mu*Sy + my + z*exp(lstd) ----
        for (int i = tid; i < P * D; i += NT) {
            int p = i / D, d = i - p * D;
            int n = min(n0 + p, N - 1);
            float g = gs[p * SD + d];
            float sy = __ldg(prm.Sy + d);
            gsp[p * SD + d] = g;
            in[d * P + p] = g * sy;
            if (dyn.has_density) {
                float ls = __ldg(prm.ws + dyn.outsaved_off + ((size_t)t * N + n) * dyn.nout + D + d);
                float lst = clamp_logstd(ls, dyn.lmax) + logf(sy);
                float z = __ldg(dyn.z + (size_t)t * dyn.zstride + (size_t)n * D + d);
                in[(D + d) * P + p] = g * z * expf(lst) * sigmoid_f(dyn.lmax - ls);
            }
        }
        net_backward<P, false>(prm, dyn, sched_i, in, out, obuf, res, red, S, t, n0);
        __syncthreads();
        // ---- through the input scaler: d[s;a] = dx * iSx ----
        for (int i = tid; i < P * (D + U); i += NT) {
            int p = i / (D + U), k = i - p * (D + U);
            float v = obuf[p * dyn.nin + k] * __ldg(prm.iSx + k);
            if (k < D) gsp[p * SD + k] += v;
            else ga[p * SD + (k - D)] += v;
        }
        __syncthreads();
        // ---- tanh squash + policy density adjoint: a = scale*tanh(u)+bias, u = mu + z*exp(lstd) ----
        for (int i = tid; i < P * U; i += NT) {
            int p = i / U, u = i - p * U;
            int n = min(n0 + p, N - 1);
            const float *op = prm.ws + pol.outsaved_off + ((size_t)t * N + n) * pol.nout;
            float sc = __ldg(prm.act_scale + u);
            float g = ga[p * SD + u];
            float du, dls = 0.f;
            if (pol.has_density) {
                float mu = __ldg(op + u), ls = __ldg(op + U + u);
                float lst = clamp_logstd(ls, pol.lmax);
                float z = __ldg(pol.z + (size_t)t * pol.zstride + (size_t)n * U + u);
                float el = expf(lst);
                float th = tanhf(mu + z * el);
                du = g * sc * (1.f - th * th);
                dls = du * z * el * sigmoid_f(pol.lmax - ls);
                out[(U + u) * P + p] = dls;
            } else {
                float th = tanhf(__ldg(op + u));
                du = g * sc * (1.f - th * th);
            }
            out[u * P + p] = du;
            if (n0 + p < N) {
                float *dd = prm.ws + pol.delta_off[pol.nlin - 1] + ((size_t)t * N + n0 + p) * pol.nout;
                dd[u] = du;
                if (pol.has_density) dd[U + u] = dls;
            }
        }
        {
            float *tmp = in; in = out; out = tmp;
        }
        net_backward<P, true>(prm, pol, sched_i, in, out, obuf, res, red, S, t, n0);
        __syncthreads();
        // ---- dL/ds_t = carried + through dynamics input + through policy input + direct cotangent ----
        for (int i = tid; i < P * D; i += NT) {
            int p = i / D, d = i - p * D;
            int n = min(n0 + p, N - 1);
            float g = gsp[p * SD + d] + obuf[p * pol.nin + d];
            if (prm.g_states) g += __ldg(prm.g_states + ((size_t)t * N + n) * D + d);
            gs[p * SD + d] = g;
        }
        __syncthreads();
    }
    if (prm.dx0) {
        for (int i = tid; i < P * D; i += NT) {
            int p = i / D, d = i - p * D;
            if (n0 + p < N) prm.dx0[(size_t)(n0 + p) * D + d] = gs[p * SD + d];
        }
    }
}

cudaError_t launch_rollout_bwd(const SweepParams &prm, int P, int smem_bytes, cudaStream_t stream) {
    const int grid = (prm.N + P - 1) / P;
    cudaError_t e;
#define PMB_LAUNCH_BWD(PP)                                                                                   \
    case PP:                                                                                                 \
        e = cudaFuncSetAttribute(rollout_bwd_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                 smem_bytes);                                                                \
        if (e != cudaSuccess) return e;                                                                      \
        rollout_bwd_kernel<PP><<<grid, NT, smem_bytes, stream>>>(prm);                                       \
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
