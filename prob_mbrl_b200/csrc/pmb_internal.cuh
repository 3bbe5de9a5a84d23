// Internal plan structures and device building blocks of the fused rollout sweeps (sm_100a).
//
// One CTA owns P particles for the whole horizon; its (P x width) activation tile never leaves
// shared memory between layers or steps.  Per step the CTA walks the policy MLP, the action
// squashing, the dynamics MLP, the Gaussian output densities and the reward (forward sweep,
// pmb_rollout_fwd.cu) or their adjoints in reverse (pmb_rollout_bwd.cu).  The hidden x hidden
// weight matrices do not fit next to the tile, so they are streamed from L2 every step as
// k-chunks through a ring of shared-memory stages filled by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx); the skinny first/last layers stay resident.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pmb_b200.h"

namespace pmb {

constexpr int NT = 256;              // threads per CTA of the sweeps
constexpr int NWARP = NT / 32;
constexpr int MAXL = PMB_MAX_LINEAR; // linear layers per net
constexpr int MAXS = 8;              // ring stages (max)
constexpr int MAXKS = 8;             // max k-split of a wide layer
constexpr int MAXSCHED = 2 * MAXL;   // streamed layers per step

// One linear op as executed inside a sweep.
//   wide  : out[p][0..Npad) = sum_{k<K} in[k][p] * Wm[k][0..Npad)   (Wm rows of Npad floats)
//   narrow: out[p][j<Nout]  = sum_{k<K} in[k][p] * Wm[j][k]         (Wm rows of K floats)
struct Lin {
    int kind;        // 0 = wide, 1 = narrow
    int K;           // reduction length as stored
    int Nout;        // true outputs
    int Npad;        // wide: padded row length (multiple of 4); narrow: Nout
    int streamed;    // wide only: 1 = through the ring, 0 = resident in smem
    int kc;          // rows per chunk (streamed)
    int nchunks;     // chunks per layer (streamed)
    int soff;        // resident: float offset inside the resident smem area
    long long goff;  // float offset of the matrix inside the packed weight area of this sweep
    long long boff;  // float offset (workspace) of the padded bias, -1 = none (forward only)
};

struct NetSweep {
    int nlin;                   // L hidden + 1
    int nout;                   // outputs of the last linear layer
    int nin;                    // inputs of the first
    Lin lin[MAXL];              // indexed by linear layer l = 0..nlin-1
    long long mask_off[MAXL];   // hidden layer l: packed mask [N][Npad_l] (workspace floats), -1 = none
    long long saved_off[MAXL];  // hidden layer l: post-dropout activations [H][N][Npad_l]
    long long delta_off[MAXL];  // policy only: adjoint of linear l's output [H][N][Npad_l or nout]
    long long outsaved_off;     // raw output of the last linear layer [H][N][nout]
    float keep[MAXL];
    int has_density;
    float lmax;
    const float *z;
    long long zstride;
};

struct StreamItem {
    long long goff;  // float offset in packed area
    int kc, nchunks, K, Npad;
};

struct SweepParams {
    int N, H, D, U;
    int stream_mode;            // 1 = synchronous copies, 2 = TMA bulk + mbarrier
    NetSweep pol, dyn;
    const float *wpack;         // packed weights of THIS sweep (fwd or bwd area)
    float *ws;                  // workspace base (floats)
    // scalers / squashing / reward
    const float *act_scale, *act_bias, *mx, *iSx, *my, *Sy;
    int KR;
    const float *rew_C, *rew_c0, *rew_Q, *rew_R;
    float rew_scale, rew_offset;
    // trajectories
    const float *x0;
    float *states, *actions, *rewards;
    // cotangents (backward)
    const float *g_states, *g_actions, *g_rewards;
    float *dx0;
    int *status;
    // streaming schedule for one step, in consumption order
    int nsched;
    int chunks_per_step;
    StreamItem sched[MAXSCHED];
    // shared memory carve-up (float offsets from the dynamic smem base)
    int res_floats;             // resident weights
    long long res_goff_unused;
    int off_act0, off_act1, off_red, off_misc, off_stage;
    int stage_floats, nstages;
    int hmax_pad;
    // resident copy list: (goff -> soff, n floats)
    int nres;
    long long res_goff[2 * MAXL];
    int res_soff[2 * MAXL];
    int res_n[2 * MAXL];
};

// ----------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared), see the Blackwell guide "Guideline 15".
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a byte-count mismatch must surface as a trapped kernel, not a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// smooth upper clamp of log-std: lmax - softplus(lmax - l)  (reference models/densities.py:97-98;
// softplus as torch: x > 20 ? x : log1p(exp(x)))
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float clamp_logstd(float l, float lmax) { return lmax - softplus_f(lmax - l); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// ----------------------------------------------------------------------------------------
// The weight stream: a ring of `nstages` smem stages consumed in a fixed cyclic schedule.
// All threads call consume() in lockstep; thread 0 issues the copies `nstages-1` chunks ahead.
// ----------------------------------------------------------------------------------------
struct Stream {
    const SweepParams *prm;
    float *stage_base;
    uint64_t *full;         // [nstages] mbarriers (mode 2)
    unsigned q;             // next chunk to consume (uniform)
    // issue cursor (thread 0 only)
    unsigned issued;
    unsigned total;
    int is_item, is_chunk;

    __device__ __forceinline__ void init(const SweepParams *p, float *smem, uint64_t *bars) {
        prm = p;
        stage_base = smem + p->off_stage;
        full = bars;
        q = 0;
        issued = 0;
        total = (unsigned)p->H * (unsigned)p->chunks_per_step;
        is_item = 0;
        is_chunk = 0;
        if (p->stream_mode == 2) {
            if (threadIdx.x == 0) {
                for (int s = 0; s < p->nstages; ++s) mbar_init(&full[s], 1);
                fence_mbar_init();
                fence_proxy_async();
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int s = 0; s + 1 < p->nstages; ++s) issue_next();
            }
        }
    }

    // thread 0: issue the next chunk of the cyclic schedule into its stage
    __device__ __forceinline__ void issue_next() {
        if (issued >= total) return;
        const StreamItem &it = prm->sched[is_item];
        int row0 = is_chunk * it.kc;
        int rows = min(it.kc, it.K - row0);
        uint32_t bytes = (uint32_t)rows * (uint32_t)it.Npad * 4u;
        int st = issued % prm->nstages;
        const float *src = prm->wpack + it.goff + (long long)row0 * it.Npad;
        mbar_expect_tx(&full[st], bytes);
        tma_bulk_g2s(stage_base + (size_t)st * prm->stage_floats, src, bytes, &full[st]);
        ++issued;
        if (++is_chunk == it.nchunks) {
            is_chunk = 0;
            if (++is_item == prm->nsched) is_item = 0;
        }
    }

    // Make chunk q of the schedule readable in its stage; returns the stage pointer.
    // Contains one __syncthreads() before any stage is overwritten, so callers may rely on it as the
    // barrier that publishes the previous layer's shared-memory writes.
    __device__ __forceinline__ const float *acquire(const StreamItem &it, int chunk) {
        const int ns = prm->nstages;
        __syncthreads();   // everyone is done with chunk q-1 -> its stage may be refilled
        const float *st = stage_base + (size_t)(q % ns) * prm->stage_floats;
        if (prm->stream_mode == 2) {
            if (threadIdx.x == 0) issue_next();
            mbar_wait(&full[q % ns], (q / ns) & 1u);
        } else {
            int row0 = chunk * it.kc;
            int rows = min(it.kc, it.K - row0);
            int n4 = rows * it.Npad / 4;
            const float4 *src = reinterpret_cast<const float4 *>(prm->wpack + it.goff + (long long)row0 * it.Npad);
            float4 *dst = reinterpret_cast<float4 *>(const_cast<float *>(st));
            for (int i = threadIdx.x; i < n4; i += NT) dst[i] = __ldg(src + i);
            __syncthreads();
        }
        ++q;
        return st;
    }
};

// ----------------------------------------------------------------------------------------
// wide layer: thread = (column quad cq, k-split group g)
// ----------------------------------------------------------------------------------------
struct WideMap {
    int cq, g, ks, active;
    __device__ __forceinline__ void set(int npad) {
        int cqn = npad >> 2;
        int gs = (cqn + 31) & ~31;          // threads per k-split group (warp multiple)
        ks = NT / gs;
        if (ks > MAXKS) ks = MAXKS;
        if (ks < 1) ks = 1;
        g = threadIdx.x / gs;
        cq = threadIdx.x - g * gs;
        active = (g < ks) && (cq < cqn);
    }
};

template <int P>
__device__ __forceinline__ void load_act(float (&a)[P], const float *src) {
    if constexpr (P % 4 == 0) {
#pragma unroll
        for (int i = 0; i < P / 4; ++i) {
            float4 v = *reinterpret_cast<const float4 *>(src + 4 * i);
            a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
        }
    } else if constexpr (P == 2) {
        float2 v = *reinterpret_cast<const float2 *>(src);
        a[0] = v.x; a[1] = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) a[i] = src[i];
    }
}

// acc[p][0..3] += sum over rows r = g, g+ks, ... < rows of act[r][p] * w[r][4cq..4cq+3]
template <int P>
__device__ __forceinline__ void wide_accum(float (&acc)[P][4], const float *__restrict__ w, int rows, int npad,
                                           const float *__restrict__ act, const WideMap &m) {
    const float *wp = w + 4 * m.cq;
#pragma unroll 4
    for (int r = m.g; r < rows; r += m.ks) {
        float4 wv = *reinterpret_cast<const float4 *>(wp + (size_t)r * npad);
        float a[P];
        load_act<P>(a, act + r * P);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            acc[p][0] = fmaf(a[p], wv.x, acc[p][0]);
            acc[p][1] = fmaf(a[p], wv.y, acc[p][1]);
            acc[p][2] = fmaf(a[p], wv.z, acc[p][2]);
            acc[p][3] = fmaf(a[p], wv.w, acc[p][3]);
        }
    }
}

// Sum the k-split partials into group 0.  Contains one __syncthreads().
template <int P>
__device__ __forceinline__ void wide_reduce(float (&acc)[P][4], float *red, int npad, const WideMap &m) {
    if (m.ks > 1) {
        if (m.active && m.g > 0) {
#pragma unroll
            for (int p = 0; p < P; ++p)
                *reinterpret_cast<float4 *>(red + ((size_t)((m.g - 1) * P + p) * npad) + 4 * m.cq) =
                    make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        }
        __syncthreads();
        if (m.active && m.g == 0) {
            for (int gg = 1; gg < m.ks; ++gg) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float4 v = *reinterpret_cast<const float4 *>(red + ((size_t)((gg - 1) * P + p) * npad) + 4 * m.cq);
                    acc[p][0] += v.x; acc[p][1] += v.y; acc[p][2] += v.z; acc[p][3] += v.w;
                }
            }
        }
    }
}

// Accumulate a whole wide layer (resident or streamed) into acc.  `act` is the [K][P] input tile.
// On return group-0 threads hold the full sums.  Always starts with a __syncthreads().
template <int P>
__device__ __forceinline__ void wide_layer(float (&acc)[P][4], const Lin &L, const StreamItem *item,
                                           const float *res, const float *act, float *red, Stream &S,
                                           const WideMap &m) {
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.f;
    if (L.streamed) {
        for (int c = 0; c < L.nchunks; ++c) {
            const float *w = S.acquire(*item, c);
            int row0 = c * L.kc;
            int rows = min(L.kc, L.K - row0);
            if (m.active) wide_accum<P>(acc, w, rows, L.Npad, act + (size_t)row0 * P, m);
        }
    } else {
        __syncthreads();
        if (m.active) wide_accum<P>(acc, res + L.soff, L.K, L.Npad, act, m);
    }
    wide_reduce<P>(acc, red, L.Npad, m);
}

// narrow layer: out[p][j] = sum_k act[k][p] * w[j][k] (+ bias[j]); one warp per output, lanes split k.
// Starts with a __syncthreads(); results are visible after the caller's next barrier.
template <int P>
__device__ __forceinline__ void narrow_layer(const Lin &L, const float *res, const float *act, float *out,
                                             const float *bias) {
    __syncthreads();
    const float *w = res + L.soff;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nout = P * L.Nout;
    for (int o = warp; o < nout; o += NWARP) {
        int p = o / L.Nout, j = o - p * L.Nout;
        const float *wj = w + (size_t)j * L.K;
        float s = 0.f;
        for (int k = lane; k < L.K; k += 32) s = fmaf(act[k * P + p], wj[k], s);
        s = warp_sum(s);
        if (lane == 0) out[p * L.Nout + j] = s + (bias ? bias[j] : 0.f);
    }
}

// cooperative copy of the resident weights into shared memory
__device__ __forceinline__ void load_resident(const SweepParams &prm, float *res) {
    for (int i = 0; i < prm.nres; ++i) {
        const float4 *src = reinterpret_cast<const float4 *>(prm.wpack + prm.res_goff[i]);
        float4 *dst = reinterpret_cast<float4 *>(res + prm.res_soff[i]);
        for (int k = threadIdx.x; k < prm.res_n[i] / 4; k += NT) dst[k] = __ldg(src + k);
    }
}

}  // namespace pmb
