// Internal plan structures and device building blocks of the fused rollout sweeps (sm_100a).
//
// One CTA owns P particles for the whole horizon; its (P x width) activation tile never leaves
// shared memory between layers or steps.  Per step the CTA walks the policy MLP, the action
// squashing, the dynamics MLP, the Gaussian output densities and the reward (forward sweep,
// pmb_rollout_fwd.cu) or their adjoints in reverse (pmb_rollout_bwd.cu).  The hidden x hidden
// weight matrices do not fit next to the tile, so they are streamed from L2 every step as large
// k-chunks through a ring of shared-memory stages filled by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx); the skinny first/last layers, the biases, the CTA's rows of
// the dropout masks and every per-step constant stay resident in shared memory.
// Inner products run on the packed FP32 pipe (fma.rn.f32x2, SASS FFMA2 -- new on sm_100).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pmb_b200.h"

namespace pmb {

constexpr int NT = 256;              // COMPUTE threads per CTA of the sweeps (8 warps)
constexpr int NWARP = NT / 32;
constexpr int NT_LAUNCH = NT + 32;   // + one producer warp that only issues TMA copies
// barrier among the compute warps only (the producer warp never joins it)
#define CTA_SYNC() asm volatile("bar.sync 1, 256;" ::: "memory")
constexpr int MAXL = PMB_MAX_LINEAR; // linear layers per net
constexpr int MAXS = 4;              // ring stages (max)
constexpr int MAXSCHED = 2 * MAXL;   // streamed layers per step
constexpr int SD = PMB_MAX_STATE;    // row stride of the small per-particle buffers (D + U <= 16)

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
    int bias_soff;   // float offset of the padded bias in the resident area, -1 = none (forward only)
    long long goff;  // float offset of the matrix inside the packed weight area of this sweep
    long long boff;  // float offset (workspace) of the padded bias, -1 = none
};

struct NetSweep {
    int nlin;                   // L hidden + 1
    int nout;                   // outputs of the last linear layer
    int nin;                    // inputs of the first
    Lin lin[MAXL];              // indexed by linear layer l = 0..nlin-1
    long long mask_off[MAXL];   // hidden layer l: packed mask [N][Npad_l] (workspace floats), -1 = none
    int mask_soff[MAXL];        // hidden layer l: this CTA's [P][Npad_l] rows in smem, -1 = read from global
    long long saved_off[MAXL];  // hidden layer l: post-dropout activations [H][N][Npad_l]
    int sav_soff[MAXL];         // backward: offset of the layer inside the per-step saved tile in smem
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

// layout of the constants block in shared memory (floats)
// (reward matrices C [KR][D], Q [KR][KR], R [U][U] and the symmetrised Q + Q^T, R + R^T: rows of SD floats)
constexpr int C_MX = 0, C_ISX = 16, C_MY = 32, C_SY = 48, C_LSY = 64, C_SCALE = 80, C_BIAS = 96, C_C0 = 112,
              C_C = 128, C_Q = 384, C_QS = 640, C_R = 896, C_RS = 1152, C_TOTAL = 1408;
static_assert(PMB_MAX_REWARD_ROWS <= SD, "reward rows are stored with the per-particle row stride");

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
    float *da_total;            // backward: total dL/da_t [H][N][U] (nullable)
    int *status;
    // moment matching (pmb_mm.cuh)
    int mm_states, mm_rewards, mm_G, mm_Ng;
    const float *z_mm;
    float *s1pre;               // [H][N][D] particles before moment matching
    float *mmstat;              // [H][G][3*SD + SD*SD] mean, z mean, 1/z std, Cholesky factor of every step
    double *mmrec;              // [2][grid][MMREC] per-CTA block statistics
    unsigned *mmctr;            // grid-barrier arrival counter (zeroed before launch)
    int off_mm;                 // shared-memory scratch of the mm step (floats)
    long long *dbg;             // profiling aid: clock64() marks of CTA 0 / thread 0 at step H/2 (nullable)
    // streaming schedule for one step, in consumption order
    int nsched;
    int chunks_per_step;
    StreamItem sched[MAXSCHED];
    // shared memory carve-up (float offsets from the dynamic smem base)
    int off_cst, off_act0, off_act1, off_red, off_misc, off_sav, off_stage;
    int sav_floats;             // backward: floats of one saved-activation tile buffer (two buffers)
    int stage_floats, nstages;
    // resident copy list: (source -> smem offset, n floats); src_ws = 1: workspace offset, 0: wpack offset
    int nres;
    long long res_goff[4 * MAXL];
    int res_soff[4 * MAXL];
    int res_n[4 * MAXL];
    int res_ws[4 * MAXL];
};

// ----------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (global -> shared)
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// timeline marks (profiling aid; compiled in, a predicated store when enabled)
#define PMB_MARK(i) do { if (dbg_on) prm.dbg[(i)] = clock64(); } while (0)

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
// exp(clamp_logstd(l, lmax)) = exp(lmax - log(1 + exp(lmax - l))) = exp(lmax) / (1 + exp(lmax - l)):
// one expf and one division on the serial chain instead of expf + log1pf + expf (same value to ~2 ulp;
// for lmax - l > 88 it underflows to 0 exactly like exp(l) does).  `elmax` = exp(lmax), precomputed.
__device__ __forceinline__ float exp_clamped_logstd(float l, float lmax, float elmax) {
    return elmax / (1.f + expf(lmax - l));
}

// ----------------------------------------------------------------------------------------
// The weight stream: a ring of `nstages` smem stages consumed in a fixed cyclic schedule.
// The 8 compute warps consume it in lockstep; a dedicated producer warp re-fills released stages.
// ----------------------------------------------------------------------------------------
struct ChunkDesc {
    const float *src;
    uint32_t bytes;
    uint32_t pad;
};
constexpr int MAXCHUNKS = 64;   // chunks per step (planner guarantees chunks_per_step <= MAXCHUNKS)

// Barriers of the weight ring (static shared memory of the kernels).
struct RingBars {
    uint64_t full[MAXS];    // producer -> consumers: chunk landed (TMA complete_tx)
    uint64_t empty[MAXS];   // consumers -> producer: all 8 compute warps are done with the stage
};

// Consumer view of the weight ring (compute warps, in lockstep).
struct Stream {
    const SweepParams *prm;
    float *stage_base;
    RingBars *bars;
    int c_stage;
    uint32_t c_parity;

    __device__ __forceinline__ void init(const SweepParams *p, float *smem, RingBars *b) {
        prm = p;
        stage_base = smem + p->off_stage;
        bars = b;
        c_stage = 0;
        c_parity = 0;
    }
    // wait for the next chunk of the schedule; returns its stage pointer
    __device__ __forceinline__ const float *acquire() {
        mbar_wait(&bars->full[c_stage], c_parity);
        return stage_base + (size_t)c_stage * prm->stage_floats;
    }
    // every lane of the warp is done reading the current stage: hand it back to the producer
    __device__ __forceinline__ void release() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&bars->empty[c_stage]);
        if (++c_stage == prm->nstages) {
            c_stage = 0;
            c_parity ^= 1u;
        }
    }
};

// Producer side: one elected thread of the extra warp walks the cyclic chunk schedule of all H steps,
// re-filling a stage as soon as the 8 compute warps released it.
__device__ __forceinline__ void ring_fill_table(const SweepParams &p, ChunkDesc *tab) {
    int k = 0;
    for (int i = 0; i < p.nsched; ++i) {
        const StreamItem &it = p.sched[i];
        for (int c = 0; c < it.nchunks; ++c, ++k) {
            const int row0 = c * it.kc;
            const int rows = min(it.kc, it.K - row0);
            tab[k].src = p.wpack + it.goff + (long long)row0 * it.Npad;
            tab[k].bytes = (uint32_t)rows * (uint32_t)it.Npad * 4u;
        }
    }
}

struct RingProducer {
    int stage, idx;
    uint32_t parity;      // parity of the NEXT wait on empty[stage]
    unsigned issued;
    __device__ __forceinline__ void init() { stage = idx = 0; parity = 0; issued = 0; }
    // issue `n` chunks (blocking on stage availability)
    __device__ __forceinline__ void issue(const SweepParams &p, float *stage_base, RingBars *bars,
                                          const ChunkDesc *tab, int n) {
        for (int i = 0; i < n; ++i) {
            if (issued >= (unsigned)p.nstages) {
                mbar_wait(&bars->empty[stage], parity);   // all 8 compute warps released the stage ...
                fence_proxy_async();                      // ... order their generic reads before the TMA write
            }
            const ChunkDesc d = tab[idx];
            mbar_expect_tx(&bars->full[stage], d.bytes);
            tma_bulk_g2s(stage_base + (size_t)stage * p.stage_floats, d.src, d.bytes, &bars->full[stage]);
            ++issued;
            if (++idx == p.chunks_per_step) idx = 0;
            if (++stage == p.nstages) {
                stage = 0;
                if (issued > (unsigned)p.nstages) parity ^= 1u;
            }
        }
    }
};

// ----------------------------------------------------------------------------------------
// wide layer: thread = (column quad cq, k-split group g); ks is a power of two
// ----------------------------------------------------------------------------------------
struct WideMap {
    int cq, g, ks, ks_log2, active;
    __device__ __forceinline__ void set(int npad) {
        const int cqn = npad >> 2;
        // threads per k-split group: power of two >= number of column quads, at least one warp
        const int gs_log2 = cqn <= 32 ? 5 : cqn <= 64 ? 6 : cqn <= 128 ? 7 : 8;
        ks_log2 = 8 - gs_log2;               // NT = 256 threads
        ks = 1 << ks_log2;
        g = threadIdx.x >> gs_log2;
        cq = threadIdx.x & ((1 << gs_log2) - 1);
        active = cq < cqn;
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

// acc[p][0..1] (column pairs) += act[r][p] * w[r][4cq..4cq+3] over this thread's rows r = g, g+ks, ...
// Packed FP32 FMA (FFMA2): two columns per instruction.
template <int P>
__device__ __forceinline__ void wide_accum(float2 (&acc)[P][2], const float *__restrict__ w, int rows, int npad,
                                           const float *__restrict__ act, const WideMap &m) {
    const int n = (rows - m.g + m.ks - 1) >> m.ks_log2;
    const float *wp = w + 4 * m.cq + (size_t)m.g * npad;
    const float *ap = act + m.g * P;
    const int wstride = npad << m.ks_log2;
    const int astride = P << m.ks_log2;
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        const float4 wv = *reinterpret_cast<const float4 *>(wp);
        float a[P];
        load_act<P>(a, ap);
        wp += wstride;
        ap += astride;
        const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float2 a2 = make_float2(a[p], a[p]);
            acc[p][0] = __ffma2_rn(a2, w01, acc[p][0]);
            acc[p][1] = __ffma2_rn(a2, w23, acc[p][1]);
        }
    }
}

// Accumulate a whole wide layer (resident or streamed), combine the k-split partials and hand every
// finished (particle p, 4 columns) tuple to `epi(p, float4 sums, active)` (called by every lane of the
// epilogue warps; `active` = the lane owns real columns).  The k-split groups share the
// epilogue work: group g finishes particles p = g, g+ks, ...  Starts with a CTA_SYNC(); contains a
// second one when ks > 1.  `act` is the [K][P] input tile.
template <int P, typename Epi>
__device__ __forceinline__ void wide_layer(const Lin &L, const StreamItem *item, const float *res,
                                           const float *act, float *red, Stream &S, const WideMap &m,
                                           Epi epi, long long *dbg = nullptr) {
    float2 acc[P][2];
    int dbi = 0;
#define PMB_WMARK() do { if (dbg) dbg[dbi++] = clock64(); } while (0)
    PMB_WMARK();
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p][0] = acc[p][1] = make_float2(0.f, 0.f);
    if (L.streamed) {
        CTA_SYNC();        // publish the input tile written by the previous phase
#pragma unroll 1
        for (int c = 0; c < L.nchunks; ++c) {
            const float *w = S.acquire();
            PMB_WMARK();
            const int row0 = c * L.kc;
            const int rows = min(L.kc, L.K - row0);
            if (m.active) wide_accum<P>(acc, w, rows, L.Npad, act + (size_t)row0 * P, m);
            S.release();
            PMB_WMARK();
        }
    } else {
        CTA_SYNC();
        PMB_WMARK();
        if (m.active) wide_accum<P>(acc, res + L.soff, L.K, L.Npad, act, m);
        PMB_WMARK();
    }
    if (m.ks == 1) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            epi(p, make_float4(acc[p][0].x, acc[p][0].y, acc[p][1].x, acc[p][1].y), m.active != 0);
        return;
    }
    const int npad = L.Npad;
    if (m.active) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            *reinterpret_cast<float4 *>(red + ((size_t)(m.g * P + p) * npad) + 4 * m.cq) =
                make_float4(acc[p][0].x, acc[p][0].y, acc[p][1].x, acc[p][1].y);
    }
    PMB_WMARK();
    CTA_SYNC();
    PMB_WMARK();
    // whole warps walk this loop (g is warp-uniform), so the epilogue may use warp collectives
#pragma unroll 1
    for (int p = m.g; p < P; p += m.ks) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m.active) {
            s = *reinterpret_cast<const float4 *>(red + ((size_t)p * npad) + 4 * m.cq);
            for (int gg = 1; gg < m.ks; ++gg) {
                const float4 v = *reinterpret_cast<const float4 *>(red + ((size_t)(gg * P + p) * npad) + 4 * m.cq);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        epi(p, s, m.active != 0);
    }
    PMB_WMARK();
#undef PMB_WMARK
}

// thin-K wide layer (first layers, K = D / D+U; output-layer adjoints, K = 2D / 2U): one thread per
// output column, all P particles, no k-split and no reduction.  epi(j, acc[P]) finishes column j.
// Starts with a CTA_SYNC().
template <int P, typename Epi>
__device__ __forceinline__ void thin_layer(const Lin &L, const float *res, const float *act, Epi epi) {
    CTA_SYNC();
    const float *w = res + L.soff;
    const int K = L.K, npad = L.Npad;
#pragma unroll 1
    for (int j = threadIdx.x; j < L.Nout; j += NT) {
        float acc[P];
#pragma unroll
        for (int p = 0; p < P; ++p) acc[p] = 0.f;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float wv = w[k * npad + j];
            float a[P];
            load_act<P>(a, act + k * P);
#pragma unroll
            for (int p = 0; p < P; ++p) acc[p] = fmaf(a[p], wv, acc[p]);
        }
        epi(j, acc);
    }
}

// narrow layer: out[p][j] = sum_k act[k][p] * w[j][k] (+ bias[j]) for P*Nout <= 256 outputs.
// Outputs across lanes, the reduction axis k split across the warps that share an output: with
// R = 256 / pow2ceil(P*Nout) k-slices per output, thread = (slice r, output o); partials meet in `red`.
// The thread mapping is fixed for the whole horizon: NarrowMap is computed once per kernel.
struct NarrowMap {
    int active;      // this thread owns an (output, k-slice)
    int final;       // this thread sums the slices of output o
    int o, j, ol2, R;
    int k0, k1;
    int woff;        // j * K
    int aoff;        // particle index p
    template <int P>
    __device__ __forceinline__ void set(const Lin &L) {
        const int K = L.K, nout = L.Nout;
        const int no = P * nout;
        ol2 = no <= 32 ? 5 : no <= 64 ? 6 : no <= 128 ? 7 : 8;
        R = NT >> ol2;
        o = threadIdx.x & ((1 << ol2) - 1);
        const int r = threadIdx.x >> ol2;
        const int kper = (K + R - 1) / R;
        k0 = min(K, r * kper);
        k1 = min(K, k0 + kper);
        active = o < no;
        final = active && r == 0;
        const int p = active ? o / nout : 0;
        j = active ? o - p * nout : 0;
        woff = j * K;
        aoff = p;
    }
};

// Starts with a CTA_SYNC() and contains a second one; results are visible after the caller's next
// barrier.
template <int P>
__device__ __forceinline__ void narrow_layer(const Lin &L, const NarrowMap &nm, const float *res, const float *act,
                                             float *out, const float *bias, float *red) {
    CTA_SYNC();
    if (nm.active) {
        const float *wj = res + L.soff + nm.woff;
        const float *ap = act + nm.aoff;
        float s = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int k = nm.k0;
        for (; k + 4 <= nm.k1; k += 4) {
            s = fmaf(ap[k * P], wj[k], s);
            s1 = fmaf(ap[(k + 1) * P], wj[k + 1], s1);
            s2 = fmaf(ap[(k + 2) * P], wj[k + 2], s2);
            s3 = fmaf(ap[(k + 3) * P], wj[k + 3], s3);
        }
        for (; k < nm.k1; ++k) s = fmaf(ap[k * P], wj[k], s);
        red[((threadIdx.x >> nm.ol2) << nm.ol2) + nm.o] = (s + s1) + (s2 + s3);
    }
    CTA_SYNC();
    if (nm.final) {
        float v = red[nm.o];
        for (int rr = 1; rr < nm.R; ++rr) v += red[(rr << nm.ol2) + nm.o];
        out[nm.o] = v + (bias ? bias[nm.j] : 0.f);
    }
}

// Output projection fused into the epilogue of the last hidden layer: every lane holds 4 finished hidden
// values v of particle p (zeros on idle lanes); the warp forms its share of out[p][j] = sum_k h[k] w[j][k]
// for all j with 4 dot products in flight and a butterfly reduction, lane 0 publishes one partial per
// (p, j, warp-of-the-group).  The consumers add the partials of the group's warps (read_out).
template <int B>
__device__ __forceinline__ void narrow_fused_batch(const float4 v, bool active, int p, int col, const float *wn, int K,
                                                   int nout, int j0, float *part, int wpg, int wig) {
    const int lane = threadIdx.x & 31;
    float s[B];
#pragma unroll
    for (int q = 0; q < B; ++q) {
        const int j = min(j0 + q, nout - 1);
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) w = *reinterpret_cast<const float4 *>(wn + (size_t)j * K + col);
        s[q] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, v.w * w.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < B; ++q) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
    }
    if (lane < B && j0 + lane < nout) {
        float r = s[0];
#pragma unroll
        for (int q = 1; q < B; ++q) r = lane == q ? s[q] : r;
        part[(p * nout + j0 + lane) * wpg + wig] = r;
    }
}
// B dot products in flight: 4 for the 2-output policy head, 12 for the 10/16-output heads (B = 12 covers
// the 10 outputs of a 5-dim state in one batch).
__device__ __forceinline__ void narrow_fused(const float4 v, bool active, int p, int col, const float *wn, int K,
                                             int nout, float *part, int wpg, int wig) {
    if (nout <= 4) {
        narrow_fused_batch<4>(v, active, p, col, wn, K, nout, 0, part, wpg, wig);
    } else if (nout <= 8) {
        narrow_fused_batch<8>(v, active, p, col, wn, K, nout, 0, part, wpg, wig);
    } else {
#pragma unroll 1
        for (int j0 = 0; j0 < nout; j0 += 12) narrow_fused_batch<12>(v, active, p, col, wn, K, nout, j0, part, wpg, wig);
    }
}
// out[p][j] from the published partials (wpg = 1, bias = nullptr: plain buffer written by narrow_layer)
__device__ __forceinline__ float read_out(const float *part, const float *bias, int p, int j, int nout, int wpg) {
    const float *q = part + (p * nout + j) * wpg;
    float v = bias ? bias[j] : 0.f;
    for (int w = 0; w < wpg; ++w) v += q[w];
    return v;
}

// cooperative copy of the resident blocks (weights, biases, this CTA's mask rows) into shared memory.
// res_ws: 0 = offset into the packed weights, 1 = offset into the workspace, >= 4 = dropout-mask rows of
// this CTA (value = row length Npad; res_n = P * Npad; row p is particle min(n0 + p, N - 1)).
__device__ __forceinline__ void load_resident(const SweepParams &prm, float *smem, int n0) {
    for (int i = 0; i < prm.nres; ++i) {
        float4 *dst = reinterpret_cast<float4 *>(smem + prm.res_soff[i]);
        if (prm.res_ws[i] >= 4) {
            const int row4 = prm.res_ws[i] / 4;
            const int n4 = prm.res_n[i] / 4;
            for (int k = threadIdx.x; k < n4; k += NT) {
                const int p = k / row4, c = k - p * row4;
                const int n = min(n0 + p, prm.N - 1);
                dst[k] = __ldg(reinterpret_cast<const float4 *>(prm.ws + prm.res_goff[i] + (long long)n * row4 * 4) + c);
            }
        } else {
            const float *src = (prm.res_ws[i] == 0 ? prm.wpack : prm.ws) + prm.res_goff[i];
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
            for (int k = threadIdx.x; k < prm.res_n[i] / 4; k += NT) dst[k] = __ldg(s4 + k);
        }
    }
}

// constants block: scalers, squashing, reward matrices (+ symmetrised copies for the adjoint)
template <typename PRM>
__device__ __forceinline__ void load_constants(const PRM &prm, float *cst) {
    const int D = prm.D, U = prm.U, KR = prm.KR;
    for (int i = threadIdx.x; i < C_TOTAL; i += NT) cst[i] = 0.f;
    CTA_SYNC();
    for (int i = threadIdx.x; i < D + U; i += NT) {
        cst[C_MX + i] = prm.mx[i];
        cst[C_ISX + i] = prm.iSx[i];
    }
    for (int i = threadIdx.x; i < D; i += NT) {
        cst[C_MY + i] = prm.my[i];
        cst[C_SY + i] = prm.Sy[i];
        cst[C_LSY + i] = logf(prm.Sy[i]);          // Sy.log(), recomputed every step by the reference
    }
    for (int i = threadIdx.x; i < U; i += NT) {
        cst[C_SCALE + i] = prm.act_scale[i];
        cst[C_BIAS + i] = prm.act_bias[i];
    }
    for (int i = threadIdx.x; i < KR; i += NT) cst[C_C0 + i] = prm.rew_c0[i];
    for (int i = threadIdx.x; i < KR * D; i += NT) cst[C_C + (i / D) * SD + (i % D)] = prm.rew_C[i];
    for (int i = threadIdx.x; i < KR * KR; i += NT) {
        int a = i / KR, b = i % KR;
        cst[C_Q + a * SD + b] = prm.rew_Q[a * KR + b];
        cst[C_QS + a * SD + b] = prm.rew_Q[a * KR + b] + prm.rew_Q[b * KR + a];
    }
    for (int i = threadIdx.x; i < U * U; i += NT) {
        int a = i / U, b = i % U;
        cst[C_R + a * SD + b] = prm.rew_R[a * U + b];
        cst[C_RS + a * SD + b] = prm.rew_R[a * U + b] + prm.rew_R[b * U + a];
    }
}

}  // namespace pmb
