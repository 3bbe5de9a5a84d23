// Tensor-core cluster sweeps (sm_100a): the imagined rollout for the shapes whose hidden x hidden layers are
// genuinely MMA-shaped (BASELINE c4 3x[400] @ 125 particles/GPU, c5 2x[512] @ 250/GPU) and for moment matching of
// the states (c3), where one thread-block cluster holds the whole particle set of a matching group.
//
// A cluster of C = 16 CTAs owns a TILE of up to 128 particles for the whole horizon (M = 128 = one UMMA tile).
// Every hidden x hidden layer  out[128 x W] = act[128 x K] . Wt[K x W]  is COLUMN-SPLIT over the cluster: CTA r
// forms the ns = W/C (padded to 16/32/64) output columns [r ns, (r+1) ns) with tcgen05.mma.kind::tf32, fp32
// accumulation in TMEM.  fp32 parity on the tensor pipe needs the 3-term split the weight-gradient kernel
// validated (pmb_wgrad.cu): x = hi + lo, hi = tf32(x);  C += A_hi B_hi + A_hi B_lo + A_lo B_hi.
//   * weights: split once per call by tc_pack_kernel into per-CTA hi/lo slices in the K-major no-swizzle
//     canonical layout (8 rows x 16 B core matrices, LBO 128 B, SBO 256 B; profiles/r01_umma_layout_probe.txt);
//   * activations: every CTA's epilogue (TMEM -> registers: bias, ReLU, dropout mask, keep) writes its columns,
//     split into hi/lo and already in the canonical layout, to a small exchange image in global memory (L2
//     resident, 2 x 512 KB per tile for 512-wide nets); after ONE cluster barrier every CTA streams the whole
//     image + its weight slice through a TMA (cp.async.bulk) + mbarrier ring, chunk by chunk, while one elected
//     thread issues the MMAs -- the tensor core reads shared memory, nothing is staged through registers.
// The skinny first / last layers (K <= 16 inputs, <= 32 outputs) stay on the FP32 pipe: the first layer is formed
// per CTA for its own columns, the output projection as per-CTA partial sums that meet in global memory
// ([rank][particle][output], fixed-order sum => bit-identical state copies on every CTA, deterministic).
// Moment matching of the states needs no communication at all: every CTA holds the full state tile.
#pragma once
#include "pmb_internal.cuh"

namespace pmb {

constexpr int TC_NT = 256;         // threads per CTA: 8 warps; thread (warp w, lane) owns particle 32 (w%4) + lane and
                                   // the column half w/4 of the CTA's slice
constexpr int TC_C = 16;           // CTAs per cluster
constexpr int TC_M = 128;          // particles per tile (UMMA M)
constexpr int TC_MAXNS = 64;       // widest column slice per CTA
constexpr int TC_NOUT = 32;        // raw outputs of a net (max)
constexpr int TC_SDP = 17;         // row stride of the per-particle shared-memory tiles (thread-per-row, conflict-free)
constexpr int TC_MAXH = MAXL - 1;  // hidden layers per net (max)

struct TcNet {
    int L;                          // hidden layers (>= 1)
    int nin, nout;                  // net inputs (<= 16), raw outputs (<= 32)
    int width[MAXL];                // hidden widths
    int npad[MAXL];                 // row stride of the saved / mask / delta arrays: width rounded to 4
    int kb[MAXL];                   // k-blocks (of 8) of hidden l as a reduction axis
    const float *W_first;           // [width0][nin]        fp32, as stored by nn.Linear
    const float *W_last;            // [nout][width_{L-1}]  fp32
    long long bias_off[MAXL + 1];   // padded biases in the workspace (index L = output projection), -1 = none
    long long mask_off[MAXL];       // dropout masks [N][npad_l] in the workspace, -1 = none
    float keep_inv[MAXL];
    long long saved_off[MAXL];      // post-dropout activations [H][N][npad_l]
    long long delta_off[MAXL + 1];  // policy: adjoints kept for the weight gradient ([L] = raw outputs), -1 = none
    long long raw_off;              // raw outputs [H][N][nout]
    long long wp_off[MAXL];         // l >= 1: hi/lo slices of linear l for THIS sweep direction, floats inside the tc
                                    // weight area: [rank][hi | lo][kb][ns x 8]
    int has_density;
    float lmax;
    const float *z;
    long long zstride;
    // shared-memory offsets (floats) of the resident fp32 operands of this CTA's columns
    int s_wfirst;                   // [16][ns]   first-layer matrix, k-major (rows >= nin are zero)
    int s_wlast;                    // [TC_NOUT][ns] output projection
    int s_bias;                     // [L + 1][max(ns, TC_NOUT)] biases (forward)
};

struct TcParams {
    int N, H, D, U;
    int C;                          // CTAs per cluster
    int TP;                         // particles per tile (<= 128)
    int ntiles;
    int ns;                         // columns per CTA (16, 32 or 64)
    int kb_stage, nstage;           // k-blocks per ring stage, ring stages
    int kbmax;                      // k-blocks of an exchange image (= C * ns / 8: every CTA writes all its columns)
    int nop;                        // rows of the partial-sum exchange: max(inputs, raw outputs) over both nets
    TcNet pol, dyn;
    float *ws;                      // workspace base
    const float *wpack;             // tc weight area (hi/lo slices)
    float *xbuf;                    // [ntiles][2][hi | lo][kbmax][128 x 8] activation exchange images
    float *opart;                   // [ntiles][2][C][nop][128] partial sums of the skinny projections (2 = pass parity)
    const float *act_scale, *act_bias, *mx, *iSx, *my, *Sy;
    int KR;
    const float *rew_C, *rew_c0, *rew_Q, *rew_R;
    float rew_scale, rew_offset;
    const float *x0;
    float *states, *actions, *rewards;
    const float *g_states, *g_actions, *g_rewards;
    float *dx0;
    float *pre;                     // backward: [H][N][2D + 3U] step-local adjoint factors (cluster_bwd_pre_kernel)
    int *status;
    // moment matching of the states (reference utils/rollout.py:20-29,121-132); the whole group lives in the tile
    int mm_states, mm_G, mm_Ng;
    const float *z_mm;
    float *s1pre;                   // [H][N][D] particles before matching
    float *mmstat;                  // [H][G][3*SD + SD*SD] mean, z mean, 1/z std, Cholesky factor
    long long *dbg;
    // shared-memory carve-up (float offsets)
    int off_cst, off_res, off_xin, off_st, off_aux, off_ring;
    int stage_floats;
    int smem_floats;
};

struct TcPackJob {
    const float *W;       // [out][in]
    int out, in;
    int transpose;        // 0: B[n][k] = W[n][k] (forward);  1: B[n][k] = W[k][n] (reverse sweep)
    long long dst_off;    // floats inside the tc weight area
    int kb;               // k-blocks of the reduction axis
};
struct TcPackJobs {
    int n, C, ns;
    TcPackJob job[4 * MAXL];
};

// ----------------------------------------------------------------------------------------
// PTX helpers (cluster, tcgen05)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t tc_cluster_id() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
// cluster-wide barrier; release/acquire at cluster scope orders the global-memory exchange (activation image,
// partial sums) written before it against the reads after it
__device__ __forceinline__ void tc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {     // K-major, no swizzle, LBO 128 B, SBO 256 B
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ float tc_tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// accumulator columns [col0, col0 + HW) of this warp's 32 TMEM lanes -> registers
template <int HW>
__device__ __forceinline__ void tc_ld_acc(uint32_t taddr, float (&v)[HW]) {
    static_assert(HW == 8 || HW == 16 || HW == 32, "column half of the slice");
    uint32_t r[HW];
#pragma unroll
    for (int c = 0; c < HW; c += 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[c]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]),
                       "=r"(r[c + 6]), "=r"(r[c + 7])
                     : "r"(taddr + (uint32_t)c));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < HW; ++c) v[c] = __uint_as_float(r[c]);
}

// ----------------------------------------------------------------------------------------
// The operand ring of one CTA: stages of [A_hi | A_lo | W_hi | W_lo] k-block groups filled by TMA bulk copies
// (thread 0 = producer) and consumed by the MMAs (thread 32 = issuer).  tcgen05.commit hands a stage back.
// ----------------------------------------------------------------------------------------
constexpr int TC_MAXSTAGE = 4;
struct TcBars {
    uint64_t full[TC_MAXSTAGE];
    uint64_t empty[TC_MAXSTAGE];
    uint64_t done;
};
struct TcRing {
    int stage;            // producer / issuer: next stage
    uint32_t parity;      // issuer: parity of the next full wait; producer: of the next empty wait
    uint32_t issued;      // producer: chunks issued so far
    uint32_t done_parity; // all threads
    __device__ __forceinline__ void init() { stage = 0; parity = 0; issued = 0; done_parity = 0; }
};

// One hidden x hidden layer of this CTA: D[128 x ns] = A[128 x 8 KB] . B[ns x 8 KB]^T over KB k-blocks.
//   a_img : exchange image of the layer input in global memory: hi at a_img, lo at a_img + lo_off (floats)
//   w_sl  : this CTA's weight slices: hi at w_sl, lo at w_sl + KB * ns * 8
// Called by all 256 threads; returns when the accumulator is complete and visible to tcgen05.ld.
__device__ __forceinline__ void tc_wide_layer(const TcParams &prm, float *ring_base, TcBars *bars, TcRing &rg,
                                              const float *a_img, long long lo_off, const float *w_sl, int KB,
                                              uint32_t tmem_d) {
    const int tid = threadIdx.x;
    const int ns = prm.ns, KBS = prm.kb_stage;
    const int nchunks = (KB + KBS - 1) / KBS;
    const int a_st = KBS * 1024, w_st = KBS * ns * 8;      // floats of one operand half in a stage
    if (tid == 0) {
        // ---- producer ----
        for (int c = 0; c < nchunks; ++c) {
            const int nkb = min(KBS, KB - c * KBS);
            if (rg.issued >= (uint32_t)prm.nstage) mbar_wait(&bars->empty[rg.stage], rg.parity);
            float *st = ring_base + (size_t)rg.stage * prm.stage_floats;
            const uint32_t ab = (uint32_t)nkb * 4096u, wb = (uint32_t)(nkb * ns * 32);
            mbar_expect_tx(&bars->full[rg.stage], 2u * ab + 2u * wb);
            const float *ag = a_img + (size_t)c * a_st;
            const float *wg = w_sl + (size_t)c * w_st;
            tma_bulk_g2s(st, ag, ab, &bars->full[rg.stage]);
            tma_bulk_g2s(st + a_st, ag + lo_off, ab, &bars->full[rg.stage]);
            tma_bulk_g2s(st + 2 * a_st, wg, wb, &bars->full[rg.stage]);
            tma_bulk_g2s(st + 2 * a_st + w_st, wg + (size_t)KB * ns * 8, wb, &bars->full[rg.stage]);
            ++rg.issued;
            if (++rg.stage == prm.nstage) {
                rg.stage = 0;
                if (rg.issued > (uint32_t)prm.nstage) rg.parity ^= 1u;
            }
        }
    } else if (tid == 32) {
        // ---- MMA issuer ----
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ns >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
        for (int c = 0; c < nchunks; ++c) {
            const int nkb = min(KBS, KB - c * KBS);
            mbar_wait(&bars->full[rg.stage], rg.parity);
            tc_fence_after();
            const uint32_t s0 = smem_u32(ring_base + (size_t)rg.stage * prm.stage_floats);
            for (int k = 0; k < nkb; ++k) {
                const uint64_t dAh = tc_desc(s0 + (uint32_t)k * 4096u);
                const uint64_t dAl = tc_desc(s0 + (uint32_t)(a_st * 4) + (uint32_t)k * 4096u);
                const uint64_t dBh = tc_desc(s0 + (uint32_t)(2 * a_st * 4) + (uint32_t)(k * ns * 32));
                const uint64_t dBl = tc_desc(s0 + (uint32_t)((2 * a_st + w_st) * 4) + (uint32_t)(k * ns * 32));
                tc_mma_tf32(tmem_d, dAh, dBh, idesc, (c | k) ? 1u : 0u);
                tc_mma_tf32(tmem_d, dAh, dBl, idesc, 1u);
                tc_mma_tf32(tmem_d, dAl, dBh, idesc, 1u);
            }
            tc_commit(&bars->empty[rg.stage]);       // the stage is free once these MMAs retired
            if (++rg.stage == prm.nstage) {
                rg.stage = 0;
                rg.parity ^= 1u;
            }
        }
        tc_commit(&bars->done);
    }
    __syncwarp();
    mbar_wait(&bars->done, rg.done_parity);
    rg.done_parity ^= 1u;
    tc_fence_after();
}

// hi/lo split of 4 consecutive k values of one particle row -> the exchange image (canonical K-major layout):
// element (row m, k) of k-block kb sits at kb*1024 + (m/8)*64 + ((k%8)/4)*32 + (m%8)*4 + (k%4)
__device__ __forceinline__ void tc_store_hilo(float *img_hi, long long lo_off, int kb, int khalf, int m, float4 v) {
    const float4 h = make_float4(tc_tf32_hi(v.x), tc_tf32_hi(v.y), tc_tf32_hi(v.z), tc_tf32_hi(v.w));
    float *p = img_hi + (size_t)kb * 1024 + (m >> 3) * 64 + khalf * 32 + (m & 7) * 4;
    *reinterpret_cast<float4 *>(p) = h;
    *reinterpret_cast<float4 *>(p + lo_off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// per-warp arrival marks of one step (profiling aid, pmb_tuning.reserved[2..3]): dbg[i * 8 + warp]
#define TC_MARK(i) do { if (dbg_step && (threadIdx.x & 31) == 0) prm.dbg[(i) * 8 + (threadIdx.x >> 5)] = clock64(); } while (0)

// Sum of the C = 16 per-CTA partials of every (output o, particle) item, in rank order, for all items of the
// exchange block [rank][o][128]: the loads of up to 4 items (64 L2 requests) are in flight before the first add.
__device__ __forceinline__ void tc_reduce_partials(const float *opart, int nitems, int nop, const float *bias, float *out) {
    const int tid = threadIdx.x;
    const size_t rstride = (size_t)nop * TC_M;
#pragma unroll 1
    for (int i0 = tid; i0 < nitems; i0 += 4 * TC_NT) {
        float v[4][TC_C];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * TC_NT;
#pragma unroll
            for (int r = 0; r < TC_C; ++r) v[q][r] = i < nitems ? __ldcg(opart + r * rstride + i) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * TC_NT;
            if (i < nitems) {
                float a = bias ? bias[i / TC_M] : 0.f;
#pragma unroll
                for (int r = 0; r < TC_C; ++r) a += v[q][r];
                out[i] = a;
            }
        }
    }
}

cudaError_t launch_tc_pack(const TcPackJobs &jobs, float *wpack, cudaStream_t stream);
cudaError_t launch_tc_fwd(const TcParams &prm, cudaStream_t stream);
cudaError_t launch_tc_bwd(const TcParams &prm, cudaStream_t stream);
int tc_max_active_clusters(int C, int smem_bytes);

}  // namespace pmb
