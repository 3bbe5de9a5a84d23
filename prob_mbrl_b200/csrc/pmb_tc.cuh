// Tensor-core cluster sweeps (sm_100a): the imagined rollout for the shapes whose hidden x hidden layers are
// genuinely MMA-shaped (BASELINE c4 3x[400] @ 125 particles/GPU, c5 2x[512] @ 250/GPU) and for moment matching of
// the states (c3), where one thread-block cluster holds the whole particle set of a matching group.
//
// A cluster of C = 16 CTAs owns a TILE of up to 128 particles for the whole horizon (M = 128 = one UMMA tile).
// Every hidden x hidden layer  out[128 x W] = act[128 x K] . Wt[K x W]  is COLUMN-SPLIT over the cluster: CTA r
// forms the ns = W/C (padded to 16/32/64) output columns [r ns, (r+1) ns) with tcgen05.mma.kind::tf32, fp32
// accumulation in TMEM.  fp32 parity on the tensor pipe needs the 3-term split the weight-gradient kernel
// validated (pmb_wgrad.cu): x = hi + lo, hi = tf32(x);  C += A_hi B_hi + A_hi B_lo + A_lo B_hi.
//   * weights (B operand): split once per call by tc_pack_kernel into per-CTA hi/lo slices in the K-major
//     no-swizzle canonical layout (8 rows x 16 B core matrices, LBO 128 B, SBO 256 B;
//     profiles/r01_umma_layout_probe.txt) and streamed by TMA (cp.async.bulk + mbarrier) through a shared-memory
//     ring that runs ahead across layer boundaries (weights do not depend on the step's data);
//   * activations (A operand): every CTA's epilogue (TMEM -> registers: bias, ReLU, dropout mask, keep) writes its
//     columns as plain fp32 to a small exchange image in global memory (L2 resident, [k-block][k half][row][4]:
//     coalesced both ways); after ONE cluster barrier TMA streams the whole image through a shared-memory ring, the
//     8 compute warps of every CTA pick their rows up with 128-bit shared loads, split them into TF32 hi/lo IN
//     REGISTERS and park them in TENSOR MEMORY with tcgen05.st -- the MMAs take A from TMEM
//     (tests/csrc/umma_tmem_a_probe.cu), so shared memory sees every activation twice (TMA write, one read) instead
//     of the five times of the first version, which staged hi/lo images and read them per MMA
//     (profiles/r02_tc_timelines_v1_v2.txt).
//   * four more warps drive the pipeline: TMA producers and MMA issuers, each issuer with its own accumulator
//     (A stage full / weight stage full -> 3 MMAs per k-block -> tcgen05.commit frees both).
// The skinny first / last layers (K <= 16 inputs, <= 32 outputs) stay on the FP32 pipe: the first layer is formed
// per CTA for its own columns, the output projection as per-CTA partial sums that meet in global memory
// ([rank][output][particle], fixed-order sum => bit-identical state copies on every CTA, deterministic).
// Moment matching of the states needs no communication at all: every CTA holds the full state tile.
#pragma once
#include "pmb_internal.cuh"

namespace pmb {

constexpr int TC_NT = 256;         // COMPUTE threads per CTA: 8 warps; thread (warp w, lane) owns particle 32 (w%4) + lane
                                   // and the column half w/4 of the CTA's slice
constexpr int TC_NTL = TC_NT + 128; // + four driver warps (TMA producers + MMA issue)
constexpr int TC_KC = 8;           // k-blocks (of 8) per pipeline chunk
constexpr int TC_NSA = 3;          // A stages in tensor memory (max; 2 x 64 columns each: hi | lo)
constexpr int TC_NSW = 4;          // weight stages in shared memory (max)
constexpr int TC_MAXITEMS = 2 * (MAXL - 1) * 16;   // weight chunks per step (schedule table)
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
    int kb_stage, nstage;           // k-blocks per weight stage (= TC_KC), weight stages
    int nsx, nsa;                   // image stages (shared memory), A stages (tensor memory; 4 ns + 128 nsa <= 512)
    int kbmax;                      // k-blocks of an exchange image (= C * ns / 8: every CTA writes all its columns)
    int nop;                        // rows of the partial-sum exchange: max(inputs, raw outputs) over both nets
    TcNet pol, dyn;
    float *ws;                      // workspace base
    const float *wpack;             // tc weight area (hi/lo slices)
    float *xbuf;                    // [ntiles][2][kbmax][k half][128][4] fp32 activation exchange images
    float *opart;                   // [ntiles][2][C][nop][128] partial sums of the skinny projections (2 = pass parity)
    const float *act_scale, *act_bias, *mx, *iSx, *my, *Sy;
    int KR;
    const float *rew_C, *rew_c0, *rew_Q, *rew_R;
    float rew_scale, rew_offset;
    const float *x0;
    float *states, *actions, *rewards;
    const float *g_states, *g_actions, *g_rewards;
    float *dx0;
    float *da_total;                // backward: total dL/da_t [H][N][U] (nullable)
    float *pre;                     // backward: [H][N][2D + 3U] step-local adjoint factors (cluster_bwd_pre_kernel)
    int *status;
    // moment matching of the states (reference utils/rollout.py:20-29,121-132); the whole group lives in the tile
    int mm_states, mm_G, mm_Ng;
    const float *z_mm;
    float *s1pre;                   // [H][N][D] particles before matching
    float *mmstat;                  // [H][G][3*SD + SD*SD] mean, z mean, 1/z std, Cholesky factor
    long long *dbg;
    int dbg_flags;                  // timing experiments only (results are wrong): 1 = skip the MMAs, 2 = skip the image
                                    // loads, 4 = skip the tcgen05.st, 8 = skip the hi/lo split
    // shared-memory carve-up (float offsets)
    int off_cst, off_res, off_xin, off_st, off_aux, off_z, off_mm, off_ring, off_xring;
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

// MMA with the A operand in tensor memory (lane = row, one 32-bit column per tf32 element)
__device__ __forceinline__ void tc_mma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
// 32 consecutive TMEM columns of this warp's 32 lanes <- registers
__device__ __forceinline__ void tc_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
          "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
          "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
          "f"(v[30]), "f"(v[31])
        : "memory");
}

// ----------------------------------------------------------------------------------------
// The layer pipeline of one CTA (12 warps).
//   driver warp 0 (thread 256): TMA producer of the weight ring (runs ahead across layers) and of the image ring
//   image stages  (shared memory, fp32 [k-block][k half][row][4]):  x_full <- complete_tx,  x_empty <- 8 compute warps
//   compute warps: LDS own row share -> TF32 hi/lo split -> tcgen05.st into an A stage of tensor memory
//   A stages      (tensor memory, hi | lo):                         a_full <- 8 compute warps, a_empty <- 4 drivers
//   weight stages (shared memory, hi | lo, UMMA canonical):         w_full <- complete_tx,  w_empty <- 4 drivers
//   driver warps 0..3: one thread each issues the MMAs of the k-blocks kb % 4 == j into ITS OWN accumulator
//   (a single thread sustains one tcgen05.mma per ~100 cycles whatever its size, profiles/r02_tc_timelines_v1_v2.txt;
//   the slices are only N = 16..64 wide, so four issuers keep the tensor pipe fed); the epilogue adds the four
//   accumulators in a fixed order.   done <- 4 drivers: the layer's accumulators are complete.
// ----------------------------------------------------------------------------------------
constexpr int TC_NDRV = 4;         // MMA-issuing driver warps = independent accumulators
constexpr int TC_NSX = 3;          // image stages in shared memory (max)
struct TcBars {
    uint64_t w_full[TC_NSW];
    uint64_t w_empty[TC_NSW];
    uint64_t x_full[TC_NSX];
    uint64_t x_empty[TC_NSX];
    uint64_t a_full[TC_NSA];
    uint64_t a_empty[TC_NSA];
    uint64_t done;
};
struct TcWItem {            // one weight chunk of the per-step schedule (this CTA's slice)
    const float *hi;        // lo at hi + lo_off
    uint32_t lo_off;        // floats
    uint32_t nkb;
};
struct TcPipe {
    // driver 0: weight chunks issued so far, of the whole kernel, per step; next schedule item; image chunks issued
    uint32_t w_issued, w_total, w_items, w_next, x_issued;
    // every thread: chunks consumed so far (weights by the drivers, image / A stages by everybody)
    uint32_t count;
    uint32_t done_parity;
    __device__ __forceinline__ void init(uint32_t items, uint32_t total) {
        w_issued = 0; w_total = total; w_items = items; w_next = 0; x_issued = 0; count = 0; done_parity = 0;
    }
};

// exchange image: element (row m, column k) at (k/8)*1024 + ((k%8)/4)*512 + m*4 + (k%4)
__device__ __forceinline__ void tc_store_img(float *img, int kb, int khalf, int m, float4 v) {
    *reinterpret_cast<float4 *>(img + (size_t)kb * 1024 + khalf * 512 + m * 4) = v;
}

// driver thread j of one layer
__device__ __forceinline__ void tc_drive_layer(const TcParams &prm, float *wring, float *xring, TcBars *bars, TcPipe &pp,
                                               const TcWItem *sched, const float *img, int KB, uint32_t tmem_base, int j) {
    const int ns = prm.ns, nsw = prm.nstage, nsx = prm.nsx, nsa = prm.nsa;
    const int nch = (KB + TC_KC - 1) / TC_KC;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ns >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const uint32_t wlo = (uint32_t)(TC_KC * ns * 8);       // floats: lo half of a weight stage
    const uint32_t tmem_dj = tmem_base + (uint32_t)(j * ns);
    const uint32_t col_a = (uint32_t)(TC_NDRV * ns);
    const uint32_t x_base = pp.count;                       // chunk counter at layer entry
    uint32_t first = 1u;
    for (int c = 0; c < nch; ++c) {
        if (j == 0) {
            // weights: keep nsw chunks in flight, across layer boundaries
            while (pp.w_issued < pp.w_total && pp.w_issued < pp.count + (uint32_t)nsw) {
                const uint32_t s = pp.w_issued % nsw;
                if (pp.w_issued >= (uint32_t)nsw) mbar_wait(&bars->w_empty[s], ((pp.w_issued / nsw) - 1u) & 1u);
                const TcWItem it = sched[pp.w_next];
                const uint32_t bytes = it.nkb * (uint32_t)(ns * 32);
                float *dst = wring + (size_t)s * prm.stage_floats;
                mbar_expect_tx(&bars->w_full[s], 2u * bytes);
                tma_bulk_g2s(dst, it.hi, bytes, &bars->w_full[s]);
                tma_bulk_g2s(dst + wlo, it.hi + it.lo_off, bytes, &bars->w_full[s]);
                ++pp.w_issued;
                if (++pp.w_next == pp.w_items) pp.w_next = 0;
            }
            // image of THIS layer: keep nsx chunks in flight
            while (pp.x_issued < x_base + (uint32_t)nch && pp.x_issued < pp.count + (uint32_t)nsx) {
                const uint32_t s = pp.x_issued % nsx;
                if (pp.x_issued >= (uint32_t)nsx) mbar_wait(&bars->x_empty[s], ((pp.x_issued / nsx) - 1u) & 1u);
                const int cc = (int)(pp.x_issued - x_base);
                const uint32_t bytes = (uint32_t)min(TC_KC, KB - cc * TC_KC) * 4096u;
                mbar_expect_tx(&bars->x_full[s], bytes);
                tma_bulk_g2s(xring + (size_t)s * (TC_KC * 1024), img + (size_t)cc * TC_KC * 1024, bytes, &bars->x_full[s]);
                ++pp.x_issued;
            }
        }
        const int nkb = min(TC_KC, KB - c * TC_KC);
        const uint32_t sw = pp.count % nsw, sa = pp.count % nsa;
        mbar_wait(&bars->w_full[sw], (pp.count / nsw) & 1u);
        mbar_wait(&bars->a_full[sa], (pp.count / nsa) & 1u);
        tc_fence_after();
        const uint32_t wb = smem_u32(wring + (size_t)sw * prm.stage_floats);
        const uint32_t ta = tmem_base + col_a + sa * 128u;
        for (int k = j; k < nkb; k += TC_NDRV) {
            if (prm.dbg_flags & 1) break;
            const uint64_t dBh = tc_desc(wb + (uint32_t)(k * ns * 32));
            const uint64_t dBl = tc_desc(wb + wlo * 4u + (uint32_t)(k * ns * 32));
            tc_mma_tf32_ta(tmem_dj, ta + 8u * k, dBh, idesc, first ? 0u : 1u);
            tc_mma_tf32_ta(tmem_dj, ta + 8u * k, dBl, idesc, 1u);
            tc_mma_tf32_ta(tmem_dj, ta + 64u + 8u * k, dBh, idesc, 1u);
            first = 0u;
        }
        tc_commit(&bars->a_empty[sa]);
        tc_commit(&bars->w_empty[sw]);
        ++pp.count;
    }
    tc_commit(&bars->done);
}

// compute warps: chunk -> this thread's share (row m, k-blocks 4*half .. 4*half+3) of an image stage -> TF32 hi / lo
// -> the chunk's A stage of tensor memory
__device__ __forceinline__ void tc_img_produce(const TcParams &prm, const float *xring, TcBars *bars, TcPipe &pp,
                                               uint32_t tmem_lane_base, int m, int half) {
    const int nsx = prm.nsx, nsa = prm.nsa;
    const uint32_t sx = pp.count % nsx, sa = pp.count % nsa;
    float x[32];
    mbar_wait(&bars->x_full[sx], (pp.count / nsx) & 1u);
    {
        const float *q = xring + (size_t)sx * (TC_KC * 1024) + (size_t)(4 * half) * 1024 + m * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 a = *reinterpret_cast<const float4 *>(q + j * 1024);
            const float4 b = *reinterpret_cast<const float4 *>(q + j * 1024 + 512);
            x[8 * j] = a.x; x[8 * j + 1] = a.y; x[8 * j + 2] = a.z; x[8 * j + 3] = a.w;
            x[8 * j + 4] = b.x; x[8 * j + 5] = b.y; x[8 * j + 6] = b.z; x[8 * j + 7] = b.w;
        }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&bars->x_empty[sx]);      // the stage may be refilled
    if (pp.count >= (uint32_t)nsa) {
        mbar_wait(&bars->a_empty[sa], ((pp.count / nsa) - 1u) & 1u);
        tc_fence_after();
    }
    const uint32_t ta = tmem_lane_base + (uint32_t)(TC_NDRV * prm.ns) + sa * 128u + 32u * (uint32_t)half;
    {
        float hi[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            hi[e] = tc_tf32_hi(x[e]);
            x[e] -= hi[e];
        }
        tc_st32(ta, hi);
    }
    tc_st32(ta + 64u, x);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&bars->a_full[sa]);
    ++pp.count;
}

// One hidden x hidden layer of this CTA: D[128 x ns] = A[128 x 8 KB] . B[ns x 8 KB]^T over KB k-blocks, A from the
// fp32 exchange image `img`.  Called by all 384 threads; returns when the accumulators are complete.
__device__ __forceinline__ void tc_wide_layer(const TcParams &prm, float *wring, float *xring, TcBars *bars, TcPipe &pp,
                                              const TcWItem *sched, const float *img, int KB, uint32_t tmem_base, int m,
                                              int half, long long *dbgp = nullptr) {
    const int warp = threadIdx.x >> 5;
    const int nch = (KB + TC_KC - 1) / TC_KC;
    int dbi = 0;
#define TC_WMARK() do { if (dbgp && (threadIdx.x & 31) == 0 && dbi < 24) dbgp[(threadIdx.x >> 5) * 24 + dbi++] = clock64(); } while (0)
    TC_WMARK();
    if (warp >= 8) {
        if ((threadIdx.x & 31) == 0) tc_drive_layer(prm, wring, xring, bars, pp, sched, img, KB, tmem_base, warp - 8);
        else pp.count += (uint32_t)nch;
        __syncwarp();
    } else {
        const uint32_t lane_base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
            tc_img_produce(prm, xring, bars, pp, lane_base, m, half);
            TC_WMARK();
        }
    }
    mbar_wait(&bars->done, pp.done_parity);
    pp.done_parity ^= 1u;
    tc_fence_after();
    TC_WMARK();
#undef TC_WMARK
}

// sum of the drivers' accumulators, columns [col0, col0 + HW) of this warp's 32 TMEM lanes (fixed order)
template <int HW>
__device__ __forceinline__ void tc_ld_acc_sum(uint32_t taddr, int ns, int nacc, float (&v)[HW]) {
    tc_ld_acc<HW>(taddr, v);
    for (int j = 1; j < nacc; ++j) {
        float w[HW];
        tc_ld_acc<HW>(taddr + (uint32_t)(j * ns), w);
#pragma unroll
        for (int c = 0; c < HW; ++c) v[c] += w[c];
    }
}

// this CTA's weight-chunk schedule of one step, in consumption order (built once by the driver thread)
__device__ __forceinline__ uint32_t tc_build_schedule(const TcParams &prm, const TcNet *const (&order)[2], bool reverse, int rank,
                                                      TcWItem *sched) {
    uint32_t n = 0;
    const int ns = prm.ns;
    for (int w = 0; w < 2; ++w) {
        const TcNet &net = *order[w];
        for (int i = 1; i < net.L; ++i) {
            const int l = reverse ? net.L - i : i;                         // reverse sweep walks l = L-1 .. 1
            const int KB = reverse ? net.kb[l] : net.kb[l - 1];
            const float *base = prm.wpack + net.wp_off[l] + (size_t)rank * 2 * KB * ns * 8;
            for (int c = 0; c * TC_KC < KB; ++c) {
                sched[n].hi = base + (size_t)c * TC_KC * ns * 8;
                sched[n].lo_off = (uint32_t)(KB * ns * 8);
                sched[n].nkb = (uint32_t)min(TC_KC, KB - c * TC_KC);
                ++n;
            }
        }
    }
    return n;
}

// per-warp arrival marks of one step (profiling aid, pmb_tuning.reserved[2..3]): dbg[i * 8 + warp]
#define TC_MARK(i) do { if (dbg_step && (threadIdx.x & 31) == 0 && threadIdx.x < TC_NT) prm.dbg[(i) * 8 + (threadIdx.x >> 5)] = clock64(); } while (0)

// Sum of the C = 16 per-CTA partials of every (output o, particle) item, in rank order, for all items of the
// exchange block [rank][o][128]: the loads of up to 4 items (64 L2 requests) are in flight before the first add.
__device__ __forceinline__ void tc_reduce_partials(const float *opart, int nitems, int nop, const float *bias, float *out) {
    const int tid = threadIdx.x;
    if (tid >= TC_NT) return;
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
