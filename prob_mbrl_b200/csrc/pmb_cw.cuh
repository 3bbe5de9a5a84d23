// Wide cluster-resident sweeps (sm_100a): the imagined rollout of two-hidden-layer nets up to 512 units wide with
// ALL weights resident in the shared memory of a 16-CTA thread-block cluster (non-portable cluster size), for up to
// 36 particles per cluster.  Same decomposition as pmb_cluster.cuh (thin layer redundantly in every CTA, wide layer
// column-split over the CTAs, narrow layer as per-CTA partial sums), re-balanced for throughput instead of latency:
//
//   * B200 holds 7 co-resident 16-CTA clusters (tests/csrc/cluster_occ_probe.cu), so one GPU's shard of the
//     2x[512] / 250-particle configuration is 36 particles per cluster: slots 0..31 in octets + 4 singles.
//   * wide layer: warp = k-slice (16 slices), lane = (particle octet, column quad): 9 particles x 4 columns per
//     thread, one LDS.128 of weights + two LDS.128 + one LDS.32 of activations feed 18 FFMA2; the 16 slices are
//     added through shared memory (the partial-sum buffer aliases the activation tile).
//   * the dropout masks of the full-width hidden layer live in registers as bit words (bit = particle slot); the
//     ReLU/dropout gates the reverse sweep needs are kept as bit words as well (64 + 64 bytes per particle, step and
//     net instead of 2 x 2 KB of activations), so the dynamics net stores no activations at all.
//   * exchange in two hops over DSMEM (st.async + mbarrier complete_tx): every CTA sends its partial sums of a
//     particle's raw outputs to the particle's OWNER CTA (p mod 16), the owner adds the 16 partials in a fixed order,
//     runs the per-particle stage (density sample, tanh squash, scalers / their adjoints) once, and broadcasts the
//     next layer's input column to all 16 CTAs.  Mailboxes are 3 KB instead of 16 x 36 x 16 floats.
#pragma once
#include "pmb_cluster.cuh"

namespace pmb {

constexpr int CW_C = 16;                  // CTAs per cluster
constexpr int CW_NT = 512;                // threads per CTA
constexpr int CW_NW = CW_NT / 32;         // warps = k-slices of the wide layer
constexpr int CW_PS = 36;                 // particle slots per cluster = row stride of the x / activation tiles
constexpr int CW_HS = 32;                 // columns of the wide layer per CTA (max)
constexpr int CW_TW = 512;                // widest hidden layer
constexpr int CW_NO = 16;                 // thin inputs / narrow outputs (max)
constexpr int CW_OWN = 3;                 // particles owned by one CTA (max): p = rank + 16 * lp
constexpr int CW_MB = CW_OWN * CW_C * CW_NO;   // floats of one mailbox: [lp][sender rank][16]
constexpr int CW_XT = CW_NO * CW_PS;      // floats of one x tile [16][36]
constexpr int CW_G2 = CW_PS * CW_C;       // words of one gate tile [36][16]

// st.async of one float into a peer CTA's shared memory + 4 bytes on the peer's mbarrier
__device__ __forceinline__ void cw_st_async_f32(uint32_t daddr, float v, uint32_t dbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(daddr), "f"(v), "r"(dbar)
                 : "memory");
}

// resident operands of one net: this CTA's [tW][32] column slice of the wide matrix, its columns of the narrow
// matrix as [4][32][4] (o >> 2, column, o & 3), its slice of the wide bias and the narrow bias (forward)
__device__ __forceinline__ void cw_load_net(const ClusterParams &prm, const CNet &n, float *smem, int rank, bool fwd) {
    const int tid = threadIdx.x;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < n.tW * (CW_HS / 4); i += CW_NT) {
        const int k = i >> 3, c4 = i & 7;
        const int gc = rank * n.hs + 4 * c4;
        float4 v = z4;
        if (4 * c4 < n.hs && gc < n.wN) v = __ldg(reinterpret_cast<const float4 *>(prm.wpack + n.w_goff + (long long)k * n.wN + gc));
        *reinterpret_cast<float4 *>(smem + n.s_ww + k * CW_HS + 4 * c4) = v;
    }
    for (int i = tid; i < CW_NO * CW_HS; i += CW_NT) {
        const int o = i / CW_HS, c = i - o * CW_HS;
        const int gc = rank * n.hs + c;
        float v = 0.f;
        if (o < n.nN && c < n.hs && gc < n.wN) v = __ldg(prm.wpack + n.n_goff + (long long)o * n.wN + gc);
        smem[n.s_nwt + (o >> 2) * (CW_HS * 4) + c * 4 + (o & 3)] = v;
    }
    if (fwd) {
        for (int c = tid; c < CW_HS; c += CW_NT) {
            const int gc = rank * n.hs + c;
            smem[n.s_wb + c] = (n.wb_off >= 0 && c < n.hs && gc < n.wN) ? __ldg(prm.ws + n.wb_off + gc) : 0.f;
        }
        for (int o = tid; o < CW_NO; o += CW_NT) smem[n.s_nb + o] = (n.nb_off >= 0 && o < n.nN) ? __ldg(prm.ws + n.nb_off + o) : 0.f;
    }
}

// per-thread constants of one net's thin layer, fixed for the whole horizon: thread = column j of the thin output
template <int TK>
struct CwThin {
    float w[TK];          // column j of the thin matrix (rows >= tK are zero)
    float b;              // bias (forward)
    float kinv;           // 1 / keep of the layer the thin op produces (forward) / gates (reverse)
    uint32_t m0, m1;      // forward: dropout-mask bits of column j for the slots 0..31 / 32..35
    bool on;              // j < tW
    __device__ __forceinline__ void init(const ClusterParams &prm, const CNet &n, int j, int n0, bool fwd) {
        on = j < n.tW;
#pragma unroll
        for (int k = 0; k < TK; ++k) w[k] = (on && k < n.tK) ? __ldg(prm.wpack + n.t_goff + (long long)k * n.tW + j) : 0.f;
        b = (fwd && on && n.tb_off >= 0) ? __ldg(prm.ws + n.tb_off + j) : 0.f;
        kinv = n.tkeep_inv;
        m0 = m1 = 0u;
        if (fwd && on) {
            for (int p = 0; p < CW_PS; ++p) {
                const int nn = min(n0 + p, prm.N - 1);
                const float v = n.tm_off >= 0 ? __ldg(prm.ws + n.tm_off + (long long)nn * n.tW + j) : 1.f;
                if (v != 0.f && v != 1.f) __trap();     // the caller declared the masks binary (pmb_problem.masks_binary)
                if (v != 0.f) {
                    if (p < 32) m0 |= 1u << p;
                    else m1 |= 1u << (p - 32);
                }
            }
        }
    }
};

// wide layer: acc[i][0..1] = columns (4 cq .. 4 cq + 3) of particle (8 po + i) for i < 8, of particle (32 + po) for i = 8,
// summed over this warp's k-slice (k = warp, warp + 16, ...)
__device__ __forceinline__ void cw_wide_accum(const float *__restrict__ ww, int K, const float *__restrict__ act,
                                              float2 (&acc)[9][2]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cq = lane & 7, po = lane >> 3;
    const float2 z2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i][0] = acc[i][1] = z2;
    const float *wp = ww + w * CW_HS + 4 * cq;
    const float *ap = act + w * CW_PS + 8 * po;
    const float *as = act + w * CW_PS + 32 + po;
    const int n = (K - w + CW_NW - 1) / CW_NW;
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
        const float4 wv = *reinterpret_cast<const float4 *>(wp);
        const float4 xa = *reinterpret_cast<const float4 *>(ap);
        const float4 xb = *reinterpret_cast<const float4 *>(ap + 4);
        const float xs = *as;
        wp += CW_NW * CW_HS;
        ap += CW_NW * CW_PS;
        as += CW_NW * CW_PS;
        const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
        acc[0][0] = cl_fma2(xa.x, w01, acc[0][0]); acc[0][1] = cl_fma2(xa.x, w23, acc[0][1]);
        acc[1][0] = cl_fma2(xa.y, w01, acc[1][0]); acc[1][1] = cl_fma2(xa.y, w23, acc[1][1]);
        acc[2][0] = cl_fma2(xa.z, w01, acc[2][0]); acc[2][1] = cl_fma2(xa.z, w23, acc[2][1]);
        acc[3][0] = cl_fma2(xa.w, w01, acc[3][0]); acc[3][1] = cl_fma2(xa.w, w23, acc[3][1]);
        acc[4][0] = cl_fma2(xb.x, w01, acc[4][0]); acc[4][1] = cl_fma2(xb.x, w23, acc[4][1]);
        acc[5][0] = cl_fma2(xb.y, w01, acc[5][0]); acc[5][1] = cl_fma2(xb.y, w23, acc[5][1]);
        acc[6][0] = cl_fma2(xb.z, w01, acc[6][0]); acc[6][1] = cl_fma2(xb.z, w23, acc[6][1]);
        acc[7][0] = cl_fma2(xb.w, w01, acc[7][0]); acc[7][1] = cl_fma2(xb.w, w23, acc[7][1]);
        acc[8][0] = cl_fma2(xs, w01, acc[8][0]);   acc[8][1] = cl_fma2(xs, w23, acc[8][1]);
    }
}
// partial sums of this warp's k-slice -> red[slice][particle][32 columns] (aliases the activation tile: the caller
// synchronises the CTA before and after)
__device__ __forceinline__ void cw_wide_park(float *__restrict__ red, const float2 (&acc)[9][2]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cq = lane & 7, po = lane >> 3;
    float *r = red + ((w * CW_PS + 8 * po) << 5) + 4 * cq;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4 *>(r + (i << 5)) = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
    *reinterpret_cast<float4 *>(red + ((w * CW_PS + 32 + po) << 5) + 4 * cq) =
        make_float4(acc[8][0].x, acc[8][0].y, acc[8][1].x, acc[8][1].y);
}
// finished value of (particle p, column = lane): the 16 k-slices in a fixed order
__device__ __forceinline__ float cw_wide_reduce(const float *__restrict__ red, int p) {
    const float *r = red + (p << 5) + (threadIdx.x & 31);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int s = 0; s < CW_NW; s += 4) {
        s0 += r[(s + 0) * CW_PS * 32];
        s1 += r[(s + 1) * CW_PS * 32];
        s2 += r[(s + 2) * CW_PS * 32];
        s3 += r[(s + 3) * CW_PS * 32];
    }
    return (s0 + s1) + (s2 + s3);
}

// Narrow layer of one particle inside the wide layer's epilogue: `v` = the lane's finished hidden value (column = lane
// of this CTA's slice, 0 on idle lanes).  out[o] = sum_lanes v * nw[o][lane] for o < NV with a reduce-scatter butterfly;
// the lanes q < NV / 4 end up holding the 16-byte chunk of the outputs 4q .. 4q + 3.
template <int NV>
__device__ __forceinline__ float4 cw_narrow_chunk(float v, const float *__restrict__ nwt) {
    const int lane = threadIdx.x & 31;
    float pr[NV];
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
        const float4 w = *reinterpret_cast<const float4 *>(nwt + i * (CW_HS * 4) + lane * 4);
        pr[4 * i] = v * w.x;
        pr[4 * i + 1] = v * w.y;
        pr[4 * i + 2] = v * w.z;
        pr[4 * i + 3] = v * w.w;
    }
    int m = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1, m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = up ? pr[n / 2 + i] : pr[i];
            const float give = up ? pr[i] : pr[n / 2 + i];
            pr[i] = keep + __shfl_xor_sync(0xffffffffu, give, m);
        }
    }
#pragma unroll
    for (; m >= 1; m >>= 1) pr[0] += __shfl_xor_sync(0xffffffffu, pr[0], m);
    // output o sits in every lane of its group of s = 32 / NV lanes
    constexpr int s = 32 / NV;
    const int src = (lane & (NV / 4 - 1)) * 4 * s;
    return make_float4(__shfl_sync(0xffffffffu, pr[0], src), __shfl_sync(0xffffffffu, pr[0], src + s),
                       __shfl_sync(0xffffffffu, pr[0], src + 2 * s), __shfl_sync(0xffffffffu, pr[0], src + 3 * s));
}
// ... and its way to the owner of the particle: mailbox [lp][sender rank][16] of CTA (p & 15)
__device__ __forceinline__ void cw_narrow_send(float v, const float *__restrict__ nwt, int nN, bool send_ok, int p, int rank,
                                               uint32_t mbox_saddr, uint32_t bar_saddr, uint32_t wstride) {
    const int lane = threadIdx.x & 31;
    float4 out;
    if (nN <= 4) out = cw_narrow_chunk<4>(v, nwt);
    else if (nN <= 8) out = cw_narrow_chunk<8>(v, nwt);
    else out = cw_narrow_chunk<16>(v, nwt);
    if (send_ok && 4 * lane < nN) {
        const uint32_t owner = (uint32_t)(p & (CW_C - 1));
        const uint32_t off = (uint32_t)((((p >> 4) * CW_C + rank) * CW_NO + 4 * lane) * 4);
        const uint32_t a0 = cl_mapa(mbox_saddr + off, 0), b0 = cl_mapa(bar_saddr, 0);
        cl_st_async_v4(a0 + owner * wstride, out, b0 + owner * wstride);
    }
}
// owner side: value o of owned particle lp, the 16 partials added in a fixed pairwise order
__device__ __forceinline__ float cw_gather(const float *mbox, int lp, int o) {
    const float *q = mbox + lp * (CW_C * CW_NO) + o;
    float v[CW_C];
#pragma unroll
    for (int r = 0; r < CW_C; ++r) v[r] = q[r * CW_NO];
#pragma unroll
    for (int w = 1; w < CW_C; w <<= 1)
#pragma unroll
        for (int r = 0; r + w < CW_C; r += 2 * w) v[r] += v[r + w];
    return v[0];
}

// per-warp arrival marks (lane 0 of every warp) of cluster 0 / rank 0 at step H/2: dbg[i * 16 + warp]
#define CW_MARK(i) do { if (dbg_step && (threadIdx.x & 31) == 0) prm.dbg[(i) * 16 + (threadIdx.x >> 5)] = clock64(); } while (0)

cudaError_t launch_cw_fwd(const ClusterParams &prm, int nclusters, cudaStream_t stream);
cudaError_t launch_cw_bwd(const ClusterParams &prm, int nclusters, cudaStream_t stream);
int cw_max_active();

}  // namespace pmb
