// C ABI of libpmb_b200.so (declared in include/pmb_b200.h) + the host-side planner that turns a
// pmb_problem into the shared-memory carve-up, the weight-stream schedule and the workspace layout
// of the sweeps.  No allocation, no host synchronisation: everything is enqueued on the caller's stream.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "pmb_host.h"
#include "pmb_internal.cuh"
#include "pmb_mm.cuh"
#include "pmb_cluster.cuh"
#include "pmb_cluster_mm.cuh"
#include "pmb_cw.cuh"
#include "pmb_tc.cuh"
#include "pmb_tc_mm.cuh"

namespace pmb {

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static inline int round4(int x) { return (x + 3) & ~3; }
static inline long long round32(long long x) { return (x + 31) & ~31LL; }

constexpr int SMEM_LIMIT_FLOATS = (232448 - 2048) / 4;   // 227 KB opt-in shared memory per CTA minus the kernels' static part
                                                        // (ring barriers, chunk table)

struct WGrad {
    long long delta_off;   // workspace floats
    int lda, M;
    long long inp_off;     // workspace floats, -1 = the states trajectory
    int ldb, Nc;
    long long w_off, b_off;   // offsets inside the flat gradient (b_off = -1: no bias)
};

struct Plan {
    int P, nsplit, stream_mode;
    SweepParams fwd, bwd;
    PackJobs jobs;            // dst pointers are workspace float OFFSETS until resolved
    long long job_dst_off[MAX_PACK_JOBS];
    long long wpack_fwd_off, wpack_bwd_off;
    long long part_off;
    long long s1pre_off, mmstat_off, mmrec_off, mmctr_off, rpre_off, rstat_off, geff_off, cmm_gbuf_off;
    // moment matching across GPUs: full-width [H][n_global] copies of the pre-matching rewards, the matched rewards,
    // their cotangents and the cotangents of the pre-matching rewards
    int mm_world;
    long long rfull_off, routfull_off, gfull_off, gefffull_off;
    int mm_G;
    long long nparam;
    long long ws_floats;
    int smem_fwd_bytes, smem_bwd_bytes;
    int n_wg;
    WGrad wg[MAXL];
    // cluster-resident sweeps (pmb_cluster.cuh): used instead of the streaming sweeps when eligible
    int cluster;              // 0 = streaming sweeps, otherwise CTAs per cluster
    int cl_nclusters;
    long long cl_pre_off;     // [H][N][2D + 3U] step-local adjoint factors of the cluster-resident reverse sweep
    ClusterParams cfwd, cbwd;
    // tensor-core cluster sweeps (pmb_tc.cuh): used instead of both when chosen
    int tc;                   // 0 = off, otherwise CTAs per cluster
    TcParams tfwd, tbwd;
    TcPackJobs tjobs;         // hi/lo weight slices of both sweep directions
    PackJobs sjobs;           // the workspace-resident small operands only (padded biases, masks)
    long long tc_wpack_off, tc_xbuf_off, tc_opart_off;
    ClusterParams tpre;       // arguments of the adjoint-factor pre-pass (cluster_bwd_pre_kernel)
    // wide cluster-resident sweeps (pmb_cw.cuh): 16-CTA clusters, two-hidden-layer nets up to 512 wide
    int cw;                   // 0 = off
    int cw_nclusters;
    long long cw_g1_off, cw_g2_off;
    ClusterParams wfwd, wbwd;
};

struct Alloc {
    long long top = 0;
    long long take(long long n) {
        long long o = top;
        top += round32(n);
        return o;
    }
};

static int check_net(const pmb_net &n, int nin, int nout, const char *what) {
    if (n.n_linear < 1 || n.n_linear > MAXL) return fail(PMB_E_UNSUPPORTED, "%s: n_linear=%d outside [1,%d]", what, n.n_linear, MAXL);
    if (n.dims[0] != nin) return fail(PMB_E_INVALID, "%s: dims[0]=%d, expected %d", what, n.dims[0], nin);
    if (n.dims[n.n_linear] != nout) return fail(PMB_E_INVALID, "%s: output size %d, expected %d", what, n.dims[n.n_linear], nout);
    for (int l = 0; l < n.n_linear; ++l) {
        if (!n.W[l]) return fail(PMB_E_INVALID, "%s: W[%d] is NULL", what, l);
        if (n.dims[l + 1] < 1) return fail(PMB_E_INVALID, "%s: dims[%d] < 1", what, l + 1);
        if (l + 1 < n.n_linear && n.dims[l + 1] > PMB_MAX_WIDTH)
            return fail(PMB_E_UNSUPPORTED, "%s: hidden width %d > %d", what, n.dims[l + 1], PMB_MAX_WIDTH);
        if (l + 1 < n.n_linear && !(n.keep[l] > 0.f)) return fail(PMB_E_INVALID, "%s: keep[%d] must be > 0", what, l);
    }
    if (n.has_density && !n.z) return fail(PMB_E_INVALID, "%s: density noise z is NULL", what);
    return PMB_OK;
}

// Fill the fwd/bwd layer tables of one net, its pack jobs and workspace regions.
static void plan_net(const pmb_net &net, int N, int H, bool is_policy, NetSweep &F, NetSweep &B, Plan &pl,
                     Alloc &wf, Alloc &wb, Alloc &ws, long long region_base[2]) {
    (void)region_base;
    const int L = net.n_linear - 1;
    int npad[MAXL];
    for (int l = 0; l < L; ++l) npad[l] = round4(net.dims[l + 1]);
    memset(&F, 0, sizeof(F));
    memset(&B, 0, sizeof(B));
    F.nlin = B.nlin = net.n_linear;
    F.nout = B.nout = net.dims[net.n_linear];
    F.nin = B.nin = net.dims[0];
    F.has_density = B.has_density = net.has_density;
    F.lmax = B.lmax = net.max_log_std;
    F.z = B.z = net.z;
    F.zstride = B.zstride = net.z_step_stride;
    auto add_job = [&](const float *src, long long dst_off, int R, int C, int SR, int SC, int ld, int tr, int area) {
        PackJob &j = pl.jobs.job[pl.jobs.n];
        j.src = src; j.dst = nullptr; j.R = R; j.C = C; j.SR = SR; j.SC = SC; j.src_ld = ld; j.transpose = tr;
        pl.job_dst_off[pl.jobs.n] = dst_off | ((long long)area << 60);
        ++pl.jobs.n;
    };
    for (int l = 0; l <= L; ++l) {
        Lin &f = F.lin[l];
        Lin &b = B.lin[l];
        const int out_l = net.dims[l + 1], in_l = net.dims[l];
        // ---------------- forward ----------------
        if (l < L) {
            f.kind = 0; f.K = (l == 0) ? in_l : npad[l - 1]; f.Nout = out_l; f.Npad = npad[l]; f.streamed = (l >= 1);
            f.goff = wf.take((long long)f.K * f.Npad);
            add_job(net.W[l], f.goff, f.K, f.Npad, out_l, in_l, in_l, 1, 0);
        } else {
            f.kind = 1; f.K = (L == 0) ? in_l : npad[L - 1]; f.Nout = out_l; f.Npad = out_l; f.streamed = 0;
            f.goff = wf.take((long long)f.Nout * f.K);
            add_job(net.W[l], f.goff, f.Nout, f.K, out_l, in_l, in_l, 0, 0);
        }
        f.boff = -1;
        if (net.b[l]) {
            f.boff = ws.take(f.Npad);
            add_job(net.b[l], f.boff, 1, f.Npad, 1, out_l, out_l, 0, 2);
        }
        // ---------------- backward ----------------
        if (l >= 1) {
            b.kind = 0; b.K = (l == L) ? out_l : npad[l]; b.Nout = in_l; b.Npad = npad[l - 1]; b.streamed = (l < L);
            b.goff = wb.take((long long)b.K * b.Npad);
            add_job(net.W[l], b.goff, b.K, b.Npad, out_l, in_l, in_l, 0, 1);
        } else {
            b.kind = 1; b.K = (L == 0) ? out_l : npad[0]; b.Nout = in_l; b.Npad = in_l; b.streamed = 0;
            b.goff = wb.take((long long)b.Nout * b.K);
            add_job(net.W[l], b.goff, b.Nout, b.K, out_l, in_l, in_l, 1, 1);
        }
        b.boff = -1;
    }
    for (int l = 0; l < L; ++l) {
        F.keep[l] = B.keep[l] = net.keep[l];
        F.mask_off[l] = B.mask_off[l] = -1;
        if (net.mask[l]) {
            long long o = ws.take((long long)N * npad[l]);
            F.mask_off[l] = B.mask_off[l] = o;
            add_job(net.mask[l], o, N, npad[l], N, net.dims[l + 1], net.dims[l + 1], 0, 2);
        }
        long long o = ws.take((long long)H * N * npad[l]);
        F.saved_off[l] = B.saved_off[l] = o;
        if (is_policy) {
            long long d = ws.take((long long)H * N * npad[l]);
            F.delta_off[l] = B.delta_off[l] = d;
        }
    }
    {
        long long o = ws.take((long long)H * N * F.nout);
        F.outsaved_off = B.outsaved_off = o;
        if (is_policy) {
            long long d = ws.take((long long)H * N * F.nout);
            F.delta_off[L] = B.delta_off[L] = d;
        }
    }
}

// shared-memory carve-up + stream schedule of one sweep
static int plan_sweep(SweepParams &S, const NetSweep *order[2], bool reverse, int P, int stream_mode,
                      int nstages_req, bool mm_states) {
    int off = 0;
    S.nres = 0;
    S.nsched = 0;
    S.off_cst = off; off += (C_TOTAL + 31) & ~31;
    int tile_rows = 4;
    int max_npad = 0;
    auto add_res = [&](long long goff, int soff, int n, int kind) {
        S.res_goff[S.nres] = goff; S.res_soff[S.nres] = soff; S.res_n[S.nres] = n; S.res_ws[S.nres] = kind;
        ++S.nres;
    };
    // resident weights + biases
    for (int n = 0; n < 2; ++n) {
        NetSweep &net = const_cast<NetSweep &>(*order[n]);
        for (int l = 0; l < net.nlin; ++l) {
            Lin &L = net.lin[l];
            tile_rows = max(tile_rows, max(L.K, L.Npad));
            if (L.streamed) max_npad = max(max_npad, L.Npad);
            if (!L.streamed) {
                int nfl = (L.kind == 0) ? L.K * L.Npad : L.Nout * L.K;
                nfl = (nfl + 3) & ~3;
                L.soff = off;
                add_res(L.goff, off, nfl, 0);
                off += (nfl + 31) & ~31;
            }
            L.bias_soff = -1;
            if (!reverse && L.boff >= 0) {
                int nfl = (L.Npad + 3) & ~3;
                L.bias_soff = off;
                add_res(L.boff, off, nfl, 1);
                off += (nfl + 31) & ~31;
            }
        }
    }
    S.off_act0 = off; off += ((tile_rows * P) + 31) & ~31;
    S.off_act1 = off; off += ((tile_rows * P) + 31) & ~31;
    S.off_red = off;  off += 1024 * P;
    S.off_misc = off; off += 640 * P;   // per-particle scratch + partials of the fused projections
    S.off_mm = off;
    if (mm_states) off += (mm_smem_floats(P) + 31) & ~31;
    // backward: two buffers for the stored hidden activations of a step
    S.off_sav = off;
    S.sav_floats = 0;
    if (reverse) {
        int sf = 0;
        for (int n = 0; n < 2; ++n) {
            NetSweep &net = const_cast<NetSweep &>(*order[n]);
            for (int h = 0; h + 1 < net.nlin; ++h) {
                net.sav_soff[h] = sf;
                sf += P * net.lin[h + 1].Npad;      // bwd lin[h+1].Npad = padded width of hidden h
            }
        }
        S.sav_floats = (sf + 31) & ~31;
        off += 2 * S.sav_floats;
    }
    // dropout-mask rows of this CTA: resident when they leave room for a useful ring
    int mask_floats = 0;
    for (int n = 0; n < 2; ++n) {
        const NetSweep &net = *order[n];
        for (int h = 0; h + 1 < net.nlin; ++h) {
            const int npad = reverse ? net.lin[h + 1].Npad : net.lin[h].Npad;
            if (net.mask_off[h] >= 0) mask_floats += ((P * npad) + 31) & ~31;
        }
    }
    const int ring_min = max_npad ? 2 * 16 * max_npad : 0;
    const bool masks_resident = off + mask_floats + ring_min <= SMEM_LIMIT_FLOATS;
    for (int n = 0; n < 2; ++n) {
        NetSweep &net = const_cast<NetSweep &>(*order[n]);
        for (int h = 0; h + 1 < net.nlin; ++h) {
            net.mask_soff[h] = -1;
            if (net.mask_off[h] >= 0 && masks_resident) {
                const int npad = reverse ? net.lin[h + 1].Npad : net.lin[h].Npad;
                net.mask_soff[h] = off;
                add_res(net.mask_off[h], off, P * npad, npad);
                off += ((P * npad) + 31) & ~31;
            }
        }
    }
    S.off_stage = off;
    S.stream_mode = stream_mode;
    if (max_npad == 0) {
        S.nstages = 1; S.stage_floats = 0; S.chunks_per_step = 0;
    } else {
        const int avail = SMEM_LIMIT_FLOATS - off;
        int ns = nstages_req >= 2 && nstages_req <= MAXS ? nstages_req : 2;
        int sf = (avail / ns) & ~31;
        while (ns > 2 && sf < 8 * max_npad) { --ns; sf = (avail / ns) & ~31; }
        if (sf < max_npad) return fail(PMB_E_UNSUPPORTED, "network too wide for the shared-memory ring");
        S.nstages = ns;
        int cps = 0, sf_used = 0;
        for (int n = 0; n < 2; ++n) {
            NetSweep &net = const_cast<NetSweep &>(*order[n]);
            for (int i = 0; i < net.nlin; ++i) {
                int l = reverse ? net.nlin - 1 - i : i;
                Lin &L = net.lin[l];
                if (!L.streamed) continue;
                const int kcmax = sf / L.Npad;
                L.nchunks = (L.K + kcmax - 1) / kcmax;
                L.kc = (L.K + L.nchunks - 1) / L.nchunks;       // balanced chunks
                sf_used = max(sf_used, L.kc * L.Npad);
                StreamItem &it = S.sched[S.nsched++];
                it.goff = L.goff; it.kc = L.kc; it.nchunks = L.nchunks; it.K = L.K; it.Npad = L.Npad;
                cps += L.nchunks;
            }
        }
        S.stage_floats = (sf_used + 31) & ~31;
        S.chunks_per_step = cps;
        if (cps > MAXCHUNKS) return fail(PMB_E_UNSUPPORTED, "weight stream needs %d chunks per step (> %d)", cps, MAXCHUNKS);
        off += ns * S.stage_floats;
    }
    if (off > SMEM_LIMIT_FLOATS) return fail(PMB_E_UNSUPPORTED, "shared-memory plan needs %d bytes > 227 KB", off * 4);
    return off * 4;
}

// ---------------------------------------------------------------------------------------------
// Cluster-resident sweeps: eligibility, operand tables and shared-memory carve-up (pmb_cluster.cuh).
// Eligible: two hidden layers per net (models.mlp as used by every example of the reference), hidden
// widths <= 256, D+U <= 16, outputs <= 16, no moment matching of the states.
// ---------------------------------------------------------------------------------------------
static bool cluster_eligible(const pmb_problem *p, int C) {
    // moment matching of the states: one matching group (the whole particle set) of <= 128 particles
    if (p->mm_states && (p->mm_groups > 1 || p->N > CMM_NMAX || (p->mm_world > 1 ? p->n_global : p->N) < 2)) return false;
    const pmb_net *nets[2] = {&p->pol, &p->dyn};
    for (int i = 0; i < 2; ++i) {
        const pmb_net &n = *nets[i];
        if (n.n_linear != 3) return false;
        const int w0 = round4(n.dims[1]), w1 = round4(n.dims[2]);
        if (w0 > CL_TW || w1 > CL_TW) return false;
        if (n.dims[0] > CL_NO || n.dims[3] > CL_NO) return false;
        if (round4((w0 + C - 1) / C) > CL_HS || round4((w1 + C - 1) / C) > CL_HS) return false;
    }
    return true;
}

static void cluster_net(const NetSweep &S, bool reverse, bool is_policy, int C, CNet &n) {
    memset(&n, 0, sizeof(n));
    const Lin &thin = reverse ? S.lin[2] : S.lin[0];
    const Lin &wide = S.lin[1];
    const Lin &narrow = reverse ? S.lin[0] : S.lin[2];
    const int ht = reverse ? 1 : 0, hw = reverse ? 0 : 1;     // hidden layer produced by the thin / wide op
    n.tK = thin.K;
    n.tW = thin.Npad;
    n.tsl = round4((n.tW + C - 1) / C);
    n.wN = wide.Npad;
    n.hs = round4((n.wN + C - 1) / C);
    n.nN = narrow.Nout;
    n.nNp = round4(n.nN);
    n.t_goff = thin.goff;
    n.w_goff = wide.goff;
    n.n_goff = narrow.goff;
    n.tb_off = reverse ? -1 : thin.boff;
    n.wb_off = reverse ? -1 : wide.boff;
    n.nb_off = reverse ? -1 : narrow.boff;
    n.tm_off = S.mask_off[ht];
    n.wm_off = S.mask_off[hw];
    n.tkeep_inv = 1.f / S.keep[ht];
    n.wkeep_inv = 1.f / S.keep[hw];
    n.tsav_off = S.saved_off[ht];
    n.wsav_off = S.saved_off[hw];
    n.tdel_off = (reverse && is_policy) ? S.delta_off[ht] : -1;
    n.wdel_off = (reverse && is_policy) ? S.delta_off[hw] : -1;
    n.odel_off = (reverse && is_policy) ? S.delta_off[2] : -1;
    n.raw_off = S.outsaved_off;
    n.nraw = S.nout;
    n.has_density = S.has_density;
    n.lmax = S.lmax;
    n.z = S.z;
    n.zstride = S.zstride;
}

static int cluster_carve(ClusterParams &P, bool reverse) {
    int off = 0;
    auto take = [&](int nfl) { int o = off; off += (nfl + 31) & ~31; return o; };
    P.off_cst = take(C_TOTAL);
    CNet *nets[2] = {&P.pol, &P.dyn};
    int tw_max = 4;
    for (int i = 0; i < 2; ++i) {
        CNet &n = *nets[i];
        n.s_tw = take(CL_NO * n.tW);        // rows >= tK stay zero (the thin layer runs whole 4-row blocks)
        n.s_ww = take(n.tW * n.hs);
        n.s_nwt = take(CL_NO * CL_HS);
        n.s_tb = take(n.tW);
        n.s_wb = take(n.hs);
        n.s_nb = take(CL_NO);
        n.s_tm = take(CL_PS * n.tW);
        n.s_wm = take(CL_PS * n.hs);
        tw_max = max(tw_max, n.tW);
    }
    P.off_xa = take(2 * CL_NO * CL_PS);
    P.off_xb = take(2 * CL_NO * CL_PS);
    P.off_act = take(tw_max * CL_PS);
    P.off_red = take(CL_KS * CL_PS * 32);
    P.off_inbox = take(2 * 2 * P.C * CL_MBOX);        // [tile][exchange][sender rank][4 slots x 16]
    P.off_misc = take(reverse ? (4 + 2 * 6) * CL_PS * SD + 32 : CL_PS * SD);
    P.off_mm = off;
    if (P.mm_states) take(2 * CMM_FLOATS);
    P.smem_floats = off;
    return off;
}

// fills pl.cluster / pl.cfwd / pl.cbwd from the streaming plan's layer tables (same workspace layout)
static int plan_cluster(const pmb_problem *p, const pmb_tuning *tune, Plan &pl) {
    pl.cluster = 0;
    const int mode = tune ? tune->stream_mode : 0;
    if (mode == 1 || mode == 2) return PMB_OK;
    int C = (tune && mode == 3) ? (tune->reserved[1] >> 4) & 15 : 0;
    if (C == 0) C = 8;
    if (C != 4 && C != 8) return fail(PMB_E_INVALID, "cluster size must be 4 or 8");
    if (!cluster_eligible(p, C)) {
        if (mode == 3) return fail(PMB_E_UNSUPPORTED, "problem is outside the cluster-resident sweeps");
        return PMB_OK;
    }
    for (int pass = 0; pass < 2; ++pass) {
        const SweepParams &S = pass ? pl.bwd : pl.fwd;
        ClusterParams &P = pass ? pl.cbwd : pl.cfwd;
        memset(&P, 0, sizeof(P));
        P.N = S.N; P.H = S.H; P.D = S.D; P.U = S.U; P.C = C;
        P.mm_states = p->mm_states; P.z_mm = p->z_mm;
        P.mm_world = p->mm_world > 1 ? p->mm_world : 1;
        P.mm_rank = p->mm_world > 1 ? p->mm_rank : 0;
        P.n_global = p->mm_world > 1 ? p->n_global : p->N;
        P.n_off = P.mm_rank * p->N;
        cluster_net(S.pol, pass == 1, true, C, P.pol);
        cluster_net(S.dyn, pass == 1, false, C, P.dyn);
        P.act_scale = S.act_scale; P.act_bias = S.act_bias; P.mx = S.mx; P.iSx = S.iSx; P.my = S.my; P.Sy = S.Sy;
        P.KR = S.KR; P.rew_C = S.rew_C; P.rew_c0 = S.rew_c0; P.rew_Q = S.rew_Q; P.rew_R = S.rew_R;
        P.rew_scale = S.rew_scale; P.rew_offset = S.rew_offset;
        if (cluster_carve(P, pass == 1) > SMEM_LIMIT_FLOATS) {
            if (mode == 3) return fail(PMB_E_UNSUPPORTED, "cluster-resident plan does not fit in shared memory");
            return PMB_OK;
        }
    }
    // particles per cluster: spread the particles over every cluster the device can hold at once
    int PG = (tune && mode == 3) ? tune->reserved[1] & 15 : 0;
    if (PG < 0 || PG > CL_PS) return fail(PMB_E_INVALID, "particles per cluster must be 1..%d", CL_PS);
    if (PG == 0) {
        // co-resident clusters of this kernel, queried once per (device, cluster size)
        static int cached[32][2];
        static bool cached_init = false;
        if (!cached_init) { for (auto &c : cached) c[0] = c[1] = -1; cached_init = true; }
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) dev = 0;
        int &mc = cached[dev][C == 8];
        if (mc < 0) mc = cluster_max_active(C, max(pl.cfwd.smem_floats, pl.cbwd.smem_floats) * 4, true);
        const int maxc = mc > 0 ? mc : (C == 8 ? 16 : 32);
        PG = (p->N + maxc - 1) / maxc;
        if (PG > CL_PS) PG = CL_PS;
        if (PG < 1) PG = 1;
    }
    if (p->mm_states) {
        // every cluster must be resident (the per-step exchange is a grid-wide barrier): spread over at most the
        // co-resident clusters
        static int mmc[32];
        static bool mmc_init = false;
        if (!mmc_init) { for (auto &c : mmc) c = -1; mmc_init = true; }
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) dev = 0;
        if (mmc[dev] < 0) mmc[dev] = cluster_max_active(C, max(pl.cfwd.smem_floats, pl.cbwd.smem_floats) * 4, true);
        const int maxc = mmc[dev] > 0 ? mmc[dev] : (C == 8 ? 15 : 30);
        if ((p->N + PG - 1) / PG > maxc) PG = (p->N + maxc - 1) / maxc;
        if (PG > CL_PS) {
            if (mode == 3) return fail(PMB_E_UNSUPPORTED, "mm_states with N=%d particles does not fit the co-resident clusters", p->N);
            return PMB_OK;
        }
    }
    pl.cfwd.PG = pl.cbwd.PG = PG;
    {
        const int st = (tune && mode == 3) ? (tune->reserved[1] >> 8) & 0xffff : 0;
        pl.cfwd.pingpong = pl.cbwd.pingpong = st ? (st - 1) : 1;
        // with moment matching the forward tiles wait for each other's records every step: letting them run free
        // (instead of alternating on the shared-memory-bound phases) brings the last record 4 % earlier (c3 forward
        // sweep 3.06 -> 2.95 ms); the reverse sweep does not care
        if (!st && p->mm_states) pl.cfwd.pingpong = 0;
    }
    pl.cl_nclusters = (p->N + PG - 1) / PG;
    pl.cluster = C;
    return PMB_OK;
}

// ---------------------------------------------------------------------------------------------
// Wide cluster-resident sweeps (pmb_cw.cuh).  Eligible: two hidden layers per net, hidden widths <= 512, D+U <= 16,
// outputs <= 16, binary dropout masks (pmb_problem.masks_binary), no moment matching of the states.
// ---------------------------------------------------------------------------------------------
static bool cw_eligible(const pmb_problem *p) {
    if (p->mm_states || !p->masks_binary) return false;
    const pmb_net *nets[2] = {&p->pol, &p->dyn};
    for (int i = 0; i < 2; ++i) {
        const pmb_net &n = *nets[i];
        if (n.n_linear != 3) return false;
        if (round4(n.dims[1]) > CW_TW || round4(n.dims[2]) > CW_TW) return false;
        if (n.dims[0] > CW_NO || n.dims[3] > CW_NO) return false;
    }
    return true;
}

static int cw_carve(ClusterParams &P, bool reverse) {
    int off = 0;
    auto take = [&](int nfl) { int o = off; off += (nfl + 31) & ~31; return o; };
    P.off_cst = reverse ? 0 : take(C_TOTAL);
    CNet *nets[2] = {&P.pol, &P.dyn};
    int tw_max = 4;
    for (int i = 0; i < 2; ++i) {
        CNet &n = *nets[i];
        n.s_ww = take(n.tW * CW_HS);
        n.s_nwt = take(CW_NO * CW_HS);
        n.s_wb = take(CW_HS);
        n.s_nb = take(CW_NO);
        tw_max = max(tw_max, n.tW);
    }
    P.off_xa = take(CW_XT);
    P.off_xb = take(CW_XT);
    P.off_act = take(max(tw_max, CW_NW * 32) * CW_PS);    // hidden tile [tW][36]; partial sums [16][36][32]
    P.off_inbox = take(2 * CW_MB);
    P.off_misc = take(reverse ? 2 * CW_G2 : 32);
    P.smem_floats = off;
    return off;
}

// fills pl.cw / pl.wfwd / pl.wbwd from the streaming plan's layer tables (same workspace layout + the gate words)
static int plan_cw(const pmb_problem *p, const pmb_tuning *tune, Plan &pl, Alloc &ws) {
    pl.cw = 0;
    const int mode = tune ? tune->stream_mode : 0;
    if (mode != 0 && mode != 5) return PMB_OK;
    if (mode == 0 && cluster_eligible(p, 8)) return PMB_OK;      // narrower nets: the 8-CTA latency-oriented sweeps
    if (!cw_eligible(p)) {
        if (mode == 5) return fail(PMB_E_UNSUPPORTED, "problem is outside the wide cluster-resident sweeps%s",
                                   p->masks_binary ? "" : " (masks_binary is not set)");
        return PMB_OK;
    }
    for (int pass = 0; pass < 2; ++pass) {
        const SweepParams &S = pass ? pl.bwd : pl.fwd;
        ClusterParams &P = pass ? pl.wbwd : pl.wfwd;
        memset(&P, 0, sizeof(P));
        P.N = p->N; P.H = p->H; P.D = p->D; P.U = p->U; P.C = CW_C;
        cluster_net(S.pol, pass == 1, true, CW_C, P.pol);
        cluster_net(S.dyn, pass == 1, false, CW_C, P.dyn);
        P.act_scale = p->act_scale; P.act_bias = p->act_bias; P.mx = p->mx; P.iSx = p->iSx; P.my = p->my; P.Sy = p->Sy;
        P.KR = p->rew_rows; P.rew_C = p->rew_C; P.rew_c0 = p->rew_c0; P.rew_Q = p->rew_Q; P.rew_R = p->rew_R;
        P.rew_scale = p->rew_scale; P.rew_offset = p->rew_offset;
        if (P.pol.hs > CW_HS || P.dyn.hs > CW_HS || cw_carve(P, pass == 1) > SMEM_LIMIT_FLOATS) {
            if (mode == 5) return fail(PMB_E_UNSUPPORTED, "wide cluster-resident plan does not fit in shared memory");
            return PMB_OK;
        }
    }
    // particles per cluster: spread the particles over every cluster the device can hold at once
    int PG = (tune && mode == 5) ? tune->reserved[1] & 63 : 0;
    if (PG < 0 || PG > CW_PS) return fail(PMB_E_INVALID, "particles per cluster must be 1..%d", CW_PS);
    if (PG == 0) {
        static int cached[32];
        static bool cached_init = false;
        if (!cached_init) { for (auto &c : cached) c = -1; cached_init = true; }
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) dev = 0;
        int &mc = cached[dev];
        if (mc < 0) mc = cw_max_active();
        const int maxc = mc > 0 ? mc : 7;
        PG = (p->N + maxc - 1) / maxc;
        if (PG > CW_PS) PG = CW_PS;
        if (PG < 1) PG = 1;
    }
    pl.wfwd.PG = pl.wbwd.PG = PG;
    pl.cw_nclusters = (p->N + PG - 1) / PG;
    pl.wfwd.ncl = pl.wbwd.ncl = pl.cw_nclusters;
    pl.cw_g1_off = ws.take((long long)p->H * pl.cw_nclusters * 2 * 2 * CW_TW);
    pl.cw_g2_off = ws.take((long long)p->H * p->N * 2 * CW_C);
    pl.cw = CW_C;
    return PMB_OK;
}

// ---------------------------------------------------------------------------------------------
// Tensor-core cluster sweeps: eligibility, operand tables, shared-memory carve-up, workspace (pmb_tc.cuh).
// Eligible: >= 1 hidden layer per net, hidden widths <= 1024, <= 32 raw outputs, D+U <= 16; with moment matching
// of the states the whole particle set must fit one tile (N <= 128, <= 16 groups).
// ---------------------------------------------------------------------------------------------
static int tc_slice_width(int maxw, int C) {
    const int per = (maxw + C - 1) / C;
    return per <= 16 ? 16 : per <= 32 ? 32 : per <= 64 ? 64 : 0;
}

static bool tc_eligible(const pmb_problem *p, int C) {
    const pmb_net *nets[2] = {&p->pol, &p->dyn};
    int maxw = 0;
    for (int i = 0; i < 2; ++i) {
        const pmb_net &n = *nets[i];
        if (n.n_linear < 2) return false;
        if (n.dims[0] > 16 || n.dims[n.n_linear] > TC_NOUT) return false;
        for (int l = 1; l < n.n_linear; ++l) maxw = max(maxw, n.dims[l]);
    }
    if (tc_slice_width(maxw, C) == 0) return false;
    if (p->mm_states) {
        const int G = p->mm_groups > 1 ? p->mm_groups : 1;
        if (p->N > TC_M || G > TCMM_MAXG) return false;
    }
    return true;
}

static void tc_net(const pmb_net &net, const NetSweep &F, bool reverse, bool is_policy, TcNet &n) {
    memset(&n, 0, sizeof(n));
    n.L = net.n_linear - 1;
    n.nin = net.dims[0];
    n.nout = net.dims[net.n_linear];
    for (int l = 0; l < n.L; ++l) {
        n.width[l] = net.dims[l + 1];
        n.npad[l] = round4(net.dims[l + 1]);
        n.kb[l] = (net.dims[l + 1] + 7) / 8;
        n.mask_off[l] = F.mask_off[l];
        n.keep_inv[l] = 1.f / F.keep[l];
        n.saved_off[l] = F.saved_off[l];
    }
    for (int l = 0; l <= n.L; ++l) {
        n.bias_off[l] = reverse ? -1 : F.lin[l].boff;
        n.delta_off[l] = is_policy ? F.delta_off[l] : -1;
    }
    n.W_first = net.W[0];
    n.W_last = net.W[n.L];
    n.raw_off = F.outsaved_off;
    n.has_density = net.has_density;
    n.lmax = net.max_log_std;
    n.z = net.z;
    n.zstride = net.z_step_stride;
}

// fills pl.tc / pl.tfwd / pl.tbwd; `ws` allocates the extra workspace regions
static int plan_tc(const pmb_problem *p, const pmb_tuning *tune, Plan &pl, Alloc &ws) {
    pl.tc = 0;
    const int mode = tune ? tune->stream_mode : 0;
    if (mode == 1 || mode == 2 || mode == 3) return PMB_OK;
    const int C = 16;
    if (!tc_eligible(p, C)) {
        if (mode == 4) return fail(PMB_E_UNSUPPORTED, "problem is outside the tensor-core cluster sweeps");
        return PMB_OK;
    }
    // Opt-in only (stream_mode 4).  Measured on B200 (profiles/r02_tc_timelines_v1_v2.txt): a tcgen05.mma costs ~100
    // cycles to issue whatever its size, a column slice of a 128-particle tile is only N = 16..64 wide and K = 8 per
    // tf32 instruction, so one 512x512 layer is 192 instructions = ~25 k cycles per CTA -- slower than the FFMA2
    // streaming sweeps at every BASELINE shape.  The auto planner therefore never picks this variant.
    if (mode != 4) return PMB_OK;
    int maxw = 0, nop = 1;
    const pmb_net *nets[2] = {&p->pol, &p->dyn};
    for (int i = 0; i < 2; ++i) {
        for (int l = 1; l < nets[i]->n_linear; ++l) maxw = max(maxw, nets[i]->dims[l]);
        nop = max(nop, max(nets[i]->dims[0], nets[i]->dims[nets[i]->n_linear]));
    }
    const int ns = tc_slice_width(maxw, C);
    const int ntiles = (p->N + TC_M - 1) / TC_M;
    const int TP = (p->N + ntiles - 1) / ntiles;
    // ---- hi/lo weight slices of every hidden x hidden layer, both directions ----
    Alloc wa;
    pl.tjobs.n = 0;
    pl.tjobs.C = C;
    pl.tjobs.ns = ns;
    for (int pass = 0; pass < 2; ++pass) {
        TcParams &T = pass ? pl.tbwd : pl.tfwd;
        memset(&T, 0, sizeof(T));
        T.N = p->N; T.H = p->H; T.D = p->D; T.U = p->U; T.C = C; T.TP = TP; T.ntiles = ntiles; T.ns = ns;
        T.kbmax = C * ns / 8;
        T.nop = nop;
        tc_net(p->pol, pl.fwd.pol, pass == 1, true, T.pol);
        tc_net(p->dyn, pl.fwd.dyn, pass == 1, false, T.dyn);
        for (int i = 0; i < 2; ++i) {
            TcNet &n = i ? T.dyn : T.pol;
            const pmb_net &src = *nets[i];
            for (int l = 1; l < n.L; ++l) {
                // forward: K = width of hidden l-1; reverse: K = width of hidden l (B[n][k] = W_l[k][n])
                const int kb = pass ? n.kb[l] : n.kb[l - 1];
                n.wp_off[l] = wa.take((long long)C * 2 * kb * ns * 8);
                TcPackJob &j = pl.tjobs.job[pl.tjobs.n++];
                j.W = src.W[l]; j.out = src.dims[l + 1]; j.in = src.dims[l]; j.transpose = pass; j.dst_off = n.wp_off[l]; j.kb = kb;
            }
        }
        T.act_scale = p->act_scale; T.act_bias = p->act_bias; T.mx = p->mx; T.iSx = p->iSx; T.my = p->my; T.Sy = p->Sy;
        T.KR = p->rew_rows; T.rew_C = p->rew_C; T.rew_c0 = p->rew_c0; T.rew_Q = p->rew_Q; T.rew_R = p->rew_R;
        T.rew_scale = p->rew_scale; T.rew_offset = p->rew_offset;
        T.mm_states = p->mm_states; T.mm_G = pl.mm_G; T.mm_Ng = p->N / pl.mm_G; T.z_mm = p->z_mm;
        // ---- shared memory ----
        int off = 0;
        auto take = [&](int nfl) { int o = off; off += (nfl + 31) & ~31; return o; };
        T.off_cst = take(C_TOTAL);
        T.off_res = off;
        for (int i = 0; i < 2; ++i) {
            TcNet &n = i ? T.dyn : T.pol;
            n.s_wfirst = take(16 * ns);
            n.s_wlast = take(TC_NOUT * ns);
            n.s_bias = pass ? 0 : take((MAXL + 1) * TC_MAXNS);
        }
        const int PW = 2 * p->D + 3 * p->U;
        T.off_xin = take(pass ? TC_M * (TC_NOUT + 1) : TC_M * TC_SDP);      // reverse: adjoint of the raw outputs
        T.off_st = take((pass ? 2 : 1) * TC_M * TC_SDP);
        T.off_aux = take(2 * nop * TC_M);
        T.off_z = take(pass ? TC_M * PW : 2 * TC_M * TC_SDP);             // forward: density noise; reverse: adjoint factors
        T.off_mm = p->mm_states ? take(TCMM_FLOATS) : off;
        T.off_ring = off;
        const int budget = SMEM_LIMIT_FLOATS - 1280 - off;         // 5 KB of slack for the static barriers + schedule table
        const int per_stage = 2 * TC_KC * ns * 8;                  // weights: hi | lo of TC_KC k-blocks
        const int per_x = TC_KC * 1024;                            // image: fp32 [TC_KC][2][128][4]
        int nsx = TC_NSX, nstage = TC_NSW;
        while (nsx * per_x + nstage * per_stage > budget && (nstage > 2 || nsx > 2)) {
            if (nstage > 2 && (nstage >= nsx + 1 || nsx <= 2)) --nstage; else --nsx;
        }
        {
            const int e = tune ? tune->reserved[1] : 0;            // tuning aid: bits 0-3 image stages, 8-11 weight stages
            if (mode == 4 && ((e >> 8) & 15) >= 1 && ((e >> 8) & 15) <= TC_NSW) nstage = min(nstage, (e >> 8) & 15);
            if (mode == 4 && (e & 15) >= 1 && (e & 15) <= TC_NSX) nsx = min(nsx, e & 15);
        }
        if (nsx * per_x + nstage * per_stage > budget)
            return mode == 4 ? fail(PMB_E_UNSUPPORTED, "tensor-core plan does not fit in shared memory") : PMB_OK;
        T.kb_stage = TC_KC;
        T.nstage = nstage;
        T.nsx = nsx;
        T.nsa = min(TC_NSA, (512 - TC_NDRV * ns) / 128);
        T.dbg_flags = (tune && mode == 4) ? (tune->reserved[1] >> 12) & 15 : 0;
        T.stage_floats = per_stage;
        off += nstage * per_stage;
        T.off_xring = off;
        off += nsx * per_x;
        T.smem_floats = off;
    }
    pl.tc_wpack_off = ws.take(wa.top);
    pl.tc_xbuf_off = ws.take((long long)ntiles * 2 * (C * ns / 8) * 1024);
    pl.tc_opart_off = ws.take((long long)ntiles * 2 * C * nop * TC_M);
    // the workspace-resident small operands (padded biases, masks): the streaming weight images are not needed
    pl.sjobs.n = 0;
    for (int i = 0; i < pl.jobs.n; ++i)
        if ((pl.job_dst_off[i] >> 60) == 2) pl.sjobs.job[pl.sjobs.n++] = pl.jobs.job[i];
    // adjoint-factor pre-pass
    ClusterParams &P = pl.tpre;
    memset(&P, 0, sizeof(P));
    P.N = p->N; P.H = p->H; P.D = p->D; P.U = p->U;
    P.pol.raw_off = pl.fwd.pol.outsaved_off; P.pol.nraw = pl.fwd.pol.nout; P.pol.has_density = p->pol.has_density;
    P.pol.lmax = p->pol.max_log_std; P.pol.z = p->pol.z; P.pol.zstride = p->pol.z_step_stride;
    P.dyn.raw_off = pl.fwd.dyn.outsaved_off; P.dyn.nraw = pl.fwd.dyn.nout; P.dyn.has_density = p->dyn.has_density;
    P.dyn.lmax = p->dyn.max_log_std; P.dyn.z = p->dyn.z; P.dyn.zstride = p->dyn.z_step_stride;
    P.act_scale = p->act_scale; P.Sy = p->Sy;
    P.KR = p->rew_rows; P.rew_C = p->rew_C; P.rew_c0 = p->rew_c0; P.rew_Q = p->rew_Q; P.rew_R = p->rew_R;
    P.rew_scale = p->rew_scale; P.rew_offset = p->rew_offset;
    pl.tc = C;
    return PMB_OK;
}

static int build_plan(const pmb_problem *p, const pmb_tuning *tune, Plan &pl) {
    if (!p) return fail(PMB_E_INVALID, "problem is NULL");
    if (p->N < 1 || p->H < 1 || p->D < 1 || p->U < 1) return fail(PMB_E_INVALID, "N, H, D, U must be >= 1");
    if (p->D + p->U > PMB_MAX_STATE) return fail(PMB_E_UNSUPPORTED, "D+U=%d > %d", p->D + p->U, PMB_MAX_STATE);
    if (p->rew_rows < 1 || p->rew_rows > PMB_MAX_REWARD_ROWS) return fail(PMB_E_INVALID, "rew_rows=%d", p->rew_rows);
    if (!p->act_scale || !p->act_bias || !p->mx || !p->iSx || !p->my || !p->Sy || !p->rew_C || !p->rew_c0 ||
        !p->rew_Q || !p->rew_R)
        return fail(PMB_E_INVALID, "a scaler / reward pointer is NULL");
    int rc;
    if ((rc = check_net(p->pol, p->D, p->pol.has_density ? 2 * p->U : p->U, "policy"))) return rc;
    if ((rc = check_net(p->dyn, p->D + p->U, p->dyn.has_density ? 2 * p->D : p->D, "dynamics"))) return rc;
    const bool mm = p->mm_states || p->mm_rewards;
    const int G = p->mm_groups > 1 ? p->mm_groups : 1;
    if (mm) {
        if (p->mm_states && !p->z_mm) return fail(PMB_E_INVALID, "mm_states needs z_mm");
        if (p->mm_rewards && !p->z_rr) return fail(PMB_E_INVALID, "mm_rewards needs z_rr");
        if (p->N % G != 0) return fail(PMB_E_INVALID, "N=%d is not divisible by mm_groups=%d", p->N, G);
        if (p->N / G < 2) return fail(PMB_E_INVALID, "moment matching needs at least 2 particles per group");
        if (p->n_global != p->N && !(p->mm_world > 1))
            return fail(PMB_E_INVALID, "n_global=%d differs from N=%d without mm_world > 1", p->n_global, p->N);
    }

    memset(&pl, 0, sizeof(pl));
    pl.mm_G = G;
    int P = tune && tune->particles_per_cta ? tune->particles_per_cta : 0;
    const bool autoP = P == 0;
    if (P == 0) P = (p->N > 8 * 148) ? 8 : (p->N > 2 * 148) ? 4 : (p->N > 148) ? 2 : 1;   // few particles: spread over more SMs
    // moment matching: every step ends in a grid barrier whose cost grows with the CTA count (c3, 100 particles:
    // 22.1 / 17.4 / 16.3 / 19.7 ms per iteration at P = 1 / 2 / 4 / 8, profiles/r02_tc_timelines_v1_v2.txt)
    if (autoP && p->mm_states && P < 4 && p->N >= 16) P = 4;
    if (P != 1 && P != 2 && P != 4 && P != 8) return fail(PMB_E_INVALID, "particles_per_cta must be 1, 2, 4 or 8");
    if (p->mm_states) {
        // every CTA of the grid takes part in a per-step barrier: CTAs must not straddle groups and the
        // whole grid must be co-resident (one CTA per SM)
        if (G > 1) while (P > 1 && (p->N / G) % P != 0) P >>= 1;
        int sms = 148, dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        while (P < 8 && (p->N + P - 1) / P > sms && (G == 1 || (p->N / G) % (2 * P) == 0)) P <<= 1;
        if ((p->N + P - 1) / P > sms)
            return fail(PMB_E_UNSUPPORTED, "mm_states with N=%d particles does not fit one co-resident grid", p->N);
    }
    pl.P = P;
    pl.stream_mode = tune && tune->stream_mode ? tune->stream_mode : 2;
    if (pl.stream_mode < 1 || pl.stream_mode > 5) return fail(PMB_E_INVALID, "stream_mode must be 0..5");
    if (pl.stream_mode >= 3) pl.stream_mode = 2;
    // split-K slices of the batched weight gradient: 64, or as many as keep a slice <= 1024 rows (the bound up to which
    // the tcgen05 kernel's TMEM accumulation stays inside the gradient budget; c5: 250 k rows -> 256 slices)
    int auto_split = 64;
    while (auto_split < 1024 && ((long long)p->H * p->N + auto_split - 1) / auto_split > 1024) auto_split <<= 1;
    pl.nsplit = tune && tune->wgrad_splits ? tune->wgrad_splits : auto_split;
    if (pl.nsplit < 1 || pl.nsplit > 1024) return fail(PMB_E_INVALID, "wgrad_splits outside [1,1024]");

    Alloc wf, wb, ws;
    long long dummy[2] = {0, 0};
    SweepParams &F = pl.fwd, &B = pl.bwd;
    plan_net(p->pol, p->N, p->H, true, F.pol, B.pol, pl, wf, wb, ws, dummy);
    plan_net(p->dyn, p->N, p->H, false, F.dyn, B.dyn, pl, wf, wb, ws, dummy);
    // packed weight areas live after the generic regions
    pl.wpack_fwd_off = ws.take(wf.top);
    pl.wpack_bwd_off = ws.take(wb.top);
    // flat gradient layout + weight-gradient jobs
    long long np = 0;
    pl.n_wg = p->pol.n_linear;
    const int L = p->pol.n_linear - 1;
    for (int l = 0; l <= L; ++l) {
        WGrad &g = pl.wg[l];
        const int out_l = p->pol.dims[l + 1], in_l = p->pol.dims[l];
        g.delta_off = F.pol.delta_off[l];
        g.lda = (l < L) ? round4(out_l) : out_l;
        g.M = out_l;
        g.inp_off = (l == 0) ? -1 : F.pol.saved_off[l - 1];
        g.ldb = (l == 0) ? in_l : round4(in_l);
        g.Nc = in_l;
        g.w_off = np; np += (long long)out_l * in_l;
        g.b_off = -1;
        if (p->pol.b[l]) { g.b_off = np; np += out_l; }
    }
    pl.nparam = np;
    pl.part_off = ws.take((long long)pl.nsplit * np);
    pl.cl_pre_off = ws.take((long long)p->H * p->N * (2 * p->D + 3 * p->U));
    if (mm) {
        const long long HN = (long long)p->H * p->N;
        const int grid = (p->N + P - 1) / P;
        pl.s1pre_off = ws.take(HN * p->D);
        pl.mmstat_off = ws.take((long long)p->H * G * (3 * SD + SD * SD));
        pl.mmrec_off = ws.take(2LL * grid * MMREC * 2);      // doubles
        pl.mmctr_off = ws.take(32);
        pl.rpre_off = ws.take(HN);
        pl.rstat_off = ws.take((long long)p->H * G * 4);
        pl.geff_off = ws.take(HN);
        pl.cmm_gbuf_off = ws.take(2LL * 2 * 2 * CMM_TILES * CMM_NQ); // [2 parities][tiles][nq] 16-byte tagged records of the cluster sweeps
        pl.mm_world = p->mm_world > 1 ? p->mm_world : 1;
        if (pl.mm_world > 1 && p->mm_rewards) {
            const long long HNg = (long long)p->H * p->n_global;
            pl.rfull_off = ws.take(HNg);
            pl.routfull_off = ws.take(HNg);
            pl.gfull_off = ws.take(HNg);
            pl.gefffull_off = ws.take(HNg);
        }
    }
    if ((rc = plan_tc(p, tune, pl, ws)) != PMB_OK) return rc;
    const long long ws_before_cw = ws.top;
    (void)ws_before_cw;
    pl.cw = 0;

    for (int pass = 0; pass < 2; ++pass) {
        SweepParams &S = pass ? B : F;
        S.N = p->N; S.H = p->H; S.D = p->D; S.U = p->U;
        S.act_scale = p->act_scale; S.act_bias = p->act_bias;
        S.mx = p->mx; S.iSx = p->iSx; S.my = p->my; S.Sy = p->Sy;
        S.KR = p->rew_rows; S.rew_C = p->rew_C; S.rew_c0 = p->rew_c0; S.rew_Q = p->rew_Q; S.rew_R = p->rew_R;
        S.rew_scale = p->rew_scale; S.rew_offset = p->rew_offset;
        S.mm_states = p->mm_states; S.mm_rewards = p->mm_rewards; S.mm_G = G; S.mm_Ng = p->N / G;
        S.z_mm = p->z_mm;
    }
    const NetSweep *fo[2] = {&F.pol, &F.dyn};
    const NetSweep *bo[2] = {&B.dyn, &B.pol};
    pl.ws_floats = ws.top;
    if (pl.tc) return PMB_OK;
    if ((rc = plan_cw(p, tune, pl, ws)) != PMB_OK) return rc;
    pl.ws_floats = ws.top;
    if (pl.cw) { pl.cluster = 0; return PMB_OK; }
    if ((rc = plan_cluster(p, tune, pl)) != PMB_OK) return rc;
    if (p->mm_world > 1 && mm) {
        if (p->mm_world > PMB_MAX_PEERS || p->mm_rank < 0 || p->mm_rank >= p->mm_world || p->n_global != p->N * p->mm_world)
            return fail(PMB_E_INVALID, "mm_world=%d mm_rank=%d n_global=%d do not describe equal shards of N=%d", p->mm_world,
                        p->mm_rank, p->n_global, p->N);
        if (G > 1) return fail(PMB_E_UNSUPPORTED, "moment matching across GPUs supports one matching group");
        if (p->mm_states && !pl.cluster)
            return fail(PMB_E_UNSUPPORTED, "moment matching of the states across GPUs runs on the cluster-resident sweeps "
                                           "(two hidden layers <= 256 wide, <= 120 particles per GPU)");
    }
    if (pl.cluster) return PMB_OK;
    const int nst = tune ? tune->reserved[1] : 0;
    if ((rc = plan_sweep(F, fo, false, P, pl.stream_mode, nst, p->mm_states != 0)) < 0) return rc;
    pl.smem_fwd_bytes = rc;
    if ((rc = plan_sweep(B, bo, true, P, pl.stream_mode, nst, p->mm_states != 0)) < 0) return rc;
    pl.smem_bwd_bytes = rc;
    return PMB_OK;
}

static void resolve(Plan &pl, float *ws) {
    for (int i = 0; i < pl.jobs.n; ++i) {
        long long v = pl.job_dst_off[i];
        int area = (int)(v >> 60);
        long long off = v & ((1LL << 60) - 1);
        long long base = area == 0 ? pl.wpack_fwd_off : area == 1 ? pl.wpack_bwd_off : 0;
        pl.jobs.job[i].dst = ws + base + off;
    }
    pl.fwd.ws = pl.bwd.ws = ws;
    pl.fwd.wpack = ws + pl.wpack_fwd_off;
    pl.bwd.wpack = ws + pl.wpack_bwd_off;
    if (pl.tc) {
        for (int pass = 0; pass < 2; ++pass) {
            TcParams &T = pass ? pl.tbwd : pl.tfwd;
            T.ws = ws;
            T.wpack = ws + pl.tc_wpack_off;
            T.xbuf = ws + pl.tc_xbuf_off;
            T.opart = ws + pl.tc_opart_off;
            T.pre = ws + pl.cl_pre_off;
            if (T.mm_states) {
                T.s1pre = ws + pl.s1pre_off;
                T.mmstat = ws + pl.mmstat_off;
            }
        }
        pl.tpre.ws = ws;
        pl.tpre.pre = ws + pl.cl_pre_off;
        pl.tpre.s1pre = pl.tfwd.mm_states ? ws + pl.s1pre_off : nullptr;
        for (int i = 0, k = 0; i < pl.jobs.n; ++i)
            if ((pl.job_dst_off[i] >> 60) == 2) pl.sjobs.job[k++].dst = ws + (pl.job_dst_off[i] & ((1LL << 60) - 1));
    }
    if (pl.cw) {
        for (int pass = 0; pass < 2; ++pass) {
            ClusterParams &W = pass ? pl.wbwd : pl.wfwd;
            W.ws = ws;
            W.wpack = ws + (pass ? pl.wpack_bwd_off : pl.wpack_fwd_off);
            W.pre = ws + pl.cl_pre_off;
            W.g1 = reinterpret_cast<unsigned *>(ws + pl.cw_g1_off);
            W.g2 = reinterpret_cast<unsigned *>(ws + pl.cw_g2_off);
        }
    }
    pl.cfwd.ws = pl.cbwd.ws = ws;
    if (pl.cluster && pl.cfwd.mm_states) {
        for (int pass = 0; pass < 2; ++pass) {
            ClusterParams &P = pass ? pl.cbwd : pl.cfwd;
            P.s1pre = ws + pl.s1pre_off;
            P.mmstat = ws + pl.mmstat_off;
            P.mmctr = reinterpret_cast<unsigned *>(ws + pl.mmctr_off) + pass;   // one counter per sweep
            P.mmrec = reinterpret_cast<double *>(ws + pl.cmm_gbuf_off);
            P.mmrec_peer[0] = P.mmrec;
            P.mmctr_peer[0] = P.mmctr;
            P.mm_base = nullptr;
            P.mm_base_next = nullptr;
        }
    }
    pl.cbwd.pre = ws + pl.cl_pre_off;
    pl.cfwd.wpack = ws + pl.wpack_fwd_off;
    pl.cbwd.wpack = ws + pl.wpack_bwd_off;
    if (pl.fwd.mm_states || pl.fwd.mm_rewards) {
        for (int pass = 0; pass < 2; ++pass) {
            SweepParams &S = pass ? pl.bwd : pl.fwd;
            S.s1pre = ws + pl.s1pre_off;
            S.mmstat = ws + pl.mmstat_off;
            S.mmrec = reinterpret_cast<double *>(ws + pl.mmrec_off);
            S.mmctr = reinterpret_cast<unsigned *>(ws + pl.mmctr_off) + pass;   // one counter per sweep
        }
    }
}

// ---- moment matching across GPUs: geometry of the exchange areas and their binding to the sweep parameters ----
struct MmExchange {
    int nq, ntiles_total;
    size_t sweep_doubles;      // [2 parities][ntiles_total * nq] doubles of one sweep
    size_t ctr_off;            // byte offset of the two arrival counters inside a rank's record area
    size_t rec_bytes, gather_bytes, state_bytes;
};
static MmExchange mm_exchange(const pmb_problem *p, const Plan &pl) {
    MmExchange x;
    x.nq = p->D + p->D * (p->D + 1) / 2;
    x.ntiles_total = 2 * (pl.cluster ? pl.cl_nclusters : 1) * (p->mm_world > 1 ? p->mm_world : 1);
    x.sweep_doubles = (size_t)2 * x.ntiles_total * x.nq * 2;      // [2 parities][tiles][nq] 16-byte tagged entries
    x.ctr_off = (2 * x.sweep_doubles * sizeof(double) + 255) & ~(size_t)255;
    x.rec_bytes = x.ctr_off + 256;
    x.gather_bytes = pmb_peer_buffer_bytes((long long)p->H * p->N, p->mm_world > 1 ? p->mm_world : 1);
    x.state_bytes = 64;
    return x;
}
// after resolve(): point the cluster sweeps at the peer-mapped record areas / counters (mm_world > 1)
static int bind_peers(Plan &pl, const pmb_problem *p) {
    if (!(p->mm_world > 1 && (p->mm_states || p->mm_rewards))) return PMB_OK;
    if (!p->mm_local_state) return fail(PMB_E_INVALID, "mm_local_state is NULL");
    for (int r = 0; r < p->mm_world; ++r)
        if ((p->mm_states && !p->mm_peer_rec[r]) || (p->mm_rewards && !p->mm_peer_gather[r]))
            return fail(PMB_E_INVALID, "exchange area of rank %d is NULL", r);
    if (!p->mm_states) return PMB_OK;
    const MmExchange x = mm_exchange(p, pl);
    unsigned long long *st = reinterpret_cast<unsigned long long *>(p->mm_local_state);
    for (int pass = 0; pass < 2; ++pass) {
        ClusterParams &P = pass ? pl.cbwd : pl.cfwd;
        for (int r = 0; r < p->mm_world; ++r) {
            P.mmrec_peer[r] = reinterpret_cast<double *>(p->mm_peer_rec[r]) + pass * x.sweep_doubles;
            P.mmctr_peer[r] = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(p->mm_peer_rec[r]) + x.ctr_off) + pass;
        }
        P.mmrec = P.mmrec_peer[p->mm_rank];
        P.mmctr = P.mmctr_peer[p->mm_rank];
        P.mm_base = st + pass;
        P.mm_base_next = st + pass;
    }
    return PMB_OK;
}

}  // namespace pmb

using namespace pmb;

#define PMB_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) return fail(PMB_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

extern "C" {

int pmb_abi_version(void) { return PMB_ABI_VERSION; }

const char *pmb_last_error(void) { return g_err; }

int pmb_check_problem(const pmb_problem *p, const pmb_tuning *tune) {
    Plan pl;
    return build_plan(p, tune, pl);
}

size_t pmb_workspace_bytes(const pmb_problem *p, const pmb_tuning *tune) {
    Plan pl;
    if (build_plan(p, tune, pl) != PMB_OK) return 0;
    return (size_t)pl.ws_floats * sizeof(float);
}

int pmb_mm_exchange_bytes(const pmb_problem *p, const pmb_tuning *tune, size_t out[3]) {
    if (!out) return fail(PMB_E_INVALID, "out is NULL");
    Plan pl;
    int rc = build_plan(p, tune, pl);
    if (rc != PMB_OK) return rc;
    const MmExchange x = mm_exchange(p, pl);
    out[0] = x.rec_bytes; out[1] = x.gather_bytes; out[2] = x.state_bytes;
    return PMB_OK;
}

int pmb_plan_describe(const pmb_problem *p, const pmb_tuning *tune, pmb_plan_info *info) {
    if (!info) return fail(PMB_E_INVALID, "info is NULL");
    Plan pl;
    int rc = build_plan(p, tune, pl);
    if (rc != PMB_OK) return rc;
    memset(info, 0, sizeof(*info));
    if (pl.tc) {
        info->variant = 2;
        info->ctas = pl.tfwd.ntiles * pl.tc;
        info->threads_per_cta = TC_NTL;
        info->cluster_size = pl.tc;
        info->particles_per_group = pl.tfwd.TP;
        info->smem_fwd_bytes = pl.tfwd.smem_floats * 4;
        info->smem_bwd_bytes = pl.tbwd.smem_floats * 4;
    } else if (pl.cw) {
        info->variant = 3;
        info->ctas = pl.cw_nclusters * pl.cw;
        info->threads_per_cta = CW_NT;
        info->cluster_size = pl.cw;
        info->particles_per_group = pl.wfwd.PG;
        info->smem_fwd_bytes = pl.wfwd.smem_floats * 4;
        info->smem_bwd_bytes = pl.wbwd.smem_floats * 4;
    } else if (pl.cluster) {
        info->variant = 1;
        info->ctas = pl.cl_nclusters * pl.cluster;
        info->threads_per_cta = CL_NT;
        info->cluster_size = pl.cluster;
        info->particles_per_group = pl.cfwd.PG;
        info->smem_fwd_bytes = pl.cfwd.smem_floats * 4;
        info->smem_bwd_bytes = pl.cbwd.smem_floats * 4;
    } else {
        info->variant = 0;
        info->ctas = (p->N + pl.P - 1) / pl.P;
        info->threads_per_cta = NT_LAUNCH;
        info->cluster_size = 1;
        info->particles_per_group = pl.P;
        info->smem_fwd_bytes = pl.smem_fwd_bytes;
        info->smem_bwd_bytes = pl.smem_bwd_bytes;
    }
    // pack + sweep (+ reward matching); [reward matching adjoint] + [adjoint-factor pre-pass] + sweep +
    // one weight-gradient kernel per policy layer + partial reduction
    info->launches_fwd = 2 + (pl.tc ? 1 : 0) + (p->mm_rewards ? 1 : 0);
    info->launches_bwd = (p->mm_rewards ? 1 : 0) + ((pl.cluster || pl.tc || pl.cw) ? 1 : 0) + 1 + pl.n_wg + 1;
    return PMB_OK;
}

size_t pmb_policy_param_count(const pmb_problem *p) {
    if (!p) return 0;
    size_t n = 0;
    for (int l = 0; l < p->pol.n_linear; ++l) {
        n += (size_t)p->pol.dims[l + 1] * p->pol.dims[l];
        if (p->pol.b[l]) n += p->pol.dims[l + 1];
    }
    return n;
}

int pmb_rollout_forward(const pmb_problem *p, const pmb_tuning *tune, const float *x0, float *states,
                        float *actions, float *rewards, void *workspace, size_t workspace_bytes,
                        int *status_dev, void *stream) {
    Plan pl;
    int rc = build_plan(p, tune, pl);
    if (rc != PMB_OK) return rc;
    if (!x0 || !states || !actions || !rewards || !workspace) return fail(PMB_E_INVALID, "NULL tensor argument");
    if (workspace_bytes < (size_t)pl.ws_floats * sizeof(float))
        return fail(PMB_E_WORKSPACE, "workspace has %zu bytes, need %zu", workspace_bytes, (size_t)pl.ws_floats * 4);
    cudaStream_t st = (cudaStream_t)stream;
    resolve(pl, (float *)workspace);
    if ((rc = bind_peers(pl, p)) != PMB_OK) return rc;
    const bool sharded_mm = p->mm_world > 1 && (p->mm_states || p->mm_rewards);
    const int phases = (tune && (tune->reserved[0] & 255)) ? (tune->reserved[0] & 255) : 7;   // profiling aid: 1 pack, 2 sweep
    if (status_dev) PMB_CUDA(cudaMemsetAsync(status_dev, 0, sizeof(int), st));
    if (phases & 1) {
        if (pl.tc) {
            PMB_CUDA(launch_pack(pl.sjobs, st));
            PMB_CUDA(launch_tc_pack(pl.tjobs, (float *)workspace + pl.tc_wpack_off, st));
        } else {
            PMB_CUDA(launch_pack(pl.jobs, st));
        }
    }
    SweepParams &F = pl.fwd;
    float *wsf = (float *)workspace;
    F.x0 = x0; F.states = states; F.actions = actions; F.status = status_dev;
    // with mm_rewards the sweep writes the pre-matching rewards; a whole-horizon kernel matches them
    F.rewards = p->mm_rewards ? wsf + pl.rpre_off : rewards;
    // (across GPUs the record areas live in peer-mapped memory and are never reset -- a peer may already be a launch
    // ahead --, the tags just keep growing; on one GPU the tags restart at 1 with every launch, so the area is zeroed)
    if (p->mm_states && !pl.tc && !sharded_mm) {
        PMB_CUDA(cudaMemsetAsync(wsf + pl.mmctr_off, 0, 32 * sizeof(float), st));
        if (pl.cluster) PMB_CUDA(cudaMemsetAsync(wsf + pl.cmm_gbuf_off, 0, sizeof(float) * 2 * 2 * 2 * CMM_TILES * CMM_NQ, st));
    }
    F.dbg = tune ? (long long *)(((unsigned long long)(unsigned)tune->reserved[3] << 32) | (unsigned)tune->reserved[2]) : nullptr;
    if (pl.tc) {
        TcParams &T = pl.tfwd;
        T.x0 = x0; T.states = states; T.actions = actions; T.rewards = F.rewards; T.status = status_dev; T.dbg = F.dbg;
        if (phases & 2) PMB_CUDA(launch_tc_fwd(T, st));
    } else if (pl.cw) {
        ClusterParams &CF = pl.wfwd;
        CF.x0 = x0; CF.states = states; CF.actions = actions; CF.rewards = F.rewards; CF.dbg = F.dbg;
        if (phases & 2) PMB_CUDA(launch_cw_fwd(CF, pl.cw_nclusters, st));
    } else if (pl.cluster) {
        ClusterParams &CF = pl.cfwd;
        CF.x0 = x0; CF.states = states; CF.actions = actions; CF.rewards = F.rewards; CF.dbg = F.dbg;
        CF.status = status_dev;
        if (phases & 2) PMB_CUDA(launch_cluster_fwd(CF, pl.cl_nclusters, st));
    } else if (phases & 2) {
        PMB_CUDA(launch_rollout_fwd(F, pl.P, pl.smem_fwd_bytes, st));
    }
    if (p->mm_rewards && sharded_mm) {
        // the rewards of ALL ranks are matched together: all-gather the pre-matching rewards over peer memory, match the
        // full [H][n_global] array (every rank computes the same statistics), keep this rank's columns
        unsigned long long *gst = reinterpret_cast<unsigned long long *>(p->mm_local_state) + 2;
        PMB_CUDA(launch_peer_gather(wsf + pl.rpre_off, wsf + pl.rfull_off, p->H, p->N, p->mm_world, p->mm_rank, p->mm_peer_gather,
                                    gst, st));
        PMB_CUDA(launch_reward_mm_fwd(wsf + pl.rfull_off, wsf + pl.routfull_off, p->z_rr, wsf + pl.rstat_off, p->n_global, p->H,
                                      1, status_dev, st));
        PMB_CUDA(launch_take_columns(wsf + pl.routfull_off, rewards, p->H, p->n_global, p->N, p->mm_rank * p->N, st));
    } else if (p->mm_rewards) {
        PMB_CUDA(launch_reward_mm_fwd(wsf + pl.rpre_off, rewards, p->z_rr, wsf + pl.rstat_off, p->N, p->H, pl.mm_G,
                                      status_dev, st));
    }
    return PMB_OK;
}

int pmb_rollout_backward(const pmb_problem *p, const pmb_tuning *tune, const float *states, const float *actions,
                         const float *rewards, const float *g_states, const float *g_actions,
                         const float *g_rewards, float *grad_flat, float *dx0, float *da_total, void *workspace,
                         size_t workspace_bytes, void *stream) {
    Plan pl;
    int rc = build_plan(p, tune, pl);
    if (rc != PMB_OK) return rc;
    if (!states || !actions || !rewards || !grad_flat || !workspace) return fail(PMB_E_INVALID, "NULL tensor argument");
    if (workspace_bytes < (size_t)pl.ws_floats * sizeof(float))
        return fail(PMB_E_WORKSPACE, "workspace has %zu bytes, need %zu", workspace_bytes, (size_t)pl.ws_floats * 4);
    cudaStream_t st = (cudaStream_t)stream;
    float *ws = (float *)workspace;
    resolve(pl, ws);
    if ((rc = bind_peers(pl, p)) != PMB_OK) return rc;
    const bool sharded_mm = p->mm_world > 1 && (p->mm_states || p->mm_rewards);
    SweepParams &B = pl.bwd;
    B.states = const_cast<float *>(states); B.actions = const_cast<float *>(actions);
    B.rewards = const_cast<float *>(rewards);
    B.g_states = g_states; B.g_actions = g_actions; B.g_rewards = g_rewards; B.dx0 = dx0; B.da_total = da_total;
    if (p->mm_rewards) {
        // adjoint of the reward matching runs first, off the serial chain; the sweep then sees the
        // pre-matching rewards and their cotangent
        B.rewards = ws + pl.rpre_off;
        if (g_rewards && sharded_mm) {
            // adjoint of the matching over all ranks' rewards: all-gather the cotangents, run the adjoint on the full
            // [H][n_global] arrays (the gathered pre-matching rewards are still in place), keep this rank's columns
            unsigned long long *gst = reinterpret_cast<unsigned long long *>(p->mm_local_state) + 2;
            PMB_CUDA(launch_peer_gather(g_rewards, ws + pl.gfull_off, p->H, p->N, p->mm_world, p->mm_rank, p->mm_peer_gather, gst, st));
            PMB_CUDA(launch_reward_mm_bwd(ws + pl.gfull_off, ws + pl.rfull_off, p->z_rr, ws + pl.rstat_off, ws + pl.gefffull_off,
                                          p->n_global, p->H, 1, st));
            PMB_CUDA(launch_take_columns(ws + pl.gefffull_off, ws + pl.geff_off, p->H, p->n_global, p->N, p->mm_rank * p->N, st));
            B.g_rewards = ws + pl.geff_off;
        } else if (g_rewards) {
            PMB_CUDA(launch_reward_mm_bwd(g_rewards, ws + pl.rpre_off, p->z_rr, ws + pl.rstat_off, ws + pl.geff_off,
                                          p->N, p->H, pl.mm_G, st));
            B.g_rewards = ws + pl.geff_off;
        }
    }
    if (p->mm_states && !pl.tc && !sharded_mm) {
        PMB_CUDA(cudaMemsetAsync(reinterpret_cast<unsigned *>(ws + pl.mmctr_off) + 1, 0, sizeof(unsigned), st));
        if (pl.cluster) PMB_CUDA(cudaMemsetAsync(ws + pl.cmm_gbuf_off, 0, sizeof(float) * 2 * 2 * 2 * CMM_TILES * CMM_NQ, st));
    }
    B.dbg = tune ? (long long *)(((unsigned long long)(unsigned)tune->reserved[3] << 32) | (unsigned)tune->reserved[2]) : nullptr;
    const int phases = (tune && (tune->reserved[0] & 255)) ? (tune->reserved[0] & 255) : 7;   // profiling aid: 2 sweep, 4 wgrad
    if (pl.tc) {
        TcParams &T = pl.tbwd;
        T.states = B.states; T.actions = B.actions; T.rewards = B.rewards;
        T.g_states = B.g_states; T.g_actions = B.g_actions; T.g_rewards = B.g_rewards; T.dx0 = B.dx0; T.dbg = B.dbg;
        T.da_total = da_total;
        ClusterParams &P = pl.tpre;
        P.states = B.states; P.actions = B.actions; P.rewards = B.rewards;
        P.g_states = B.g_states; P.g_actions = B.g_actions; P.g_rewards = B.g_rewards;
        if (phases & 2) {
            PMB_CUDA(launch_bwd_pre(P, st));
            PMB_CUDA(launch_tc_bwd(T, st));
        }
    } else if (pl.cw) {
        ClusterParams &CB = pl.wbwd;
        CB.states = B.states; CB.actions = B.actions; CB.rewards = B.rewards;
        CB.g_states = B.g_states; CB.g_actions = B.g_actions; CB.g_rewards = B.g_rewards; CB.dx0 = B.dx0; CB.dbg = B.dbg;
        CB.da_total = da_total;
        if (phases & 2) {
            PMB_CUDA(launch_bwd_pre(CB, st));
            PMB_CUDA(launch_cw_bwd(CB, pl.cw_nclusters, st));
        }
    } else if (pl.cluster) {
        ClusterParams &CB = pl.cbwd;
        CB.states = B.states; CB.actions = B.actions; CB.rewards = B.rewards;
        CB.g_states = B.g_states; CB.g_actions = B.g_actions; CB.g_rewards = B.g_rewards; CB.dx0 = B.dx0; CB.dbg = B.dbg;
        CB.da_total = da_total;
        if (phases & 2) PMB_CUDA(launch_cluster_bwd(CB, pl.cl_nclusters, st));
    } else if (phases & 2) {
        PMB_CUDA(launch_rollout_bwd(B, pl.P, pl.smem_bwd_bytes, st));
    }
    if (!(phases & 4)) return PMB_OK;
    // batched policy weight gradient over the (H*N) axis
    const long long R = (long long)p->H * p->N;
    float *part = ws + pl.part_off;
    for (int l = 0; l < pl.n_wg; ++l) {
        const WGrad &g = pl.wg[l];
        const float *A = ws + g.delta_off;
        const float *Bm = g.inp_off < 0 ? states : ws + g.inp_off;
        PMB_CUDA(launch_wgrad(A, g.lda, g.M, Bm, g.ldb, g.Nc, R, pl.nsplit, part + g.w_off,
                              g.b_off >= 0 ? part + g.b_off : nullptr, pl.nparam, st,
                              tune ? tune->reserved[4] : 0));
    }
    PMB_CUDA(launch_reduce_partials(part, pl.nparam, pl.nsplit, grad_flat, st));
    return PMB_OK;
}

int pmb_clip_adam_step(const pmb_adam_tensor *table_dev, int n_tensors, float max_norm, float lr, float beta1,
                       float beta2, float eps, long long step, long long *step_dev, float *scratch_dev,
                       const int *skip_if_nonzero, void *stream) {
    if (!table_dev || n_tensors < 1 || !scratch_dev || (!step_dev && step < 1))
        return fail(PMB_E_INVALID, "bad optimiser arguments");
    PMB_CUDA(launch_clip_adam(table_dev, n_tensors, max_norm, lr, beta1, beta2, eps, step, step_dev, scratch_dev,
                              skip_if_nonzero, (cudaStream_t)stream));
    return PMB_OK;
}

}  // extern "C"
