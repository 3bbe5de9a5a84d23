// Dynamics-model fit on the GPU: one minibatch iteration of the reference's
// utils.train_regressor (reference utils/train_regressor.py:58-165, default branch) for a Regressor built by
// models.mlp with concrete-dropout layers (models/core.py:121-187, models/modules.py:73-171) and a
// DiagGaussianDensity output (models/densities.py:87-144):
//     loss = -mean_b log N(y_b | mean_b, exp(log_std_b)^2) + reg_weight * R(theta) / N
// on the WHITENED dataset.  The adjoint formulas are those of oracle/train_regressor_oracle.py differentiated by
// hand; the uniform noise u and the hard Bernoulli samples b of every dropout layer are INPUTS (drawn with torch's
// generator in the reference's order, so the RNG stream is the reference's).
//
//   fit_rows_kernel   : forward + backward of 4 minibatch rows per CTA (weights from L2: 85 K floats for 2x[200]);
//                       writes the per-layer output adjoints and inputs for the weight gradients, the per-row
//                       log-likelihood and the per-row adjoint of every dropout probability
//   launch_wgrad      : dW_l, db_l = adjoint^T [input | 1] over the minibatch rows (pmb_wgrad.cu, FFMA2 tiles)
//   fit_finalize_kernel: + regulariser gradients (weights, biases, dropout logits), dropout-logit gradients summed
//                       over the rows in a fixed order, mean log-likelihood
// The optimiser step is pmb_clip_adam_step (max_norm = 0: train_regressor does not clip).
#include <stdio.h>
#include <string.h>

#include "pmb_host.h"
#include "pmb_internal.cuh"

namespace pmb {

constexpr int FIT_RB = 4;        // minibatch rows per CTA
constexpr int FIT_NT = 256;

struct FitParams {
    int N, M, L;                       // dataset rows, minibatch rows, hidden layers
    int dims[MAXL + 1];                // dims[0] = inputs, dims[l+1] = outputs of linear l; dims[L+1] = 2 * Dout
    const float *W[MAXL], *b[MAXL];
    const float *logit_p[MAXL];
    const float *u[MAXL], *hard[MAXL]; // [M][h_l] uniform noise, hard Bernoulli sample
    float temp[MAXL], reg_scale[MAXL], drop_reg[MAXL];
    float lmax, reg_weight;
    const float *Xw, *Yw;              // whitened dataset [N][dims[0]], [N][Dout]
    const long long *idx;              // [M] minibatch rows
    // scratch (workspace)
    float *inp[MAXL];                  // input of linear l for every row [M][dims[l]]
    float *delta[MAXL];                // adjoint of linear l's output [M][dims[l+1]]
    float *gmask[MAXL];                // per-row adjoint of the dropout probability's logit [M][h_l]
    float *ll;                         // [M] per-row log-likelihood
    float *mask_out[MAXL];             // [M][h_l] the concrete mask of this iteration (module buffer refresh), nullable
    float *p_out[MAXL];                // [h_l] sigmoid(logit_p) of this iteration, nullable
    int hmax;                          // widest layer (shared-memory tile stride)
    // outputs
    float *grad;                       // flat, parameters() order: W0, b0, logit_p0, W1, ..., W_L, b_L
    long long w_off[MAXL], b_off[MAXL], p_off[MAXL];
    float *loglik;                     // [1] mean log-likelihood of the minibatch
};

// out[r][j] = bias[j] + sum_k in[r][k] W[j][k]  for the CTA's FIT_RB rows; warps own output columns, lanes walk k
__device__ __forceinline__ void fit_linear(const float *__restrict__ W, const float *__restrict__ bias, int K, int Nout,
                                           const float *in, int in_ld, float *out, int out_ld) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < Nout; j += FIT_NT / 32) {
        float acc[FIT_RB];
#pragma unroll
        for (int r = 0; r < FIT_RB; ++r) acc[r] = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float w = __ldg(W + (size_t)j * K + k);
#pragma unroll
            for (int r = 0; r < FIT_RB; ++r) acc[r] = fmaf(in[r * in_ld + k], w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < FIT_RB; ++r) acc[r] = warp_sum(acc[r]);
        if (lane == 0) {
            const float bj = bias ? __ldg(bias + j) : 0.f;
#pragma unroll
            for (int r = 0; r < FIT_RB; ++r) out[r * out_ld + j] = acc[r] + bj;
        }
    }
}

__global__ void __launch_bounds__(FIT_NT) fit_rows_kernel(const __grid_constant__ FitParams prm) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int L = prm.L, M = prm.M, H = prm.hmax;
    const int row0 = blockIdx.x * FIT_RB;
    // shared memory: x [RB][16] | per hidden layer: relu output r [RB][H], mask m [RB][H] | buf a, buf b [RB][H]
    float *xs = sm;
    float *rbuf = xs + FIT_RB * 16;
    float *mbuf = rbuf + (size_t)L * FIT_RB * H;
    float *bufa = mbuf + (size_t)L * FIT_RB * H;
    float *bufb = bufa + FIT_RB * H;
    const int Din = prm.dims[0], Dout = prm.dims[L + 1] / 2;
    for (int i = tid; i < FIT_RB * Din; i += FIT_NT) {
        const int r = i / Din, k = i - r * Din;
        const int row = min(row0 + r, M - 1);
        const float v = __ldg(prm.Xw + (size_t)prm.idx[row] * Din + k);
        xs[r * 16 + k] = v;
        if (row0 + r < M) prm.inp[0][(size_t)(row0 + r) * Din + k] = v;
    }
    __syncthreads();
    // ---------------- forward (models/core.py:169-187 with normalize=False, train-mode CDropout) ----------------
    // bufa: pre-activations of the current layer, bufb: post-dropout activations = input of the next layer
    const float *in = xs;
    int in_ld = 16;
    for (int l = 0; l < L; ++l) {
        const int K = prm.dims[l], h = prm.dims[l + 1];
        fit_linear(prm.W[l], prm.b[l], K, h, in, in_ld, bufa, H);
        __syncthreads();
        float *rl = rbuf + (size_t)l * FIT_RB * H, *ml = mbuf + (size_t)l * FIT_RB * H;
        for (int i = tid; i < FIT_RB * h; i += FIT_NT) {
            const int r = i / h, j = i - r * h;
            const int row = min(row0 + r, M - 1);
            // modules.py:102-118: probs = sigmoid((logit_p + log((u + 1e-7) / (1 - (u - 1e-7)))) / temp),
            // mask = (b - probs).detach() + probs
            const float uu = __ldg(prm.u[l] + (size_t)row * h + j), hb = __ldg(prm.hard[l] + (size_t)row * h + j);
            const float cp = __ldg(prm.logit_p[l] + j) + logf((uu + 1e-7f) / (1.f - (uu - 1e-7f)));
            const float probs = 1.f / (1.f + expf(-cp / prm.temp[l]));
            const float mk = (hb - probs) + probs;
            const float rr = fmaxf(bufa[r * H + j], 0.f);
            rl[r * H + j] = rr;
            ml[r * H + j] = mk;
            bufb[r * H + j] = rr * mk;
            if (row0 + r < M) {
                // d mask / d logit_p = probs (1 - probs) / temp   (parked here, multiplied by the adjoint below)
                prm.gmask[l][(size_t)(row0 + r) * h + j] = probs * (1.f - probs) / prm.temp[l];
                prm.inp[l + 1][(size_t)(row0 + r) * h + j] = rr * mk;
                if (prm.mask_out[l]) prm.mask_out[l][(size_t)(row0 + r) * h + j] = mk;
            }
        }
        __syncthreads();
        in = bufb;
        in_ld = H;
    }
    // output projection -> (mean, raw log-std) in bufa
    fit_linear(prm.W[L], prm.b[L], prm.dims[L], 2 * Dout, in, in_ld, bufa, H);
    __syncthreads();
    // ---------------- Gaussian NLL (densities.py:87-144) and its adjoint ----------------
    // log_std = lmax - softplus(lmax - raw);  log p = -0.5 sum ((mean - y) / std)^2 - sum log_std - D * 0.5 log(2 pi)
    // loss = -(1/M) sum_rows log p  =>  d loss/d mean = (mean - y) / std^2 / M,
    //                                   d loss/d raw  = (1 - ((mean - y) / std)^2) * sigmoid(lmax - raw) / M
    float *dout = bufb;      // [RB][2 Dout]: the hidden tile is no longer needed (the backward pass reads rbuf / mbuf)
    if (tid < FIT_RB * Dout) {
        const int r = tid / Dout, d = tid - r * Dout;
        const int row = min(row0 + r, M - 1);
        const float y = __ldg(prm.Yw + (size_t)prm.idx[row] * Dout + d);
        const float mean = bufa[r * H + d], raw = bufa[r * H + Dout + d];
        const float ls = prm.lmax - softplus_f(prm.lmax - raw);
        const float istd = expf(-ls);
        const float e = (mean - y) * istd;
        const float invM = 1.f / (float)M;
        dout[r * H + d] = e * istd * invM;
        dout[r * H + Dout + d] = (1.f - e * e) * sigmoid_f(prm.lmax - raw) * invM;
        bufa[r * H + 2 * Dout + d] = -0.5f * e * e - ls - 0.91893853320467274178f;     // log-likelihood term
    }
    __syncthreads();
    if (tid < FIT_RB && row0 + tid < M) {
        float s = 0.f;
        for (int d = 0; d < Dout; ++d) s += bufa[tid * H + 2 * Dout + d];
        prm.ll[row0 + tid] = s;
    }
    for (int i = tid; i < FIT_RB * 2 * Dout; i += FIT_NT) {
        const int r = i / (2 * Dout), o = i - r * 2 * Dout;
        if (row0 + r < M) prm.delta[L][(size_t)(row0 + r) * 2 * Dout + o] = dout[r * H + o];
    }
    __syncthreads();
    // ---------------- backward through the hidden layers ----------------
    const float *dl = dout;       // adjoint of linear (l+1)'s output, [RB][H]
    int nj = 2 * Dout;
    for (int l = L - 1; l >= 0; --l) {
        const int h = prm.dims[l + 1];
        const float *Wn = prm.W[l + 1];        // [nj][h]
        const float *rl = rbuf + (size_t)l * FIT_RB * H, *ml = mbuf + (size_t)l * FIT_RB * H;
        float *dn = (dl == bufa) ? bufb : bufa;
        for (int k = tid; k < h; k += FIT_NT) {
            float acc[FIT_RB];
#pragma unroll
            for (int r = 0; r < FIT_RB; ++r) acc[r] = 0.f;
            for (int j = 0; j < nj; ++j) {
                const float w = __ldg(Wn + (size_t)j * h + k);
#pragma unroll
                for (int r = 0; r < FIT_RB; ++r) acc[r] = fmaf(dl[r * H + j], w, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < FIT_RB; ++r) {
                // post = relu(pre) * mask:  d mask = d post * relu,  d pre = d post * mask * [pre > 0]
                const float rr = rl[r * H + k], mk = ml[r * H + k];
                const float dpre = rr > 0.f ? acc[r] * mk : 0.f;
                dn[r * H + k] = dpre;
                if (row0 + r < M) {
                    prm.delta[l][(size_t)(row0 + r) * h + k] = dpre;
                    prm.gmask[l][(size_t)(row0 + r) * h + k] *= acc[r] * rr;
                }
            }
        }
        __syncthreads();
        dl = dn;
        nj = h;
    }
}

// regulariser (modules.py:234-274, 87-93, 32-33), dropout-logit gradients, mean log-likelihood
__global__ void __launch_bounds__(256) fit_finalize_kernel(const __grid_constant__ FitParams prm) {
    const int L = prm.L, M = prm.M;
    const float rw = prm.reg_weight / (float)prm.N;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
    for (int l = 0; l < L; ++l) {
        const int h = prm.dims[l + 1], nout = prm.dims[l + 2];
        const float *Wn = prm.W[l + 1];                 // the Linear layer AFTER dropout l: [nout][h]
        const float sc = prm.reg_scale[l];
        // weights of the next layer: + rw * 2 * scale * p_j * W[i][j]
        for (long long i = gtid; i < (long long)nout * h; i += gsz) {
            const int j = (int)(i % h);
            const float p = sigmoid_f(__ldg(prm.logit_p[l] + j));
            prm.grad[prm.w_off[l + 1] + i] += rw * 2.f * sc * p * __ldg(Wn + i);
        }
        for (long long i = gtid; i < nout; i += gsz)
            if (prm.b[l + 1]) prm.grad[prm.b_off[l + 1] + i] += rw * 2.f * sc * __ldg(prm.b[l + 1] + i);
        // dropout logits: data term (fixed-order sum over the rows) + regulariser
        for (long long j = gtid; j < h; j += gsz) {
            float g = 0.f;
            for (int m = 0; m < M; ++m) g += prm.gmask[l][(size_t)m * h + j];
            float col = 0.f;
            for (int i = 0; i < nout; ++i) {
                const float w = __ldg(Wn + (size_t)i * h + j);
                col = fmaf(w, w, col);
            }
            const float p = sigmoid_f(__ldg(prm.logit_p[l] + j));
            // d/dp [scale p col + drop_reg (p log p + (1-p) log(1-p))] * p (1 - p)
            g += rw * (sc * col + prm.drop_reg[l] * (logf(p) - logf(1.f - p))) * p * (1.f - p);
            prm.grad[prm.p_off[l] + j] = g;
            if (prm.p_out[l]) prm.p_out[l][j] = p;
        }
    }
    if (gtid == 0) {
        float s = 0.f;
        for (int m = 0; m < M; ++m) s += prm.ll[m];
        prm.loglik[0] = s / (float)M;
    }
}

}  // namespace pmb

using namespace pmb;

static thread_local char g_fit_err[256] = "";
extern "C" const char *pmb_fit_last_error(void) { return g_fit_err; }

static int fit_fail(int code, const char *msg) {
    snprintf(g_fit_err, sizeof(g_fit_err), "%s", msg);
    return code;
}

static int fit_layout(const pmb_fit_problem *p, FitParams &F, long long &ws_floats, long long &nparam) {
    if (!p) return fit_fail(PMB_E_INVALID, "fit problem is NULL");
    const pmb_net &n = p->net;
    const int L = n.n_linear - 1;
    if (L < 1 || n.n_linear > MAXL) return fit_fail(PMB_E_UNSUPPORTED, "fit: 1..5 hidden layers");
    if (p->N < 1 || p->M < 1) return fit_fail(PMB_E_INVALID, "fit: N, M must be >= 1");
    if (n.dims[0] > 16 || n.dims[L + 1] > 32 || (n.dims[L + 1] & 1)) return fit_fail(PMB_E_UNSUPPORTED, "fit: <= 16 inputs, <= 32 (mean, log-std) outputs");
    memset(&F, 0, sizeof(F));
    F.N = p->N; F.M = p->M; F.L = L;
    int hmax = 64;        // >= 3 * Dout: the NLL stage parks its per-dim terms next to the raw outputs
    long long np = 0, ws = 0;
    auto take = [&](long long nfl) { long long o = ws; ws += (nfl + 31) & ~31LL; return o; };
    for (int l = 0; l <= L + 1; ++l) F.dims[l] = n.dims[l];
    for (int l = 0; l <= L; ++l) {
        if (!n.W[l]) return fit_fail(PMB_E_INVALID, "fit: W is NULL");
        F.W[l] = n.W[l]; F.b[l] = n.b[l];
        F.w_off[l] = np; np += (long long)n.dims[l + 1] * n.dims[l];
        F.b_off[l] = -1;
        if (n.b[l]) { F.b_off[l] = np; np += n.dims[l + 1]; }
        if (l < L) {
            if (n.dims[l + 1] > PMB_MAX_WIDTH) return fit_fail(PMB_E_UNSUPPORTED, "fit: hidden width > 1024");
            if (!p->logit_p[l] || !p->u[l] || !p->hard[l]) return fit_fail(PMB_E_INVALID, "fit: dropout operands are NULL");
            hmax = max(hmax, n.dims[l + 1]);
            F.logit_p[l] = p->logit_p[l]; F.u[l] = p->u[l]; F.hard[l] = p->hard[l];
            F.temp[l] = p->temp[l]; F.reg_scale[l] = p->reg_scale[l]; F.drop_reg[l] = p->drop_reg[l];
            F.p_off[l] = np; np += n.dims[l + 1];
        }
    }
    F.hmax = (hmax + 3) & ~3;
    F.lmax = n.max_log_std; F.reg_weight = p->reg_weight;
    F.Xw = p->Xw; F.Yw = p->Yw;
    // workspace offsets (resolved by the caller): stored as integers in the pointer fields' stead
    long long off_inp[MAXL], off_delta[MAXL], off_gmask[MAXL];
    for (int l = 0; l <= L; ++l) {
        off_inp[l] = take((long long)p->M * n.dims[l]);
        off_delta[l] = take((long long)p->M * n.dims[l + 1]);
        if (l < L) off_gmask[l] = take((long long)p->M * n.dims[l + 1]);
    }
    const long long off_ll = take(p->M);
    for (int l = 0; l <= L; ++l) {
        F.inp[l] = (float *)0 + off_inp[l];
        F.delta[l] = (float *)0 + off_delta[l];
        if (l < L) F.gmask[l] = (float *)0 + off_gmask[l];
    }
    F.ll = (float *)0 + off_ll;
    ws_floats = ws;
    nparam = np;
    return PMB_OK;
}

extern "C" size_t pmb_fit_workspace_bytes(const pmb_fit_problem *p) {
    FitParams F;
    long long ws, np;
    if (fit_layout(p, F, ws, np) != PMB_OK) return 0;
    return (size_t)ws * sizeof(float);
}

extern "C" size_t pmb_fit_param_count(const pmb_fit_problem *p) {
    FitParams F;
    long long ws, np;
    if (fit_layout(p, F, ws, np) != PMB_OK) return 0;
    return (size_t)np;
}

extern "C" int pmb_fit_gradient(const pmb_fit_problem *p, const long long *idx_dev, float *grad_flat, float *loglik_dev,
                                void *workspace, size_t workspace_bytes, void *stream) {
    FitParams F;
    long long wsf, np;
    int rc = fit_layout(p, F, wsf, np);
    if (rc != PMB_OK) return rc;
    if (!idx_dev || !grad_flat || !loglik_dev || !workspace || !p->Xw || !p->Yw) return fit_fail(PMB_E_INVALID, "fit: NULL argument");
    if (workspace_bytes < (size_t)wsf * sizeof(float)) return fit_fail(PMB_E_WORKSPACE, "fit: workspace too small");
    float *ws = (float *)workspace;
    const int L = F.L;
    for (int l = 0; l <= L; ++l) {
        F.inp[l] = ws + (F.inp[l] - (float *)0);
        F.delta[l] = ws + (F.delta[l] - (float *)0);
        if (l < L) {
            F.gmask[l] = ws + (F.gmask[l] - (float *)0);
            F.mask_out[l] = p->mask_out[l];
            F.p_out[l] = p->p_out[l];
        }
    }
    F.ll = ws + (F.ll - (float *)0);
    F.idx = idx_dev;
    F.grad = grad_flat;
    F.loglik = loglik_dev;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = ((size_t)FIT_RB * 16 + (size_t)(2 * L + 2) * FIT_RB * F.hmax) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fit_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fit_fail(PMB_E_CUDA, cudaGetErrorString(e));
    fit_rows_kernel<<<(p->M + FIT_RB - 1) / FIT_RB, FIT_NT, smem, st>>>(F);
    if ((e = cudaGetLastError()) != cudaSuccess) return fit_fail(PMB_E_CUDA, cudaGetErrorString(e));
    // weight / bias gradients of every linear layer over the minibatch rows (one slice: no partial reduction)
    for (int l = 0; l <= L; ++l) {
        e = launch_wgrad(F.delta[l], F.dims[l + 1], F.dims[l + 1], F.inp[l], F.dims[l], F.dims[l], p->M, 1,
                         grad_flat + F.w_off[l], F.b_off[l] >= 0 ? grad_flat + F.b_off[l] : nullptr, 0, st, 2);
        if (e != cudaSuccess) return fit_fail(PMB_E_CUDA, cudaGetErrorString(e));
    }
    fit_finalize_kernel<<<64, 256, 0, st>>>(F);
    if ((e = cudaGetLastError()) != cudaSuccess) return fit_fail(PMB_E_CUDA, cudaGetErrorString(e));
    return PMB_OK;
}
