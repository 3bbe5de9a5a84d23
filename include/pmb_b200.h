/*
 * pmb_b200.h -- C ABI of the B200-native imagined-rollout library (libpmb_b200.so).
 *
 * The reference (mcgillmrl/prob_mbrl) has NO FFI / operator-plugin interface: its seam for this
 * path is a Python call with duck-typed nn.Module arguments.  This header
 * is therefore the interface a maintainer would bind from the reference's Python (ctypes stub in
 * INTEGRATION.md); every entry point names the reference code it replaces (file:line relative
 * to the reference root).
 *
 * Conventions
 *   - plain C, raw DEVICE pointers (tensor.data_ptr()), fp32 row-major contiguous tensors;
 *   - asynchronous on `stream` (a cudaStream_t passed as void*), no allocation, no host sync;
 *   - every function returns 0 on success, a negative PMB_E_* code otherwise
 *     (pmb_last_error() gives the text); the Python host turns nonzero into RuntimeError, the
 *     reference's error convention for this path (utils/rollout.py:154-157,
 *     algorithms/mc_pilco.py:122-131);
 *   - numerical failure on the device (non positive-definite particle covariance in moment
 *     matching, reference utils/rollout.py:25) is reported through the int32 status word
 *     `status_dev` (device memory): 0 = ok, otherwise 1 + index of the first failing step.
 */
#ifndef PMB_B200_H
#define PMB_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMB_ABI_VERSION 6
#define PMB_MAX_LINEAR 6      /* linear layers per network (hidden + output projection) */
#define PMB_MAX_WIDTH 1024    /* widest hidden layer the fused sweep accepts */
#define PMB_MAX_REWARD_ROWS 16
#define PMB_MAX_PEERS 16         /* GPUs of one node in a peer-memory exchange */ /* rows of the distance map C (2 for the env tip rewards, D for losses.quadratic_*) */
#define PMB_MAX_STATE 16      /* D + U <= 16 */

enum {
    PMB_OK = 0,
    PMB_E_INVALID = -1,     /* bad descriptor (dims, null pointers) */
    PMB_E_UNSUPPORTED = -2, /* shape outside the fused path (width > PMB_MAX_WIDTH, ...) */
    PMB_E_WORKSPACE = -3,   /* workspace too small */
    PMB_E_CUDA = -4         /* a CUDA runtime call failed (see pmb_last_error) */
};

/* One MLP, nn.Linear layout.  Mirrors what BSequential.forward walks
 * (reference models/modules.py:215-232) for the pattern (Linear, ReLU, [B|C]Dropout) x L, Linear. */
typedef struct pmb_net {
    int n_linear;                      /* L hidden layers + 1 */
    int dims[PMB_MAX_LINEAR + 1];      /* dims[0] = inputs, dims[i+1] = outputs of linear i */
    const float *W[PMB_MAX_LINEAR];    /* [dims[i+1]][dims[i]]  (weight of fc{i} / fc_out) */
    const float *b[PMB_MAX_LINEAR];    /* [dims[i+1]] or NULL */
    const float *mask[PMB_MAX_LINEAR]; /* hidden layer i: dropout mask rows [>= N][dims[i+1]] or NULL;
                                          BDropout.noise (modules.py:61) / CDropout.concrete_noise (modules.py:160) */
    float keep[PMB_MAX_LINEAR];        /* divisor after the mask: BDropout p = 1 - rate, CDropout 1 */
    int has_density;                   /* output is (mean, log_std) of a DiagGaussianDensity (densities.py:87-121) */
    const float *z;                    /* density noise [N][out] (PEGASUS: constant over steps) */
    long long z_step_stride;           /* floats between consecutive steps' noise (0 = constant) */
    float max_log_std;                 /* DiagGaussianDensity.max_log_std */
} pmb_net;

/* One rollout problem = the arguments of utils.rollout(states, dynamics, policy, steps, ...)
 * (reference utils/rollout.py:62-79) after the host has read the modules. */
typedef struct pmb_problem {
    int N;                        /* particles on this device */
    int H;                        /* horizon (steps) */
    int D;                        /* state dims */
    int U;                        /* action dims */
    pmb_net pol;                  /* Policy.model   (models/core.py:221-248) */
    pmb_net dyn;                  /* DynamicsModel.model + output_density (models/core.py:265-303) */
    const float *act_scale;       /* [U] Policy.scale */
    const float *act_bias;        /* [U] Policy.bias  */
    const float *mx, *iSx;        /* [D+U] Regressor input scaler (models/core.py:177) */
    const float *my, *Sy;         /* [D]   Regressor output scaler (densities.py:100-107) */
    int rew_rows;                 /* rows of C (2 for the env *Reward modules; D with C = I for a full quadratic form
                                     on the state, reference losses.py:67-75) */
    const float *rew_C;           /* [rew_rows][D]   delta = C s' + c0 (envs/cartpole/env.py:54-75) */
    const float *rew_c0;          /* [rew_rows] */
    const float *rew_Q;           /* [rew_rows][rew_rows] */
    const float *rew_R;           /* [U][U] */
    float rew_scale, rew_offset;  /* r = scale * exp(-cost) + offset */
    int mm_states, mm_rewards;    /* moment matching flags (utils/rollout.py:121-145) */
    int mm_groups;                /* 0/1 = one group; G = independent contiguous blocks of N/G rows */
    const float *z_mm;            /* [>= N][D]  rows 0..N-1 are used, rotated by the step index (rollout.py:53-59) */
    const float *z_rr;            /* [>= N][1] */
    int n_global;                 /* particles of ALL ranks when the matching spans several GPUs (mm_world > 1: the ranks
                                     hold equal contiguous shards, this one the rows mm_rank * N ..); otherwise N */
    int masks_binary;             /* 1 = every dropout mask value is exactly 0 or 1 (true for BDropout noise and for
                                     CDropout's concrete_noise = (b - probs).detach() + probs, which rounds to b exactly
                                     in fp32; reference models/modules.py:61,113-116).  Lets the planner keep the masks
                                     as bit words (wide cluster-resident sweeps); 0 = unknown (those sweeps are not used) */
    /* Moment matching across the GPUs of one node (the reference is single-process): the
     * per-step statistics are exchanged over NVLink peer memory inside the sweeps.  mm_world <= 1: off. */
    int mm_world, mm_rank;
    void *mm_peer_rec[PMB_MAX_PEERS];     /* [rank]: that rank's exchange area for the per-step records of the states
                                             (pmb_mm_exchange_bytes()[0] bytes, zeroed once, entry mm_rank = the own one,
                                             the others mapped with pmb_peer_open) */
    void *mm_peer_gather[PMB_MAX_PEERS];  /* [rank]: ... for the whole-horizon exchange of the rewards ([1] bytes) */
    void *mm_local_state;                 /* local device memory ([2] bytes, zeroed once): exchange epochs */
} pmb_problem;

/* Tunables (0 = library default). */
typedef struct pmb_tuning {
    int particles_per_cta;   /* 1, 2, 4 or 8 */
    int stream_mode;         /* sweep variant: 0 = auto (FFMA2 cluster-resident sweeps for two-hidden-layer nets <= 256
                                wide without moment matching of the states; otherwise 2),
                                1/2 = streaming sweeps (hidden x hidden weights through a TMA + mbarrier ring),
                                3 = cluster-resident sweeps required (all weights in the shared memory of a
                                thread-block cluster; PMB_E_UNSUPPORTED when the problem is outside them),
                                4 = tensor-core cluster sweeps required (tcgen05 3xTF32 hidden x hidden layers,
                                16-CTA cluster per 128-particle tile; PMB_E_UNSUPPORTED when outside them; opt-in:
                                measured slower than the FFMA2 variants at every BASELINE shape),
                                5 = wide cluster-resident sweeps required (two-hidden-layer nets up to 512 wide in the
                                shared memory of a 16-CTA cluster, up to 36 particles per cluster; needs
                                pmb_problem.masks_binary; auto picks them when the nets are too wide for 3) */
    int wgrad_splits;        /* split-K slices of the batched policy weight gradient */
    int reserved[5];         /* reserved[0]: profiling aid, bit mask of phases to run (1 pack, 2 sweep,
                                4 weight gradient); 0 = all.  reserved[1]: ring stages (2..4), 0 = default; with
                                stream_mode 3: bits 0-3 particles per cluster (1..8, 0 = auto), bits 4-7 CTAs per
                                cluster (4 or 8, 0 = 8), bits 8-23: 1 = the two particle tiles of a CTA run
                                unsynchronised, 2 (and 0 = default) = they alternate on the shared-memory-bound phases.
                                reserved[2..3]: low/high half of a device pointer to >= 64 int64 that receives
                                clock64() timeline marks of one step (profiling aid), 0 = off.
                                reserved[4]: hidden x hidden weight gradient on tcgen05 (TF32 x3 split): 0 = auto (when a split-K
                                slice is <= 1024 rows), 1 = always, 2 = never (FFMA2 kernel) */
} pmb_tuning;

int pmb_abi_version(void);
const char *pmb_last_error(void);

/* Validate a descriptor without touching the device: PMB_OK, PMB_E_INVALID or PMB_E_UNSUPPORTED. */
int pmb_check_problem(const pmb_problem *p, const pmb_tuning *tune);

/* Bytes of device workspace pmb_rollout_forward/backward need for `p` (activations kept for the
 * reverse sweep, packed weights, per-layer deltas, split-K partials). */
size_t pmb_workspace_bytes(const pmb_problem *p, const pmb_tuning *tune);

/* What the planner chose for `p` (no device work): which sweep variant pmb_rollout_forward/backward will launch
 * and with which geometry.  Used by hosts for reporting (bench.py's roofline) and by the tests. */
typedef struct pmb_plan_info {
    int variant;             /* 0 = streaming sweeps (TMA weight ring), 1 = cluster-resident FFMA2 sweeps,
                                2 = tensor-core cluster sweeps, 3 = wide cluster-resident FFMA2 sweeps (16-CTA cluster) */
    int ctas;                /* CTAs of one sweep launch */
    int threads_per_cta;
    int cluster_size;        /* CTAs per thread-block cluster (1 for the streaming sweeps) */
    int particles_per_group; /* particles per CTA (streaming) / per cluster (cluster-resident, tensor-core) */
    int smem_fwd_bytes;      /* dynamic shared memory per CTA of the forward / reverse sweep */
    int smem_bwd_bytes;
    int launches_fwd;        /* kernels one pmb_rollout_forward / pmb_rollout_backward call launches */
    int launches_bwd;
} pmb_plan_info;
int pmb_plan_describe(const pmb_problem *p, const pmb_tuning *tune, pmb_plan_info *info);

/* Number of floats of the flat policy gradient: sum over linear layers of W (+ b when present), in
 * policy.parameters() order  W0, b0, W1, b1, ... */
size_t pmb_policy_param_count(const pmb_problem *p);

/* Forward sweep.  Replaces the loop body of utils.rollout (reference utils/rollout.py:93-163):
 * Policy.forward, DynamicsModel.forward, reward_func, optional mm_resample_, for H steps.
 *   x0      [N][D]        initial particles
 *   states  [H+1][N][D]   states[0] = x0
 *   actions [H][N][U]
 *   rewards [H][N]
 * Keeps in `workspace` what pmb_rollout_backward needs. */
int pmb_rollout_forward(const pmb_problem *p, const pmb_tuning *tune, const float *x0,
                        float *states, float *actions, float *rewards,
                        void *workspace, size_t workspace_bytes, int *status_dev, void *stream);

/* Reverse sweep + batched policy weight gradient.  Replaces loss.backward() through the rollout
 * (reference algorithms/mc_pilco.py:197) for arbitrary cotangents on the three outputs:
 *   g_states [H+1][N][D], g_actions [H][N][U], g_rewards [H][N]   (each may be NULL = zeros)
 *   grad_flat [pmb_policy_param_count]   OVERWRITTEN with dL/dtheta_policy (parameters() order)
 *   dx0       [N][D] or NULL             dL/dx0
 *   da_total  [H][N][U] or NULL          TOTAL dL/da_t of every step (direct cotangent + reward + through the
 *                                        dynamics): what a hook on actions[t] sees in the reference; prioritized
 *                                        replay scores initial states by its norm (algorithms/mc_pilco.py:160-188)
 * Must follow a pmb_rollout_forward on the same problem/workspace. */
int pmb_rollout_backward(const pmb_problem *p, const pmb_tuning *tune,
                         const float *states, const float *actions, const float *rewards,
                         const float *g_states, const float *g_actions, const float *g_rewards,
                         float *grad_flat, float *dx0, float *da_total,
                         void *workspace, size_t workspace_bytes, void *stream);

/* One tensor of the optimiser step (all device pointers, `n` floats each). */
typedef struct pmb_adam_tensor {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    long long n;
} pmb_adam_tensor;

/* clip_grad_norm_(params, max_norm) followed by torch.optim.Adam.step()
 * (reference algorithms/mc_pilco.py:209-214; Adam set-up examples/deep_pilco_no_mm.py:163-166).
 *   table_dev   device array of n_tensors descriptors
 *   max_norm    <= 0 disables clipping
 *   step        1-based step count of THIS update (bias correction), used when step_dev is NULL
 *   step_dev    optional device counter: when non-NULL it is incremented on the device and used as
 *               the step count, so the call can be replayed from a CUDA graph
 *   scratch_dev device scratch of >= 1024 floats; scratch_dev[0] receives the total grad norm.
 *   skip_if_nonzero optional device int32 (the status word of pmb_rollout_forward): when it is non-zero at
 *               execution time the whole update is a no-op (parameters, moments and step_dev untouched) --
 *               the reference raises inside rollout and skips the iteration without an update
 *               (algorithms/mc_pilco.py:122-131). */
int pmb_clip_adam_step(const pmb_adam_tensor *table_dev, int n_tensors, float max_norm,
                       float lr, float beta1, float beta2, float eps, long long step,
                       long long *step_dev, float *scratch_dev, const int *skip_if_nonzero, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Dynamics-model fit: one minibatch iteration of utils.train_regressor
 * (reference utils/train_regressor.py:58-165, default branch) for a Regressor = (Linear, ReLU, CDropout) x L,
 * Linear + DiagGaussianDensity (models/core.py:121-187, models/modules.py:73-171, models/densities.py:87-144).
 * The uniform noise and the hard Bernoulli samples of the concrete-dropout layers are inputs, drawn by the host in
 * the reference's order (modules.py:102-118,135-139) so that the RNG stream is the reference's.
 * ------------------------------------------------------------------------------------------- */
typedef struct pmb_fit_problem {
    int N;                                   /* rows of the whitened dataset */
    int M;                                   /* minibatch rows */
    pmb_net net;                             /* dims / W / b / max_log_std; mask, keep, z are unused */
    const float *logit_p[PMB_MAX_LINEAR];    /* [h_l] CDropout.logit_p of hidden layer l */
    const float *u[PMB_MAX_LINEAR];          /* [M][h_l] uniform noise (torch.rand_like, modules.py:137) */
    const float *hard[PMB_MAX_LINEAR];       /* [M][h_l] Bernoulli(probs) sample (modules.py:115) */
    float temp[PMB_MAX_LINEAR];              /* CDropout.temp */
    float reg_scale[PMB_MAX_LINEAR];         /* CDropout.regularizer_scale (modules.py:21-22) */
    float drop_reg[PMB_MAX_LINEAR];          /* CDropout.dropout_regularizer */
    float reg_weight;                        /* train_regressor(reg_weight=) */
    const float *Xw, *Yw;                    /* whitened dataset [N][dims[0]], [N][dims[last] / 2] (train_regressor.py:75-76) */
    float *mask_out[PMB_MAX_LINEAR];         /* optional [M][h_l]: this iteration's concrete masks (CDropout.concrete_noise) */
    float *p_out[PMB_MAX_LINEAR];            /* optional [h_l]: sigmoid(logit_p) as the forward pass saw it (CDropout.p, modules.py:118) */
} pmb_fit_problem;

size_t pmb_fit_workspace_bytes(const pmb_fit_problem *p);
/* floats of the flat gradient, model.parameters() order: fc0.weight, fc0.bias, drop0.logit_p, fc1.weight, ... */
size_t pmb_fit_param_count(const pmb_fit_problem *p);
const char *pmb_fit_last_error(void);
/* Gradient of  loss = -mean_b log N(y_b | mean_b, std_b) + reg_weight * regularization_loss() / N  (train_regressor.py:
 * 122-131) w.r.t. every trainable tensor for the minibatch rows idx_dev[0..M) (int64), and the mean log-likelihood of
 * the minibatch (the progress-bar value, :143).  Follow with pmb_clip_adam_step(max_norm = 0) for optimizer.step(). */
int pmb_fit_gradient(const pmb_fit_problem *p, const long long *idx_dev, float *grad_flat, float *loglik_dev,
                     void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: the one collective of a particle-sharded iteration -- the sum of every rank's flat
 * policy gradient (+ loss) -- over NVLink / NVSwitch peer memory, stream-ordered and capturable in the iteration's CUDA
 * graph (the reference has no distributed code; this replaces the torch.distributed.all_reduce a port would call).
 * One process per GPU: every rank allocates an exchange buffer (pmb_peer_alloc), the 64-byte handles travel over the
 * host-side process group, every rank opens the other ranks' buffers (pmb_peer_open).
 * ------------------------------------------------------------------------------------------- */
const char *pmb_peer_last_error(void);
/* Sizes of the three exchange areas a sharded moment-matched rollout needs (pmb_problem.mm_peer_rec / mm_peer_gather /
 * mm_local_state); out[0..2] bytes.  PMB_E_UNSUPPORTED when the problem is outside the cluster-resident sweeps. */
int pmb_mm_exchange_bytes(const pmb_problem *p, const pmb_tuning *tune, size_t out[3]);
/* Bytes of one rank's exchange buffer for vectors of n floats: [2 parities][world][n] floats + [world] flags. */
size_t pmb_peer_buffer_bytes(long long n, int world);
/* cudaMalloc + zero + cudaIpcGetMemHandle (handle64: 64 bytes out). */
int pmb_peer_alloc(size_t bytes, void **ptr, void *handle64);
/* cudaIpcOpenMemHandle of another rank's buffer (enables peer access lazily) / its close / free of the own buffer. */
int pmb_peer_open(const void *handle64, void **ptr);
int pmb_peer_close(void *ptr);
int pmb_peer_free(void *ptr);
/* dst[i] = sum over ranks of src[i] (i < n), the world slots added in rank order: deterministic and bitwise identical on
 * every rank.  peer_bufs: HOST array of `world` device pointers (entry `rank` = the own buffer, the others as opened
 * with pmb_peer_open, all sized pmb_peer_buffer_bytes(n, world)); state_dev: 3 zero-initialised uint64 in local device
 * memory (exchange epoch and block counters).  src == dst is allowed.  Every rank must issue the same sequence of
 * calls on its buffer set; a rank that never arrives traps the waiting kernel after ~2^28 polls. */
int pmb_peer_allreduce(const float *src, float *dst, long long n, int world, int rank, void *const *peer_bufs,
                       unsigned long long *state_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PMB_B200_H */
