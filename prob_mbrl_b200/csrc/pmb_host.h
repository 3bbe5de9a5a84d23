// Host-side declarations shared by the translation units of libpmb_b200.so.
#pragma once
#include <cuda_runtime.h>
#include "../../include/pmb_b200.h"

namespace pmb {

struct SweepParams;

struct PackJob {
    const float *src;   // source matrix [SR][src_ld]
    float *dst;         // destination [R][C], zero padded
    int R, C, SR, SC, src_ld, transpose;
};
constexpr int MAX_PACK_JOBS = 64;
struct PackJobs {
    int n;
    PackJob job[MAX_PACK_JOBS];
};

cudaError_t launch_pack(const PackJobs &jobs, cudaStream_t stream);
cudaError_t launch_rollout_fwd(const SweepParams &prm, int P, int smem_bytes, cudaStream_t stream);
cudaError_t launch_rollout_bwd(const SweepParams &prm, int P, int smem_bytes, cudaStream_t stream);
cudaError_t launch_wgrad(const float *A, int lda, int M, const float *B, int ldb, int Nc, long long R, int nsplit,
                         float *w_part, float *bias_part, long long part_stride, cudaStream_t stream, int use_umma);
cudaError_t launch_reduce_partials(const float *part, long long n, int nsplit, float *out, cudaStream_t stream);
cudaError_t launch_reward_mm_fwd(const float *rpre, float *rout, const float *z_rr, float *rstat, int N, int H, int G,
                                 int *status, cudaStream_t stream);
cudaError_t launch_reward_mm_bwd(const float *gout, const float *rpre, const float *z_rr, const float *rstat, float *gin,
                                 int N, int H, int G, cudaStream_t stream);
cudaError_t launch_clip_adam(const pmb_adam_tensor *tab, int nt, float max_norm, float lr, float beta1, float beta2,
                             float eps, long long step, long long *step_dev, float *scratch, const int *skip,
                             cudaStream_t stream);

cudaError_t launch_peer_gather(const float *src, float *full, int H, int Nl, int world, int rank, void *const *peer_bufs,
                               unsigned long long *state_dev, cudaStream_t stream);
cudaError_t launch_take_columns(const float *full, float *local, int H, int Ng, int Nl, int off, cudaStream_t stream);

}  // namespace pmb
