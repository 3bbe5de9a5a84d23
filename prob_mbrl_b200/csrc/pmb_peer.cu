// The one collective of a sharded iteration -- the sum of every rank's flat policy gradient (+ loss) -- over NVLink /
// NVSwitch peer memory, without NCCL and without leaving the iteration's CUDA graph.
//
// Every rank owns an exchange buffer [2 parities][world][n] floats + [world] 64-bit flags that every other rank of the
// node has mapped (CUDA IPC).  push: a rank stores its vector into slot [epoch & 1][rank] of EVERY rank's buffer (P2P
// stores), fences at system scope and raises flag[rank] = epoch on every rank.  pull: a rank waits until all `world`
// flags of its own buffer reached the epoch and adds the slots in rank order -- deterministic, and bitwise identical on
// every rank (so clip + Adam stay identical without a broadcast).  Slots are double-buffered by epoch parity: a peer can
// be at most one exchange ahead (it needs this rank's next vector to go further).  The writer never waits, the reader
// only waits for writers, so the exchange cannot deadlock whatever the launch skew between the processes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/pmb_b200.h"
#include "pmb_host.h"

namespace pmb {
const char *peer_err = "";
}

namespace {

constexpr int PEER_MAX = 16;

struct PeerPtrs {
    float *buf[PEER_MAX];
    unsigned long long *flag[PEER_MAX];
};

// state[0] = epoch of the last completed exchange, state[1] / state[2] = finished-block counters of push / pull
__global__ void __launch_bounds__(256) peer_push_kernel(const float *__restrict__ src, long long n, int world, int rank, PeerPtrs pp,
                                                        unsigned long long *state) {
    const unsigned long long e = state[0] + 1ull;
    const long long slot = ((long long)(e & 1ull) * world + rank) * n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = src[i];
        for (int p = 0; p < world; ++p) pp.buf[p][slot + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(&state[1], 1ull) + 1ull;
        if (done == gridDim.x) {            // last block: every store of this rank is ordered before the flags
            state[1] = 0ull;
            __threadfence_system();
            for (int p = 0; p < world; ++p)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pp.flag[p] + rank), "l"(e) : "memory");
        }
    }
}

__global__ void __launch_bounds__(256) peer_pull_kernel(float *__restrict__ dst, long long n, int world, int rank, PeerPtrs pp,
                                                        unsigned long long *state) {
    const unsigned long long e = state[0] + 1ull;
    if (threadIdx.x == 0) {
        for (int r = 0; r < world; ++r) {
            unsigned long long v;
            unsigned long long spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(pp.flag[rank] + r) : "memory");
                if (++spins > (1ull << 28)) __trap();      // a rank that never arrives must not hang the GPU forever (~minutes)
            } while (v < e);
        }
    }
    __syncthreads();
    const float *base = pp.buf[rank] + (long long)(e & 1ull) * world * n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float a = __ldcg(base + i);
        for (int r = 1; r < world; ++r) a += __ldcg(base + (long long)r * n + i);
        dst[i] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(&state[2], 1ull) + 1ull;
        if (done == gridDim.x) {
            state[2] = 0ull;
            state[0] = e;
        }
    }
}

// all-gather flavour of pull: wait for every rank's slot, then lay the blocks out as [H][world * Nl] (rank r's block is
// [H][Nl]: its particles of every step)
__global__ void __launch_bounds__(256) peer_gather_kernel(float *__restrict__ dst, int H, int Nl, int world, int rank, PeerPtrs pp,
                                                          unsigned long long *state) {
    const unsigned long long e = state[0] + 1ull;
    if (threadIdx.x == 0) {
        for (int r = 0; r < world; ++r) {
            unsigned long long v, spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(pp.flag[rank] + r) : "memory");
                if (++spins > (1ull << 28)) __trap();
            } while (v < e);
        }
    }
    __syncthreads();
    const long long n = (long long)H * Nl;
    const float *base = pp.buf[rank] + (long long)(e & 1ull) * world * n;
    const long long total = n * world, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int r = (int)(i / n);
        const long long j = i - (long long)r * n;
        const int t = (int)(j / Nl), c = (int)(j - (long long)t * Nl);
        dst[(long long)t * world * Nl + (long long)r * Nl + c] = __ldcg(base + i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(&state[2], 1ull) + 1ull;
        if (done == gridDim.x) {
            state[2] = 0ull;
            state[0] = e;
        }
    }
}

__global__ void __launch_bounds__(256) take_columns_kernel(const float *__restrict__ full, float *__restrict__ local, int H, int Ng,
                                                           int Nl, int off) {
    const long long total = (long long)H * Nl;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / Nl), c = (int)(i - (long long)t * Nl);
        local[i] = full[(long long)t * Ng + off + c];
    }
}

static void fill_ptrs(PeerPtrs &pp, void *const *peer_bufs, long long n, int world) {
    memset(&pp, 0, sizeof(pp));
    const size_t flag_off = (((size_t)(2 * (long long)world * n) * sizeof(float)) + 255) & ~(size_t)255;
    for (int p = 0; p < world; ++p) {
        pp.buf[p] = (float *)peer_bufs[p];
        pp.flag[p] = (unsigned long long *)((char *)peer_bufs[p] + flag_off);
    }
}

}  // namespace

namespace pmb {

// all-gather of every rank's [H][Nl] block over peer memory into full [H][world * Nl] (whole-horizon reward matching of
// a sharded rollout); peer_bufs sized pmb_peer_buffer_bytes(H * Nl, world)
cudaError_t launch_peer_gather(const float *src, float *full, int H, int Nl, int world, int rank, void *const *peer_bufs,
                               unsigned long long *state_dev, cudaStream_t stream) {
    PeerPtrs pp;
    const long long n = (long long)H * Nl;
    fill_ptrs(pp, peer_bufs, n, world);
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148) blocks = 148;
    peer_push_kernel<<<blocks, 256, 0, stream>>>(src, n, world, rank, pp, state_dev);
    peer_gather_kernel<<<blocks, 256, 0, stream>>>(full, H, Nl, world, rank, pp, state_dev);
    return cudaGetLastError();
}
cudaError_t launch_take_columns(const float *full, float *local, int H, int Ng, int Nl, int off, cudaStream_t stream) {
    int blocks = (int)(((long long)H * Nl + 255) / 256);
    if (blocks > 148) blocks = 148;
    take_columns_kernel<<<blocks, 256, 0, stream>>>(full, local, H, Ng, Nl, off);
    return cudaGetLastError();
}

}  // namespace pmb

extern "C" {

const char *pmb_peer_last_error(void) { return pmb::peer_err; }

size_t pmb_peer_buffer_bytes(long long n, int world) {
    if (n < 1 || world < 1 || world > PEER_MAX) return 0;
    return (size_t)(2 * (long long)world * n) * sizeof(float) + (size_t)world * sizeof(unsigned long long) + 256;
}

int pmb_peer_alloc(size_t bytes, void **ptr, void *handle64) {
    if (!ptr || !handle64 || bytes == 0) { pmb::peer_err = "bad arguments"; return PMB_E_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) { pmb::peer_err = cudaGetErrorString(e); return PMB_E_CUDA; }
    memcpy(handle64, &h, 64);
    return PMB_OK;
}

int pmb_peer_open(const void *handle64, void **ptr) {
    if (!ptr || !handle64) { pmb::peer_err = "bad arguments"; return PMB_E_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { pmb::peer_err = cudaGetErrorString(e); return PMB_E_CUDA; }
    return PMB_OK;
}

int pmb_peer_close(void *ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) { pmb::peer_err = cudaGetErrorString(e); return PMB_E_CUDA; }
    return PMB_OK;
}

int pmb_peer_free(void *ptr) {
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) { pmb::peer_err = cudaGetErrorString(e); return PMB_E_CUDA; }
    return PMB_OK;
}

int pmb_peer_allreduce(const float *src, float *dst, long long n, int world, int rank, void *const *peer_bufs,
                       unsigned long long *state_dev, void *stream) {
    if (!src || !dst || !peer_bufs || !state_dev || n < 1 || world < 1 || world > PEER_MAX || rank < 0 || rank >= world) {
        pmb::peer_err = "bad arguments";
        return PMB_E_INVALID;
    }
    for (int p = 0; p < world; ++p)
        if (!peer_bufs[p]) { pmb::peer_err = "NULL peer buffer"; return PMB_E_INVALID; }
    PeerPtrs pp;
    fill_ptrs(pp, peer_bufs, n, world);
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148) blocks = 148;
    cudaStream_t st = (cudaStream_t)stream;
    peer_push_kernel<<<blocks, 256, 0, st>>>(src, n, world, rank, pp, state_dev);
    peer_pull_kernel<<<blocks, 256, 0, st>>>(dst, n, world, rank, pp, state_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pmb::peer_err = cudaGetErrorString(e); return PMB_E_CUDA; }
    return PMB_OK;
}

}  // extern "C"
