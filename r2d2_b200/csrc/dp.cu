// Data-parallel exchange step over NVLink peer memory: the gradient all-reduce of the multi-GPU learner and
// the one-scalar exchange behind the importance weights of a global sampler, as two kernels of our own instead of NCCL calls.
//
// The reference has one learner and no collective (worker.py:363-365 is a local backward + optimizer step); with one learner
// per GPU the flat gradient of loss_sum (4.33 M fp32) has to be summed over ranks between r2d2_net_backward and K5.
//
// Every rank maps every other rank's gradient buffer (symmetric memory: same size, peer pointers, and -- behind an NVSwitch --
// one MULTICAST address that fans a store out to all ranks and reduces a load over all ranks inside the switch).  Rank r owns
// slice r of the range being reduced:
//     entry barrier   every rank's gradients are final                     (flags in peer memory, release/acquire at .sys)
//     reduce          v = multimem.ld_reduce.add(slice r)                   (in-switch sum over all ranks; P2P loads otherwise)
//     broadcast       multimem.st(slice r) = v                              (P2P stores otherwise)
//     exit barrier    every slice has landed everywhere
// so each rank moves 1/N of the bytes and all ranks end up with bit-identical sums.  The local row count rides in a padding
// slot of the buffer (written before the entry barrier, turned into grad_scale = 1/rows and zeroed again after the exit
// barrier).  The dense range (98 % of the bytes) is reduced from a side stream while the conv layers' backward still runs.
// The backward kernels are single-wave grids of equal CTAs, so what costs time is the SLOWEST SM: 16 CTAs of 512 threads next
// to them stretched the concurrent weight-gradient kernel from 98 to 157 us (so did NCCL's kernel: 147 us).  The launch is
// therefore spread thin -- one warp-sized CTA per SM, no shared memory, 8 loads in flight per thread.
//
// Barriers count epochs in device memory (nothing changes on the host between calls), so the launches can be replayed from a
// CUDA graph.  A rank that waits more than a minute for a peer traps instead of hanging the GPU.
#include "common.cuh"

namespace r2d2 {

constexpr int DP_MAX_RANKS = 16;
constexpr int DP_CHANNELS = 2;                     // independent flag sets: two reductions may be in flight (side + main stream)

struct DpShared {                                  // peer-visible control block (symmetric memory, zeroed by the host)
    unsigned int flag[2 * DP_CHANNELS][DP_MAX_RANKS];      // [2 * channel + (0 entry | 1 exit)][writer rank] = epoch
    unsigned int is_flag[DP_MAX_RANKS];
    double is_min[4][DP_MAX_RANKS];                // [epoch & 3][writer rank]: a post may run up to two epochs ahead of the applies
};

struct DpLocal {                                   // this rank's own state (plain device memory)
    unsigned int epoch[DP_CHANNELS];
    unsigned int done[DP_CHANNELS];                // CTAs that have finished their part of the slice
    unsigned int is_epoch;                         // applies done
    unsigned int is_post_epoch;                    // posts done (a learner that samples ahead posts batch i+1 before it applies batch i)
    unsigned int error;
};

struct DpPeers {
    float* grad[DP_MAX_RANKS];                     // this process's mapping of rank i's gradient buffer
    DpShared* ctl[DP_MAX_RANKS];
    float* mc_grad;                                // multicast mapping of the gradient buffers, or nullptr
    int rank, world;
};

struct DpHandle {
    DpPeers peers;
    DpLocal* local;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_f4(float* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_f4(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Spin until *flag has reached `epoch` (wrap-safe); a peer that never arrives is a lost rank: record it and trap.  The bound
// has to cover honest host-side skew between ranks (start-up, a profiler attaching, a checkpoint being written): ~60 s.
constexpr long long kDpTimeoutClocks = 120000000000ll;
__device__ __forceinline__ void dp_wait_flag(const unsigned int* flag, unsigned int epoch, DpLocal* L, unsigned int code) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        if (clock64() - t0 > kDpTimeoutClocks) {
            L->error = code;
            __threadfence_system();
            __trap();
        }
    }
}

template <bool MULTICAST>
__global__ void __launch_bounds__(512) dp_allreduce_kernel(const DpPeers P, DpLocal* __restrict__ L, long long off, long long len,
                                                           int channel, const int* __restrict__ rows_dev, long long rows_slot,
                                                           float* __restrict__ grad_scale) {
    const int tid = threadIdx.x, rank = P.rank, world = P.world;
    const unsigned int ep = *reinterpret_cast<volatile unsigned int*>(&L->epoch[channel]) + 1u;   // stable until the LAST CTA leaves
    // ---- entry barrier: this rank's gradients (earlier kernels of the stream) and its row count are final
    if (blockIdx.x == 0) {
        if (tid == 0 && rows_dev != nullptr) {
            P.grad[rank][rows_slot] = (float)*rows_dev;
            __threadfence_system();
        }
        __syncthreads();
        if (tid < world) st_release_sys(&P.ctl[tid]->flag[2 * channel][rank], ep);
    }
    if (tid < world) dp_wait_flag(&P.ctl[rank]->flag[2 * channel][tid], ep, L, 1u + (unsigned)channel);
    __syncthreads();
    // ---- slice `rank` of [off, off + len): reduce over all ranks, write the sum to all ranks
    const long long n4 = len >> 2;
    const long long per = (n4 + world - 1) / world;
    const long long lo = per * rank, hi = (lo + per < n4) ? lo + per : n4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    constexpr int U = 8;                                                      // loads in flight per thread (NVLink round trip is ~3 us)
    for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + tid; i0 < hi; i0 += U * stride) {
        if (MULTICAST) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i0 + u * stride < hi) v[u] = multimem_ld_reduce_f4(P.mc_grad + off + 4 * (i0 + u * stride));
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i0 + u * stride < hi) multimem_st_f4(P.mc_grad + off + 4 * (i0 + u * stride), v[u]);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = i0 + u * stride;
                if (i >= hi) break;
                const long long e = off + 4 * i;
                float4 s = ld_relaxed_sys_f4(P.grad[0] + e);
                for (int r = 1; r < world; ++r) {                              // fixed order: every run gives the same bits
                    const float4 v = ld_relaxed_sys_f4(P.grad[r] + e);
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                }
                for (int r = 0; r < world; ++r) st_relaxed_sys_f4(P.grad[r] + e, s);
            }
        }
    }
    // ---- exit barrier: the last CTA of this rank tells the peers and waits for theirs
    __syncthreads();
    int mine = 0;
    if (tid == 0) {
        __threadfence_system();
        mine = (atomicAdd(&L->done[channel], 1u) == gridDim.x - 1) ? 1 : 0;
        __threadfence_system();                                                 // the other CTAs' stores (seen through `done`) before our signal
    }
    if (!__syncthreads_or(mine)) return;                                        // (no shared memory: the CTAs must fit next to full-smem GEMM CTAs)
    if (tid < world) {
        st_release_sys(&P.ctl[tid]->flag[2 * channel + 1][rank], ep);
        dp_wait_flag(&P.ctl[rank]->flag[2 * channel + 1][tid], ep, L, 3u + (unsigned)channel);
    }
    __syncthreads();
    if (tid == 0) {
        if (grad_scale != nullptr) {
            volatile float* slot = P.grad[rank] + rows_slot;
            *grad_scale = 1.0f / *slot;                                         // 1 / (rows of the GLOBAL batch)
            *slot = 0.0f;                                                        // padding again before the optimizer's norm
        }
        L->done[channel] = 0u;
        __threadfence();
        *reinterpret_cast<volatile unsigned int*>(&L->epoch[channel]) = ep;
    }
}

// Importance weights of ONE prioritized sampler over all shards (priority_tree.py:39-41 normalises by the batch minimum of
// p_i / root): rank s sampled with its own root and minimum, so its weights are off by ((min_s / root_s) / m)^-beta with m the
// minimum of min_s / root_s over ranks.  Two one-CTA kernels so that no rank ever waits for another one here:
//   post   (after sampling, on a side stream)  local minimum over the sampled leaves -> one double + flag to every rank
//   apply  (right before K2, ~1 ms later: the values have long arrived)  factor -> is_weights scaled in place.
// apply is ordered after this rank's own post by the flag in its own control block, not by the stream.  Posts and applies
// count their epochs separately: a learner that samples batch i+1 while update i runs posts i+1 before it applies i (values
// are kept for four epochs; a rank can be at most two posts ahead of the slowest apply).
__global__ void __launch_bounds__(256) dp_is_post_kernel(const DpPeers P, DpLocal* __restrict__ L, const double* __restrict__ nodes,
                                                         long long leaf_base, const long long* __restrict__ idx, int n) {
    __shared__ double s[8];
    const int tid = threadIdx.x, rank = P.rank, world = P.world;
    double m = 1e300;
    for (int i = tid; i < n; i += blockDim.x) m = fmin(m, nodes[leaf_base + idx[i]]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) s[tid >> 5] = m;
    __syncthreads();
    const unsigned int ep = *reinterpret_cast<volatile unsigned int*>(&L->is_post_epoch) + 1u;
    double local = s[0];
    for (int w = 1; w < 8; ++w) local = fmin(local, s[w]);
    local /= nodes[0];
    if (tid < world) {
        DpShared* peer = P.ctl[tid];
        *reinterpret_cast<volatile double*>(&peer->is_min[ep & 3u][rank]) = local;       // own control block included
        st_release_sys(&peer->is_flag[rank], ep);                                        // release: ordered after the value by this thread
    }
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile unsigned int*>(&L->is_post_epoch) = ep;
}

__global__ void __launch_bounds__(256) dp_is_apply_kernel(const DpPeers P, DpLocal* __restrict__ L, double beta,
                                                          float* __restrict__ is_weights, int rows, float* __restrict__ factor_out) {
    __shared__ double s_factor;
    const int tid = threadIdx.x, rank = P.rank, world = P.world;
    const unsigned int ep = *reinterpret_cast<volatile unsigned int*>(&L->is_epoch) + 1u;
    if (tid < world) dp_wait_flag(&P.ctl[rank]->is_flag[tid], ep, L, 5u);
    __syncthreads();
    if (tid == 0) {
        double g = 1e300;
        for (int r = 0; r < world; ++r) g = fmin(g, *reinterpret_cast<volatile double*>(&P.ctl[rank]->is_min[ep & 3u][r]));
        const double local = *reinterpret_cast<volatile double*>(&P.ctl[rank]->is_min[ep & 3u][rank]);
        s_factor = pow(local / g, -beta);
        if (factor_out != nullptr) *factor_out = (float)s_factor;
    }
    __syncthreads();
    const float f = (float)s_factor;
    for (int i = tid; i < rows; i += blockDim.x) is_weights[i] *= f;
    if (tid == 0) *reinterpret_cast<volatile unsigned int*>(&L->is_epoch) = ep;
}

}  // namespace r2d2

using namespace r2d2;

// These kernels run NEXT to GEMM / conv CTAs that need ~200 KB of shared memory.  With the default carveout preference an SM
// that holds only our (shared-memory-free) CTAs is configured for a large L1, and a 200 KB CTA cannot be placed on it until
// it has drained: measured, the next backward kernel started 58 us late, when the all-reduce CTAs were leaving.  Asking for
// the maximum shared-memory carveout keeps every SM in the configuration its neighbours need.
template <class Kernel>
static cudaError_t prefer_max_shared(Kernel kern, unsigned long long* configured_devices) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (*configured_devices & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) *configured_devices |= bit;
    return e;
}

static cudaError_t dp_configure_kernels() {
    static unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    cudaError_t e = prefer_max_shared(dp_allreduce_kernel<true>, &c0);
    if (e == cudaSuccess) e = prefer_max_shared(dp_allreduce_kernel<false>, &c1);
    if (e == cudaSuccess) e = prefer_max_shared(dp_is_post_kernel, &c2);
    if (e == cudaSuccess) e = prefer_max_shared(dp_is_apply_kernel, &c3);
    return e;
}

extern "C" {

size_t r2d2_dp_ctl_bytes(void) { return (sizeof(DpShared) + 255) / 256 * 256; }

int r2d2_dp_create(int rank, int world, const unsigned long long* grad_ptrs, unsigned long long mc_grad_ptr,
                   const unsigned long long* ctl_ptrs, void** handle) {
    R2D2_REQUIRE(handle && grad_ptrs && ctl_ptrs, "null argument");
    R2D2_REQUIRE(world >= 1 && world <= DP_MAX_RANKS && rank >= 0 && rank < world, "rank / world out of range");
    DpHandle* h = new DpHandle();
    memset(&h->peers, 0, sizeof(h->peers));
    for (int r = 0; r < world; ++r) {
        h->peers.grad[r] = reinterpret_cast<float*>(grad_ptrs[r]);
        h->peers.ctl[r] = reinterpret_cast<DpShared*>(ctl_ptrs[r]);
        if (!h->peers.grad[r] || !h->peers.ctl[r] || (grad_ptrs[r] & 15) || (ctl_ptrs[r] & 15)) {
            delete h;
            set_error("r2d2_dp_create: peer pointer %d is null or not 16-byte aligned", r);
            return R2D2_ERR_ARG;
        }
    }
    h->peers.mc_grad = reinterpret_cast<float*>(mc_grad_ptr);
    h->peers.rank = rank;
    h->peers.world = world;
    cudaError_t e = dp_configure_kernels();
    if (e == cudaSuccess) e = cudaMalloc(&h->local, sizeof(DpLocal));
    if (e == cudaSuccess) e = cudaMemset(h->local, 0, sizeof(DpLocal));
    if (e != cudaSuccess) {
        delete h;
        set_error("r2d2_dp_create: %s", cudaGetErrorString(e));
        return R2D2_ERR_CUDA;
    }
    *handle = h;
    return R2D2_OK;
}

void r2d2_dp_destroy(void* handle) {
    DpHandle* h = static_cast<DpHandle*>(handle);
    if (!h) return;
    cudaFree(h->local);
    delete h;
}

int r2d2_dp_allreduce(void* handle, long long off, long long len, int channel, const int* rows_dev, long long rows_slot,
                      float* grad_scale_dev, int ctas, int threads, int use_multicast, void* stream) {
    DpHandle* h = static_cast<DpHandle*>(handle);
    R2D2_REQUIRE(h, "null handle");
    R2D2_REQUIRE(channel >= 0 && channel < DP_CHANNELS, "channel out of range");
    R2D2_REQUIRE(off >= 0 && len >= 0 && (off & 3) == 0 && (len & 3) == 0, "range must be whole float4s");
    R2D2_REQUIRE((rows_dev == nullptr) == (grad_scale_dev == nullptr), "row count and grad_scale go together");
    R2D2_REQUIRE(rows_dev == nullptr || (rows_slot >= off && rows_slot < off + len), "the row-count slot must lie inside the range");
    R2D2_REQUIRE(ctas >= 1 && ctas <= 1024 && threads >= 32 && threads <= 512 && threads % 32 == 0, "launch shape out of range");
    if (use_multicast && h->peers.mc_grad != nullptr)
        dp_allreduce_kernel<true><<<ctas, threads, 0, as_stream(stream)>>>(h->peers, h->local, off, len, channel, rows_dev, rows_slot, grad_scale_dev);
    else
        dp_allreduce_kernel<false><<<ctas, threads, 0, as_stream(stream)>>>(h->peers, h->local, off, len, channel, rows_dev, rows_slot, grad_scale_dev);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

int r2d2_dp_is_post(void* handle, const double* nodes, long long leaf_base, const long long* idx, int n, void* stream) {
    DpHandle* h = static_cast<DpHandle*>(handle);
    R2D2_REQUIRE(h && nodes && idx, "null argument");
    R2D2_REQUIRE(n >= 1 && leaf_base >= 0, "bad sizes");
    dp_is_post_kernel<<<1, 256, 0, as_stream(stream)>>>(h->peers, h->local, nodes, leaf_base, idx, n);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

int r2d2_dp_is_apply(void* handle, double beta, float* is_weights, int rows, float* factor_out, void* stream) {
    DpHandle* h = static_cast<DpHandle*>(handle);
    R2D2_REQUIRE(h && is_weights, "null argument");
    R2D2_REQUIRE(rows >= 0, "bad sizes");
    dp_is_apply_kernel<<<1, 256, 0, as_stream(stream)>>>(h->peers, h->local, beta, is_weights, rows, factor_out);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

int r2d2_dp_error(void* handle, unsigned int* out) {
    DpHandle* h = static_cast<DpHandle*>(handle);
    R2D2_REQUIRE(h && out, "null argument");
    R2D2_CUDA_CHECK(cudaMemcpy(out, &h->local->error, sizeof(unsigned int), cudaMemcpyDeviceToHost));
    return R2D2_OK;
}

}  // extern "C"
