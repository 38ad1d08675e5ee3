// K3: HBM-resident float64 sum tree (priority_tree.py:4-45 of the reference).
//
// Layout is the reference's: nodes[0] is the root, children of i are 2i+1 / 2i+2, the leaves
// start at 2^(L-1)-1.  Arithmetic is the reference's too -- every ancestor is the float64 sum
// left+right recomputed from its children, the descent compares/subtracts float64 prefix sums --
// so that, on identical leaf contents and identical uniforms, sampled indices are bit-identical.
//
// Bound: HBM/L2 latency.  The whole tree (<= 64 MB at 2^22 leaves) fits B200's 126 MB L2.
// Algorithmic bytes: sample 8*L per draw + 12 out; update 8 + 24*(L-1) per index.
#include <math.h>
#include <stdarg.h>

#include "common.cuh"

namespace r2d2 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace r2d2

struct r2d2_tree {
    int64_t capacity;
    int num_layers;
    int64_t leaf_base;
    int64_t num_nodes;
    double* nodes;      // [num_nodes]
    int* owner;         // [2^(L-1)] duplicate arbitration: highest batch position wins
    int64_t* cur;       // [scratch_n] per-item node cursor for the level sweep
    double* leaf_p;     // [scratch_n] sampled leaf priorities
    unsigned long long* min_bits;  // [1] running min of sampled priorities (as ordered bits)
    int64_t scratch_n;
    float alpha_f32;    // NumPy evaluates f32_array ** 0.9 with the exponent cast to float32
    double beta;
    uint64_t calls;
};

namespace r2d2 {

constexpr int kSmallN = 1024;  // one CTA handles the whole call (the learner's n = 64 case)

// Indices come from the host surface (PriorityTree.update, ReplayBuffer.update_priorities): anything outside the leaf range
// [0, leaf_base] is dropped here instead of becoming an out-of-bounds access (NumPy raises IndexError in the reference).
__device__ __forceinline__ bool slot_in_tree(int64_t slot, int64_t leaf_base) { return slot >= 0 && slot <= leaf_base; }

__device__ __forceinline__ bool keep_slot(int64_t slot, int64_t old_ptr, int64_t cur_ptr, int64_t spb) {
    // worker.py:247-256: drop slots whose block was overwritten since the batch was sampled
    if (old_ptr < 0 || cur_ptr == old_ptr) return true;
    if (cur_ptr > old_ptr) return (slot < old_ptr * spb) || (slot >= cur_ptr * spb);
    return (slot < old_ptr * spb) && (slot >= cur_ptr * spb);
}

__device__ __forceinline__ double leaf_from_td(float td, float alpha) {
    // float32 pow like NumPy: evaluate in double, round once to float32, widen.
    return (double)(float)pow((double)td, (double)alpha);
}

// ---------------------------------------------------------------------------- update, small n
template <bool kFromTd>
__global__ void __launch_bounds__(kSmallN) tree_update_small(double* __restrict__ nodes, int* __restrict__ owner,
                                                             const int64_t* __restrict__ idx,
                                                             const float* __restrict__ td,
                                                             const double* __restrict__ leaf_in, int n,
                                                             int64_t leaf_base, int num_layers, float alpha,
                                                             int64_t old_ptr, int64_t cur_ptr, int64_t spb) {
    const int i = threadIdx.x;
    int64_t slot = -1;
    bool keep = false;
    if (i < n) {
        slot = idx[i];
        keep = slot_in_tree(slot, leaf_base) && keep_slot(slot, old_ptr, cur_ptr, spb);
        if (keep) atomicMax(&owner[slot], i);
    }
    __syncthreads();
    int64_t node = -1;
    if (keep && __ldcg(&owner[slot]) == i) {
        node = leaf_base + slot;
        nodes[node] = kFromTd ? leaf_from_td(td[i], alpha) : leaf_in[i];
    }
    __syncthreads();
    if (node >= 0) owner[slot] = -1;
    for (int level = 0; level < num_layers - 1; ++level) {
        __threadfence_block();
        __syncthreads();
        if (node > 0) {
            node = (node - 1) >> 1;
            // all co-owners of this parent write the same value: benign
            nodes[node] = nodes[2 * node + 1] + nodes[2 * node + 2];
        }
    }
}

// ---------------------------------------------------------------------------- update, large n
__global__ void tree_mark(int* __restrict__ owner, const int64_t* __restrict__ idx, int64_t n, int64_t leaf_base, int64_t old_ptr,
                          int64_t cur_ptr, int64_t spb) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t slot = idx[i];
    if (slot_in_tree(slot, leaf_base) && keep_slot(slot, old_ptr, cur_ptr, spb)) atomicMax(&owner[slot], (int)i);
}

template <bool kFromTd>
__global__ void tree_write_leaves(double* __restrict__ nodes, int* __restrict__ owner, int64_t* __restrict__ cur,
                                  const int64_t* __restrict__ idx, const float* __restrict__ td,
                                  const double* __restrict__ leaf_in, int64_t n, int64_t leaf_base, float alpha,
                                  int64_t old_ptr, int64_t cur_ptr, int64_t spb) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t slot = idx[i];
    int64_t node = -1;
    if (slot_in_tree(slot, leaf_base) && keep_slot(slot, old_ptr, cur_ptr, spb) && __ldcg(&owner[slot]) == (int)i) {
        node = leaf_base + slot;
        nodes[node] = kFromTd ? leaf_from_td(td[i], alpha) : leaf_in[i];
    }
    cur[i] = node;
}

__global__ void tree_reset_owner(int* __restrict__ owner, const int64_t* __restrict__ cur, int64_t n,
                                 int64_t leaf_base) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (cur[i] >= 0) owner[cur[i] - leaf_base] = -1;
}

__global__ void tree_sweep_level(double* __restrict__ nodes, int64_t* __restrict__ cur, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t node = cur[i];
    if (node <= 0) return;
    node = (node - 1) >> 1;
    nodes[node] = nodes[2 * node + 1] + nodes[2 * node + 2];
    cur[i] = node;
}

// ---------------------------------------------------------------------------- sampling
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__device__ __forceinline__ double philox_unit_double(uint64_t seed, uint64_t call, uint64_t i) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    uint64_t bits = (((uint64_t)c[0] << 32) | c[1]) >> 11;  // 53 random bits
    return (double)bits * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ int64_t descend(const double* __restrict__ nodes, int num_layers, double prefix) {
    // priority_tree.py:33-36: go left iff prefix < left child's mass, else subtract it and go right
    int64_t node = 0;
    for (int level = 0; level < num_layers - 1; ++level) {
        const double left = nodes[2 * node + 1];
        if (prefix < left) {
            node = 2 * node + 1;
        } else {
            prefix = prefix - left;
            node = 2 * node + 2;
        }
    }
    return node;
}

__device__ __forceinline__ double stratified_prefix(double total, int64_t n, int64_t i, double r) {
    // priority_tree.py:27-30: np.arange(0, total, w)[i] == i*w ; np.random.uniform(0, w) == w*r
    // explicit _rn ops: an FMA contraction would round differently from NumPy's mul-then-add
    const double w = __ddiv_rn(total, (double)n);
    return __dadd_rn(__dmul_rn((double)i, w), __dmul_rn(w, r));
}

__global__ void __launch_bounds__(kSmallN) tree_sample_small(const double* __restrict__ nodes, int n,
                                                             const double* __restrict__ unit, uint64_t seed,
                                                             uint64_t call, int64_t leaf_base, int num_layers,
                                                             double beta, int64_t* __restrict__ idx_out,
                                                             float* __restrict__ w32, double* __restrict__ w64) {
    __shared__ double s_min[kSmallN / 32];
    const int i = threadIdx.x;
    double p = INFINITY;
    int64_t node = 0;
    if (i < n) {
        const double r = unit ? unit[i] : philox_unit_double(seed, call, i);
        node = descend(nodes, num_layers, stratified_prefix(nodes[0], n, i, r));
        p = nodes[node];
    }
    double m = p;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((i & 31) == 0) s_min[i >> 5] = m;
    __syncthreads();
    if (i < 32) {
        m = (i < (blockDim.x >> 5)) ? s_min[i] : INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (i == 0) s_min[0] = m;
    }
    __syncthreads();
    if (i < n) {
        const double w = pow(p / s_min[0], -beta);   // priority_tree.py:39-41
        idx_out[i] = node - leaf_base;
        w32[i] = (float)w;
        if (w64) w64[i] = w;
    }
}

__global__ void tree_sample_descend(const double* __restrict__ nodes, int64_t n, const double* __restrict__ unit,
                                    uint64_t seed, uint64_t call, int64_t leaf_base, int num_layers,
                                    int64_t* __restrict__ idx_out, double* __restrict__ leaf_p,
                                    unsigned long long* __restrict__ min_bits) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double p = INFINITY;
    if (i < n) {
        const double r = unit ? unit[i] : philox_unit_double(seed, call, i);
        const int64_t node = descend(nodes, num_layers, stratified_prefix(nodes[0], n, i, r));
        p = nodes[node];
        idx_out[i] = node - leaf_base;
        leaf_p[i] = p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p = fmin(p, __shfl_xor_sync(0xffffffffu, p, o));
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0 && p != INFINITY) atomicMin(min_bits, (unsigned long long)__double_as_longlong(p));
}

__global__ void tree_sample_weights(const double* __restrict__ leaf_p, const unsigned long long* __restrict__ min_bits,
                                    int64_t n, double beta, float* __restrict__ w32, double* __restrict__ w64) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double pmin = __longlong_as_double((long long)*min_bits);
    const double w = pow(leaf_p[i] / pmin, -beta);
    w32[i] = (float)w;
    if (w64) w64[i] = w;
}

static int ensure_scratch(r2d2_tree* t, int64_t n) {
    if (n <= t->scratch_n) return R2D2_OK;
    if (t->cur) cudaFree(t->cur);
    if (t->leaf_p) cudaFree(t->leaf_p);
    t->cur = nullptr; t->leaf_p = nullptr; t->scratch_n = 0;
    R2D2_CUDA_CHECK(cudaMalloc(&t->cur, n * sizeof(int64_t)));
    R2D2_CUDA_CHECK(cudaMalloc(&t->leaf_p, n * sizeof(double)));
    t->scratch_n = n;
    return R2D2_OK;
}

template <bool kFromTd>
static int tree_update_impl(r2d2_tree* t, const int64_t* idx, const float* td, const double* leaf, int64_t n,
                            int64_t old_ptr, int64_t cur_ptr, int64_t spb, cudaStream_t s) {
    if (n == 0) return R2D2_OK;   // priority_tree.py handles an empty update as a no-op
    if (n <= kSmallN) {
        const int threads = (int)((n + 31) / 32 * 32);
        tree_update_small<kFromTd><<<1, threads, 0, s>>>(t->nodes, t->owner, idx, td, leaf, (int)n, t->leaf_base,
                                                       t->num_layers, t->alpha_f32, old_ptr, cur_ptr, spb);
        R2D2_LAUNCH_CHECK();
        return R2D2_OK;
    }
    int rc = ensure_scratch(t, n);
    if (rc) return rc;
    const int threads = 256, blocks = cdiv(n, threads);
    tree_mark<<<blocks, threads, 0, s>>>(t->owner, idx, n, t->leaf_base, old_ptr, cur_ptr, spb);
    tree_write_leaves<kFromTd><<<blocks, threads, 0, s>>>(t->nodes, t->owner, t->cur, idx, td, leaf, n, t->leaf_base,
                                                        t->alpha_f32, old_ptr, cur_ptr, spb);
    tree_reset_owner<<<blocks, threads, 0, s>>>(t->owner, t->cur, n, t->leaf_base);
    for (int level = 0; level < t->num_layers - 1; ++level)
        tree_sweep_level<<<blocks, threads, 0, s>>>(t->nodes, t->cur, n);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

}  // namespace r2d2

using namespace r2d2;

extern "C" {

const char* r2d2_last_error(void) { return g_err; }
int r2d2_abi_version(void) { return 1; }

int r2d2_device_ok(void) {
    int dev = 0;
    cudaDeviceProp prop;
    R2D2_CUDA_CHECK(cudaGetDevice(&dev));
    R2D2_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("device %s is sm_%d%d; this library holds sm_100a code only", prop.name, prop.major, prop.minor);
        return R2D2_ERR_STATE;
    }
    return R2D2_OK;
}

int r2d2_tree_create(int64_t capacity, double prio_exponent, double is_exponent, r2d2_tree** out) {
    R2D2_REQUIRE(out != nullptr && capacity >= 1 && capacity <= (1ll << 30), "bad capacity");
    r2d2_tree* t = new r2d2_tree();
    memset(t, 0, sizeof(*t));
    t->capacity = capacity;
    int L = 1;
    while (capacity > (1ll << (L - 1))) ++L;      // priority_tree.py:6-8
    t->num_layers = L;
    t->leaf_base = (1ll << (L - 1)) - 1;
    t->num_nodes = (1ll << L) - 1;
    t->alpha_f32 = (float)prio_exponent;
    t->beta = is_exponent;
    const int64_t leaves = 1ll << (L - 1);
    R2D2_CUDA_CHECK(cudaMalloc(&t->nodes, t->num_nodes * sizeof(double)));
    R2D2_CUDA_CHECK(cudaMalloc(&t->owner, leaves * sizeof(int)));
    R2D2_CUDA_CHECK(cudaMalloc(&t->min_bits, sizeof(unsigned long long)));
    R2D2_CUDA_CHECK(cudaMemset(t->nodes, 0, t->num_nodes * sizeof(double)));
    R2D2_CUDA_CHECK(cudaMemset(t->owner, 0xFF, leaves * sizeof(int)));   // -1
    R2D2_CUDA_CHECK(cudaDeviceSynchronize());
    *out = t;
    return R2D2_OK;
}

int r2d2_tree_destroy(r2d2_tree* t) {
    if (!t) return R2D2_OK;
    cudaFree(t->nodes); cudaFree(t->owner); cudaFree(t->min_bits);
    if (t->cur) cudaFree(t->cur);
    if (t->leaf_p) cudaFree(t->leaf_p);
    delete t;
    return R2D2_OK;
}

int r2d2_tree_num_layers(const r2d2_tree* t) { return t ? t->num_layers : -1; }
int64_t r2d2_tree_num_nodes(const r2d2_tree* t) { return t ? t->num_nodes : -1; }
double* r2d2_tree_nodes(r2d2_tree* t) { return t ? t->nodes : nullptr; }

int r2d2_tree_update(r2d2_tree* t, const int64_t* idx, const float* td, int64_t n, int64_t old_ptr,
                     int64_t cur_ptr, int64_t seq_per_block, void* stream) {
    R2D2_REQUIRE(t && n >= 0 && (n == 0 || (idx && td)), "bad arguments");
    return tree_update_impl<true>(t, idx, td, nullptr, n, old_ptr, cur_ptr, seq_per_block, as_stream(stream));
}

int r2d2_tree_set_leaves(r2d2_tree* t, const int64_t* idx, const double* leaf, int64_t n, void* stream) {
    R2D2_REQUIRE(t && n >= 0 && (n == 0 || (idx && leaf)), "bad arguments");
    return tree_update_impl<false>(t, idx, nullptr, leaf, n, -1, 0, 1, as_stream(stream));
}

int r2d2_tree_sample(r2d2_tree* t, int64_t n, const double* unit_uniforms, uint64_t seed, int64_t* idx_out,
                     float* isw_out_f32, double* isw_out_f64, void* stream) {
    R2D2_REQUIRE(t && n >= 1 && idx_out && isw_out_f32, "bad arguments");
    cudaStream_t s = as_stream(stream);
    const uint64_t call = t->calls++;
    if (n <= kSmallN) {
        const int threads = (int)((n + 31) / 32 * 32);
        tree_sample_small<<<1, threads, 0, s>>>(t->nodes, (int)n, unit_uniforms, seed, call, t->leaf_base,
                                              t->num_layers, t->beta, idx_out, isw_out_f32, isw_out_f64);
        R2D2_LAUNCH_CHECK();
        return R2D2_OK;
    }
    int rc = ensure_scratch(t, n);
    if (rc) return rc;
    const int threads = 256, blocks = cdiv(n, threads);
    R2D2_CUDA_CHECK(cudaMemsetAsync(t->min_bits, 0x7f, sizeof(unsigned long long), s));
    tree_sample_descend<<<blocks, threads, 0, s>>>(t->nodes, n, unit_uniforms, seed, call, t->leaf_base, t->num_layers,
                                                 idx_out, t->leaf_p, t->min_bits);
    tree_sample_weights<<<blocks, threads, 0, s>>>(t->leaf_p, t->min_bits, n, t->beta, isw_out_f32, isw_out_f64);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

}  // extern "C"
