// K5: global-norm gradient clipping + Adam, fused over the flat parameter buffer.
//
// Replaces  nn.utils.clip_grad_norm_(params, 40); Adam(lr=1e-4, eps=1e-3).step()
// (worker.py:289,364-365) with two launches over the 4.33 M-element flat buffers:
//   1. per-CTA partial sums of g^2 (deterministic order),
//   2. every CTA re-reduces the partials (592 floats, L2 resident), derives the clip coefficient
//      and applies torch's single-tensor Adam update (lerp form of the first moment).
// Bound: HBM, 4 reads + 3 writes of 4 B per parameter = 28 B/param -> 121 MB per step.
#include "common.cuh"

namespace r2d2 {

constexpr int kNormBlocks = 148 * 4;
constexpr int kNormThreads = 256;

__global__ void __launch_bounds__(kNormThreads) sumsq_partial_kernel(const float* __restrict__ g, int64_t n,
                                                                   double* __restrict__ partial, int64_t* __restrict__ step_inc = nullptr) {
    if (step_inc != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *step_inc += 1;     // the update count the NEXT launch (Adam) reads
    __shared__ double s[kNormThreads / 32];
    double acc = 0.0;
    const int64_t n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += (double)g[i] * g[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kNormThreads / 32; ++w) t += s[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        const double* __restrict__ partial, const float* __restrict__ grad_scale,
                                                        float max_norm, float lr, float beta1, float beta2, float eps,
                                                        float bc1, float bc2_sqrt, float* __restrict__ norm_out,
                                                        const int64_t* __restrict__ step_dev, const int32_t* __restrict__ rows_dev) {
    __shared__ double s_tot;
    if (step_dev) {                                          // device-resident update count (CUDA-graph replays): same formulas as the host path
        const double st = (double)*step_dev;
        bc1 = (float)(1.0 - pow((double)beta1, st));
        bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, st));
    }
    if (threadIdx.x < 32) {
        double t = 0.0;
        for (int i = threadIdx.x; i < kNormBlocks; i += 32) t += partial[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) s_tot = t;
    }
    __syncthreads();
    const float scale = rows_dev ? __fdiv_rn(1.f, (float)*rows_dev) : (grad_scale ? *grad_scale : 1.f);
    const float total_norm = (float)sqrt(s_tot) * fabsf(scale);
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
    const float coef = fminf(max_norm / (total_norm + 1e-6f), 1.f) * scale;
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total_norm;
    const float step_size = lr / bc1;
    auto adam = [&](float gr, float& pi, float& mi, float& vi) {
        const float gi = gr * coef;
        mi = mi + (gi - mi) * (1.f - beta1);                 // exp_avg.lerp_(grad, 1-beta1)
        vi = vi * beta2 + (1.f - beta2) * gi * gi;           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi = pi - step_size * (mi / denom);
    };
    // 16-byte accesses over the aligned bulk (cudaMalloc'd flat buffers), scalar tail
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                           reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    const int64_t n4 = aligned ? n / 4 : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 g4 = reinterpret_cast<const float4*>(g)[i];
        float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        adam(g4.x, p4.x, m4.x, v4.x); adam(g4.y, p4.y, m4.y, v4.y); adam(g4.z, p4.z, m4.z, v4.z); adam(g4.w, p4.w, m4.w, v4.w);
        reinterpret_cast<float4*>(p)[i] = p4; reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
    }
    for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam(g[i], pi, mi, vi);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

}  // namespace r2d2

using namespace r2d2;

extern "C" {

/* clip_grad_norm_(max_norm) + Adam step on flat buffers (worker.py:364-365).
 *   grad_scale (device float*, may be NULL): gradients are multiplied by it first (1/rows of the
 *   mean loss, possibly after a cross-rank reduction).  partial_ws: device double[592] scratch.
 *   step: 1-based update count (bias correction).  norm_out (device float*, may be NULL): the
 *   pre-clip global gradient norm. */
int r2d2_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                   const float* grad_scale, double* partial_ws, float max_norm, float lr, float beta1, float beta2,
                   float eps, int64_t step, float* norm_out, void* stream) {
    R2D2_REQUIRE(params && grads && exp_avg && exp_avg_sq && partial_ws && n > 0 && step >= 1, "bad arguments");
    cudaStream_t s = as_stream(stream);
    sumsq_partial_kernel<<<kNormBlocks, kNormThreads, 0, s>>>(grads, n, partial_ws);
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    clip_adam_kernel<<<kNormBlocks, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, partial_ws, grad_scale, max_norm, lr,
                                               beta1, beta2, eps, bc1, bc2_sqrt, norm_out, nullptr, nullptr);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

/* The same update with the update count read from DEVICE memory (step_dev: int64, 1-based, already incremented by the
 * caller on the same stream) and, optionally, the gradient scale derived from a device row count (rows_dev: int32,
 * scale = 1 / rows; overrides grad_scale).  Nothing in the argument list changes between updates, so the whole learner
 * update can be captured once in a CUDA graph and replayed. */
static int clip_adam_dev_impl(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* grad_scale, const int32_t* rows_dev, double* partial_ws, float max_norm, float lr, float beta1,
                             float beta2, float eps, int64_t* step_dev, bool increment, float* norm_out, void* stream) {
    R2D2_REQUIRE(params && grads && exp_avg && exp_avg_sq && partial_ws && n > 0 && step_dev, "bad arguments");
    cudaStream_t s = as_stream(stream);
    sumsq_partial_kernel<<<kNormBlocks, kNormThreads, 0, s>>>(grads, n, partial_ws, increment ? step_dev : nullptr);
    clip_adam_kernel<<<kNormBlocks, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n, partial_ws, grad_scale, max_norm, lr,
                                               beta1, beta2, eps, 1.f, 1.f, norm_out, step_dev, rows_dev);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

int r2d2_clip_adam_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const float* grad_scale, const int32_t* rows_dev, double* partial_ws, float max_norm, float lr, float beta1,
                       float beta2, float eps, const int64_t* step_dev, float* norm_out, void* stream) {
    return clip_adam_dev_impl(params, grads, exp_avg, exp_avg_sq, n, grad_scale, rows_dev, partial_ws, max_norm, lr, beta1, beta2, eps,
                              const_cast<int64_t*>(step_dev), false, norm_out, stream);
}

/* The same, and the update count is incremented on the device first (by the norm kernel, before the Adam kernel reads it):
 * the caller needs no launch of its own to keep the count. */
int r2d2_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* grad_scale, const int32_t* rows_dev, double* partial_ws, float max_norm, float lr, float beta1,
                        float beta2, float eps, int64_t* step_dev, float* norm_out, void* stream) {
    return clip_adam_dev_impl(params, grads, exp_avg, exp_avg_sq, n, grad_scale, rows_dev, partial_ws, max_norm, lr, beta1, beta2, eps,
                              step_dev, true, norm_out, stream);
}

}  // extern "C"
