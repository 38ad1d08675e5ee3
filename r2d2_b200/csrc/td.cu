// K2: fused double-Q n-step TD kernel.
//
// Replaces, in one launch: the double-Q argmax/gather, value-rescaled n-step target, IS-weighted
// MSE, |TD| and the per-sequence mixed priority of the reference learner
// (worker.py:346-359, 268-276, 383-390), plus dLoss/dQ for the backward pass.
//
// Numerics: h / h^-1 are evaluated with exactly the reference's float32 operation sequence using
// non-contracted IEEE ops (__fadd_rn/__fmul_rn/__fsqrt_rn/__fdiv_rn): h^-1 subtracts two nearly
// equal numbers twice, which amplifies rounding to ~3e-5 absolute, so an algebraically equal
// but differently ordered evaluation would eat a third of the 1e-4 parity budget.
//
// Bound: HBM, ~ rows*(3*A*4 + 17) bytes in, rows*(4 + 4*A) out (0.42 MB at rows=2560, A=9):
// launch-latency dominated; one CTA, one warp per sequence, deterministic reductions.
#include "common.cuh"

namespace r2d2 {

__device__ __forceinline__ float sign_f(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// worker.py:383-385   value.sign()*((value.abs()+1).sqrt()-1) + eps*value
__device__ __forceinline__ float value_rescale(float x) {
    const float eps = 1e-3f;
    float d = __fadd_rn(__fsqrt_rn(__fadd_rn(fabsf(x), 1.f)), -1.f);
    return __fadd_rn(__fmul_rn(sign_f(x), d), __fmul_rn(eps, x));
}

// worker.py:387-390   temp = ((1 + 4*eps*(|v|+1+eps)).sqrt() - 1) / (2*eps); sign(v)*(temp^2 - 1)
__device__ __forceinline__ float inverse_value_rescale(float x) {
    const float eps = 1e-3f, four_eps = (float)(4 * 1e-3), two_eps = (float)(2 * 1e-3);
    float a = __fadd_rn(__fadd_rn(fabsf(x), 1.f), eps);
    float s = __fadd_rn(__fmul_rn(a, four_eps), 1.f);
    float temp = __fdiv_rn(__fadd_rn(__fsqrt_rn(s), -1.f), two_eps);
    return __fmul_rn(sign_f(x), __fadd_rn(__fmul_rn(temp, temp), -1.f));
}

constexpr int kTdThreads = 512;

__global__ void __launch_bounds__(kTdThreads) td_loss_kernel(
    const float* __restrict__ q, const float* __restrict__ qn_on, const float* __restrict__ qn_tg,
    const uint8_t* __restrict__ action, const float* __restrict__ R, const float* __restrict__ G,
    const float* __restrict__ isw, const uint8_t* __restrict__ learn, int B, int A, float* __restrict__ td_out,
    float* __restrict__ prio_out, float* __restrict__ loss_sum_out, int32_t* __restrict__ rows_out,
    float* __restrict__ dq_out) {
    extern __shared__ int s_offset[];            // [B+1] exclusive prefix of learning steps
    __shared__ float s_loss[kTdThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    if (warp == 0) {                              // warp-scan the (tiny) length vector
        int running = 0;
        for (int base = 0; base < B; base += 32) {
            int n = base + lane;
            int v = (n < B) ? (int)learn[n] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (n < B) s_offset[n] = running + incl - v;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) { s_offset[B] = running; *rows_out = running; }
    }
    __syncthreads();

    float warp_loss = 0.f;
    for (int n = warp; n < B; n += nwarps) {      // one warp per sequence
        const int off = s_offset[n], l = s_offset[n + 1] - off;
        float mx = 0.f, sm = 0.f;
        for (int t = lane; t < l; t += 32) {
            const int r = off + t;
            const float* qo = qn_on + (size_t)r * A;
            int best = 0;
            float bv = qo[0];
            for (int a = 1; a < A; ++a) {          // first maximum, like torch.argmax
                float v = qo[a];
                if (v > bv) { bv = v; best = a; }
            }
            const float q_tgt = qn_tg[(size_t)r * A + best];
            const float target = value_rescale(__fadd_rn(R[r], __fmul_rn(G[r], inverse_value_rescale(q_tgt))));
            const int act = action[r];
            const float q_a = q[(size_t)r * A + act];
            const float diff = __fadd_rn(q_a, -target);
            const float w = isw[r];
            const float tdv = fabsf(__fadd_rn(target, -q_a));
            td_out[r] = tdv;
            warp_loss += __fmul_rn(w, __fmul_rn(diff, diff));
            mx = fmaxf(mx, tdv);
            sm += tdv;
            if (dq_out) {
                for (int a = 0; a < A; ++a) dq_out[(size_t)r * A + a] = (a == act) ? 2.f * w * diff : 0.f;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            sm += __shfl_xor_sync(0xffffffffu, sm, o);
        }
        // worker.py:273   0.9*max + 0.1*mean over the sequence's learning rows
        if (lane == 0) prio_out[n] = (l > 0) ? __fadd_rn(__fmul_rn(0.9f, mx), __fmul_rn(0.1f, __fdiv_rn(sm, (float)l))) : 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_loss += __shfl_xor_sync(0xffffffffu, warp_loss, o);
    if (lane == 0) s_loss[warp] = warp_loss;
    __syncthreads();
    if (tid == 0) {
        float total = 0.f;
        for (int w = 0; w < nwarps; ++w) total += s_loss[w];
        *loss_sum_out = total;
    }
}

}  // namespace r2d2

using namespace r2d2;

extern "C" int r2d2_td_loss(const float* q, const float* qn_online, const float* qn_target, const uint8_t* action,
                            const float* n_step_reward, const float* n_step_gamma, const float* is_weights,
                            const uint8_t* learning_steps, int B, int A, float* td_out, float* prio_out,
                            float* loss_sum_out, int32_t* rows_out, float* dq_out, void* stream) {
    R2D2_REQUIRE(q && qn_online && qn_target && action && n_step_reward && n_step_gamma && is_weights &&
                     learning_steps && td_out && prio_out && loss_sum_out && rows_out,
                 "null pointer");
    R2D2_REQUIRE(B >= 1 && B <= 8192 && A >= 1 && A <= 64, "bad B/A");
    td_loss_kernel<<<1, kTdThreads, (B + 1) * sizeof(int), as_stream(stream)>>>(
        q, qn_online, qn_target, action, n_step_reward, n_step_gamma, is_weights, learning_steps, B, A, td_out,
        prio_out, loss_sum_out, rows_out, dq_out);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}
