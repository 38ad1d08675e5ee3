// Persistent LSTM recurrence (forward): ALL time steps of BOTH networks in one cooperative launch.
//
// Replaces T per-step GEMM launches of the stepwise path (net.cu: net_recurrence) for batches of <= 64 sequences.
//   grid = 64 CTAs per network; CTA (net, s) owns gate columns [32 s, 32 s + 32) = hidden units [8 s, 8 s + 8)
//   * its W_hh slice [32][512] (bf16 hi+lo, 64 KB) is staged into shared memory ONCE and stays there;
//   * per step, warp w of the 8 producer warps owns k-block w (hidden units [64w, 64w+64)): its lane 0 acquires the
//     counter of that k-block (incremented by the 8 CTAs that produce those units), the warp cp.asyncs the 64 x 64
//     hi+lo tile of h_{t-1} (16 KB) and arms full[w]  ->  96 tcgen05.mma (M=64, N=32, K=16, bf16x3) into a
//     32-column TMEM accumulator, consumed k-block by k-block as they land  ->  epilogue: 128 threads each own
//     (sequence b, 4 hidden units): gates = acc + XP[t], LSTM cell with c and the previous h held in REGISTERS
//     across steps, stores h_t (split), c_t, gates  ->  CTA barrier + one release-add on the counter of its k-block.
//   Fine-grained (per 64-unit) flags let loads and MMAs of early k-blocks overlap the wait for the slowest producer;
//   the two networks use separate counters, so their step latencies overlap across SMs.
// UMMA M=64 accumulator layout (cta_group::1): row m lives in TMEM lane (m & 15) + 32 * (m >> 4), i.e. the first 16
// lanes of each 32-lane quadrant (cute tmem_frg: Shape<(16,4),N> : Stride<(1,32),128>).
#pragma once
#include "umma2.cuh"

namespace r2d2 {

constexpr int REC_H = 512, REC_G4 = 2048, REC_SLICE = 32, REC_CTAS_PER_NET = REC_G4 / REC_SLICE;   // 64
constexpr int REC_KB = REC_H / UM_BK;                                                              // 8 k-blocks of 64
constexpr int REC_A_TILE = 64 * UM_BK * 2;          // 8 KB  (64 rows x 64 k bf16)
constexpr int REC_B_TILE = REC_SLICE * UM_BK * 2;   // 4 KB
constexpr int REC_NACC = 1;                         // TMEM accumulators the MMAs of a step rotate over (1: rotating is slower, measured)
constexpr int REC_TMEM_COLS = 32;                   // >= REC_NACC * 32, power of two
constexpr int REC_SMEM = 2 * REC_KB * (REC_A_TILE + REC_B_TILE) + 1024 + 256;

struct RecFwdParams {
    const bf16* Whi[2]; const bf16* Wlo[2];   // Whh_p [2048][512] (gate-interleaved rows)
    const float* XP[2];                       // [T*B][2048]
    bf16* Hhi[2]; bf16* Hlo[2];               // HsX [(T+1)*B][512], block 0 = h0
    float* Cs[2];                             // [T*B][512]
    float* Gs[2];                             // [T*B][2048] or nullptr
    const float* c0; int ld_c0;               // stored cell state (hidden + H, row stride 2H)
    const int* len;                           // [B] steps each sequence advances
    unsigned int* bar;                        // [2][8] per-network, per-k-block step counters, zero before launch
    int B, T, net_base;                       // net_base: slot of the first CTA group (single-network launches)
    int fast;
    unsigned long long* trace;                // optional [T][8] globaltimer stamps of CTA 0 (debug)
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// gate nonlinearities on the ex2 unit: absolute error ~2e-7 (ex2.approx is 2 ulp), far inside the parity budget
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(-2.f * fabsf(x));               // in (0, 1]: no overflow
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(UM_THREADS, 1) rec_fwd_kernel(const RecFwdParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t sA = raw + pad;                                       // A: [plane][kb] tiles of 8 KB
    const uint32_t sB = sA + 2 * REC_KB * REC_A_TILE;                    // B: [plane][kb] tiles of 4 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * REC_KB * (REC_A_TILE + REC_B_TILE));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + REC_KB + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = blockIdx.x / REC_CTAS_PER_NET, slice = blockIdx.x % REC_CTAS_PER_NET;
    const int net = P.net_base + grp;
    const int B = P.B, T = P.T;
    const int n0 = slice * REC_SLICE;
    const bool want_lo = !P.fast;

    if (tid == 0) {
        for (int kb = 0; kb < REC_KB; ++kb) mbar_init(smem_u32(&bars[kb]), 1);      // armed by the warp that owns k-block kb
        mbar_init(smem_u32(&bars[REC_KB]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == UM_PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(REC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < UM_PRODUCERS / 32) {
        // ---- W_hh slice -> smem once: 32 rows x 64 chunks (x2 planes)
        for (int u = tid; u < REC_SLICE * 64; u += UM_PRODUCERS) {
            const int row = u >> 6, ch = u & 63, kb = ch >> 3, j = ch & 7;
            const uint32_t dst = (uint32_t)(kb * REC_B_TILE + (row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
            const size_t src = (size_t)(n0 + row) * REC_H + ch * 8;
            cp_async16(sB + dst, P.Whi[net] + src, true);
            if (want_lo) cp_async16(sB + REC_KB * REC_B_TILE + dst, P.Wlo[net] + src, true);
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async_smem();

        // ---- epilogue ownership: sequence b = 16*(warp&3) + lane (lanes < 16), 4 hidden units
        const int b = 16 * (warp & 3) + lane;
        const int half = warp >> 2;                       // columns [16*half, 16*half + 16) of the slice
        const bool owner = lane < 16 && b < B;
        const int j0 = slice * 8 + half * 4;              // first of this thread's 4 hidden units
        float c_reg[4] = {0.f, 0.f, 0.f, 0.f}, h_reg[4] = {0.f, 0.f, 0.f, 0.f};
        int my_len = 0;
        if (owner) {
            my_len = P.len[b];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                c_reg[u] = P.c0[(size_t)b * P.ld_c0 + j0 + u];
                h_reg[u] = split_load(P.Hhi[net], P.Hlo[net], (size_t)b * REC_H + j0 + u);
            }
        }
        // A staging plan: warp w stages k-block w (64 rows x 8 chunks x 2 planes): lane -> chunk j = lane & 7, rows (lane >> 3) + 4 i
        const int a_j = lane & 7, a_r0 = lane >> 3;
        unsigned int* my_flag = P.bar + net * REC_KB + warp;                 // producers of hidden units [64 warp, +64)
        for (int t = 0; t < T; ++t) {
            float4 xp[4];
            if (owner) {
                const float4* xsrc = reinterpret_cast<const float4*>(P.XP[net] + ((size_t)t * B + b) * REC_G4 + n0 + 16 * half);
#pragma unroll
                for (int q = 0; q < 4; ++q) xp[q] = __ldg(xsrc + q);
            }
            const bool tr = P.trace && blockIdx.x == 0 && tid == 0;
            if (tr) P.trace[t * 8 + 0] = gtime();
            if (t > 0) {                                   // the 8 CTAs owning this k-block have published h_{t-1}
                if (lane == 0) {
                    const unsigned int target = (unsigned int)(8 * t);
                    for (uint32_t spins = 0; ld_acquire_u32(my_flag) < target; ++spins)
                        if (spins > (1u << 28)) __trap();
                }
                __syncwarp();
            }
            if (tr) P.trace[t * 8 + 1] = gtime();
            {
                const int kb = warp;
                const bf16* hsrc_hi = P.Hhi[net] + (size_t)t * B * REC_H + kb * 64 + a_j * 8;     // HsX block t = h_{t-1}
                const bf16* hsrc_lo = P.Hlo[net] + (size_t)t * B * REC_H + kb * 64 + a_j * 8;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int row = a_r0 + 4 * i;
                    const uint32_t dst = (uint32_t)(kb * REC_A_TILE + (row >> 3) * 1024 + (row & 7) * 128 + ((a_j ^ (row & 7)) << 4));
                    const bool ok = row < B;
                    const size_t src = ok ? (size_t)row * REC_H : 0;
                    cp_async16(sA + dst, hsrc_hi + src, ok);
                    if (want_lo) cp_async16(sA + REC_KB * REC_A_TILE + dst, hsrc_lo + src, ok);
                }
                cp_async_commit();
                cp_async_wait<0>();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[kb]));
            }
            if (tr) P.trace[t * 8 + 2] = gtime();
            // ---- epilogue of step t
            mbar_wait(smem_u32(&bars[REC_KB]), (uint32_t)t & 1u);
            tc_fence_after();
            if (tr) P.trace[t * 8 + 3] = gtime();
            float acc[16];
            tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(16 * half), acc);
#pragma unroll
            for (int a = 1; a < REC_NACC; ++a) {           // the step's MMAs were spread over REC_NACC accumulators
                float part[16];
                tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a * REC_SLICE + 16 * half), part);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] += part[i];
            }
            if (owner) {
                const bool live = t < my_len;
                const size_t row = (size_t)t * B + b;
                float hn4[4], cn4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 x = xp[u];
                    const float gi = fast_sigmoid(acc[4 * u] + x.x);
                    const float gf = fast_sigmoid(acc[4 * u + 1] + x.y);
                    const float gg = fast_tanh(acc[4 * u + 2] + x.z);
                    const float go = fast_sigmoid(acc[4 * u + 3] + x.w);
                    const float cn = gf * c_reg[u] + gi * gg;
                    const float hn = go * fast_tanh(cn);
                    if (P.Gs[net]) *reinterpret_cast<float4*>(P.Gs[net] + row * REC_G4 + n0 + 16 * half + 4 * u) = make_float4(gi, gf, gg, go);
                    cn4[u] = live ? cn : c_reg[u];
                    hn4[u] = live ? hn : h_reg[u];
                    c_reg[u] = cn4[u];
                    h_reg[u] = hn4[u];
                }
                *reinterpret_cast<float4*>(P.Cs[net] + row * REC_H + j0) = make_float4(cn4[0], cn4[1], cn4[2], cn4[3]);
                uint32_t hh[2], ll[2];
                split2(hn4[0], hn4[1], hh[0], ll[0]);
                split2(hn4[2], hn4[3], hh[1], ll[1]);
                const size_t ho = ((size_t)(t + 1) * B + b) * REC_H + j0;           // HsX block t+1
                *reinterpret_cast<uint2*>(P.Hhi[net] + ho) = make_uint2(hh[0], hh[1]);
                *reinterpret_cast<uint2*>(P.Hlo[net] + ho) = make_uint2(ll[0], ll[1]);
                // hi/lo re-rounding: keep the register copy equal to what other CTAs will read
                h_reg[0] = __uint_as_float(hh[0] << 16) + __uint_as_float(ll[0] << 16);
                h_reg[1] = __uint_as_float(hh[0] & 0xFFFF0000u) + __uint_as_float(ll[0] & 0xFFFF0000u);
                h_reg[2] = __uint_as_float(hh[1] << 16) + __uint_as_float(ll[1] << 16);
                h_reg[3] = __uint_as_float(hh[1] & 0xFFFF0000u) + __uint_as_float(ll[1] & 0xFFFF0000u);
            }
            tc_fence_before();
            if (tr) P.trace[t * 8 + 4] = gtime();
            asm volatile("bar.sync 1, %0;" ::"n"(UM_PRODUCERS) : "memory");          // every owner's stores precede ...
            if (tid == 0) red_release_add(P.bar + net * REC_KB + (slice >> 3), 1u);   // ... this gpu-scope release (cumulative)
            if (tr) P.trace[t * 8 + 5] = gtime();
        }
    } else {
        // ---------------------------------------------------------------- MMA issuer
        // The WHOLE warp runs this loop so that the descriptor arithmetic stays warp-uniform (uniform registers feed
        // UTCHMMA directly); only the elected lane issues.  A divergent single-thread loop pays vector->uniform register
        // moves per MMA (~74 clk per instruction measured).
        {
            const uint32_t idesc = umma_idesc_bf16_mn(64, REC_SLICE);
            const bool leader = elect_one();
            // broadcast the bases through a shuffle so the compiler can keep everything derived from them uniform
            const uint32_t uA = __shfl_sync(0xffffffffu, sA, 0), uB = __shfl_sync(0xffffffffu, sB, 0);
            const uint32_t uT = __shfl_sync(0xffffffffu, tmem_base, 0);
            for (int t = 0; t < T; ++t) {
                const uint32_t ph = (uint32_t)t & 1u;
#pragma unroll
                for (int kb = 0; kb < REC_KB; ++kb) {
                    mbar_wait(smem_u32(&bars[kb]), ph);
                    tc_fence_after();
                    const uint64_t a_hi = umma_desc_sw128(uA + kb * REC_A_TILE), a_lo = umma_desc_sw128(uA + (REC_KB + kb) * REC_A_TILE);
                    const uint64_t b_hi = umma_desc_sw128(uB + kb * REC_B_TILE), b_lo = umma_desc_sw128(uB + (REC_KB + kb) * REC_B_TILE);
#pragma unroll
                    for (int k = 0; k < UM_BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);
                        // consecutive MMAs go to different accumulators: a chain on ONE accumulator costs ~70 clk per
                        // instruction (measured), independent ones pipeline at the 16-clk issue rate
                        if (leader) {
                            uint32_t accum = (kb | k) ? 1u : 0u;
                            if (want_lo) {
                                umma_bf16(uT, a_lo + adv, b_hi + adv, idesc, accum);
                                umma_bf16(uT, a_hi + adv, b_lo + adv, idesc, 1u);
                                accum = 1u;
                            }
                            umma_bf16(uT, a_hi + adv, b_hi + adv, idesc, accum);
                        }
                    }
                }
                if (leader) umma_commit(smem_u32(&bars[REC_KB]));
                __syncwarp();
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == UM_PRODUCERS / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(REC_TMEM_COLS) : "memory");
    }
}

static inline cudaError_t launch_rec_fwd(const RecFwdParams& P, int nets, cudaStream_t s) {
    static unsigned long long configured = 0;                       // bit per device ordinal
    {
        cudaError_t e = ensure_dynamic_smem(rec_fwd_kernel, REC_SMEM, &configured);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaMemsetAsync(P.bar, 0, 2 * REC_KB * sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&P};
    return cudaLaunchCooperativeKernel((const void*)rec_fwd_kernel, dim3(nets * REC_CTAS_PER_NET), dim3(UM_THREADS), args, REC_SMEM, s);
}

// ================================================================================================
// Persistent BPTT recurrence (backward): all T steps of  dgates_t -> dh_{t-1} = dgates_t . W_hh  in one
// cooperative launch (online network only).
//
//   grid = 128 CTAs.  CTA c plays two roles every step:
//   * pointwise role for hidden units [4c, 4c+4): thread (b, u) owns (sequence b, unit 4c+u) with dc carried in a
//     register; dh_t = dH[t] + sum of the 8 K-slice partials of dgates_{t+1}.W_hh (fixed order: deterministic),
//     LSTM cell backward -> d(pre-activation gates) DG_t (split) ;
//   * GEMM role (ki = c / 16, ji = c % 16): partial[ki][b][32 ji + n] = DG_t[b][256 ki .. +256) . W_hh[256 ki.., 32 ji + n]
//     with its W_hh^T block [32][256] (bf16 hi+lo, 32 KB) resident in shared memory, UMMA M=64, N=32, 48 MMAs.
//   Dependencies are tracked with release/acquire counters per 64-unit K slice (flagDG[ki], 16 producers a step)
//   and per 32-unit output slice (flagDH[ji], 8 producers a step); partial buffers alternate with the step parity.
// ================================================================================================
constexpr int RB_KB = 4;                               // k-blocks of 64 per K slice of 256
constexpr int RB_A_TILE = 64 * UM_BK * 2;              // 8 KB
constexpr int RB_B_TILE = 32 * UM_BK * 2;              // 4 KB
constexpr int RB_SMEM = 2 * RB_KB * (RB_A_TILE + RB_B_TILE) + 1024 + 256;

struct RecBwdParams {
    const bf16* WThi; const bf16* WTlo;       // WhhT_p [512][2048]: row j, col n' (gate-interleaved)
    const float* dH;                          // [T*B][512]   dLoss/dh_t from the head
    const float* Gs; const float* Cs;         // saved gates [T*B][2048], cell states [T*B][512]
    const float* c0; int ld_c0;               // stored cell state
    const int* len;                           // [B] number of steps carrying gradient (burn-in + learning)
    bf16* DGhi; bf16* DGlo;                   // [T*B][2048] out (split)
    float* partial;                           // [2][8][64][512] fp32 scratch
    unsigned int* flags;                      // [8] flagDG + [16] flagDH, zero before launch
    int B, T, fast;
    unsigned long long* trace = nullptr;      // optional [T][8] globaltimer stamps of CTA 0 (cluster kernel, debug)
    unsigned int* started = nullptr;          // optional counter, +1 when the cluster kernel has begun executing (r2d2_net_shadow_gate)
};

__global__ void __launch_bounds__(UM_THREADS, 1) rec_bwd_kernel(const RecBwdParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t sA = raw + pad;
    const uint32_t sB = sA + 2 * RB_KB * RB_A_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * RB_KB * (RB_A_TILE + RB_B_TILE));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + RB_KB + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.x, ki = c >> 4, ji = c & 15;
    const int B = P.B, T = P.T;
    const bool want_lo = !P.fast;
    unsigned int* flagDG = P.flags;          // [8]
    unsigned int* flagDH = P.flags + 8;      // [16]

    if (tid == 0) {
        for (int kb = 0; kb < RB_KB; ++kb) mbar_init(smem_u32(&bars[kb]), UM_PRODUCERS / 32);
        mbar_init(smem_u32(&bars[RB_KB]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == UM_PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < UM_PRODUCERS / 32) {
        // ---- resident B: W_hh^T block rows j in [32 ji, +32), cols n' in [256 ki, +256)
        for (int u = tid; u < 32 * 32; u += UM_PRODUCERS) {
            const int row = u >> 5, ch = u & 31, kb = ch >> 3, j = ch & 7;
            const uint32_t dst = (uint32_t)(kb * RB_B_TILE + (row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
            const size_t src = (size_t)(32 * ji + row) * REC_G4 + 256 * ki + ch * 8;
            cp_async16(sB + dst, P.WThi + src, true);
            if (want_lo) cp_async16(sB + RB_KB * RB_B_TILE + dst, P.WTlo + src, true);
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async_smem();

        // pointwise ownership: (sequence pb, unit pj)
        const int pb = tid >> 2, pj = 4 * c + (tid & 3);
        const bool p_own = pb < B;
        const int p_len = p_own ? P.len[pb] : 0;
        float dcrec = 0.f;
        // GEMM epilogue ownership (UMMA M=64 layout): row eb, 16 columns
        const int eb = 16 * (warp & 3) + lane, ehalf = warp >> 2;
        const bool e_own = lane < 16 && eb < B;
        const int a_j = tid & 7, a_r0 = tid >> 3;           // A staging: rows a_r0, a_r0 + 32

        for (int t = T - 1, step = 0; t >= 0; --t, ++step) {
            // ================= pointwise role: DG_t for units [4c, 4c+4) =================
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            float dh = 0.f, ct = 0.f, cp = 0.f;
            const bool live = p_own && t < p_len;
            if (live) {                                    // loads that do not depend on the previous step
                const size_t row = (size_t)t * B + pb;
                g = *reinterpret_cast<const float4*>(P.Gs + row * REC_G4 + 4 * pj);
                dh = P.dH[row * REC_H + pj];
                ct = P.Cs[row * REC_H + pj];
                cp = t ? P.Cs[((size_t)(t - 1) * B + pb) * REC_H + pj] : P.c0[(size_t)pb * P.ld_c0 + pj];
            }
            if (step > 0) {                                // partials of dgates_{t+1}.W_hh for my output slice are complete
                if (tid == 0) {
                    const unsigned int target = 8u * (unsigned int)step;
                    for (uint32_t spins = 0; ld_acquire_u32(flagDH + (c >> 3)) < target; ++spins)
                        if (spins > (1u << 28)) __trap();
                }
                asm volatile("bar.sync 1, %0;" ::"n"(UM_PRODUCERS) : "memory");
                if (live) {
                    const float* pp = P.partial + (size_t)((t + 1) & 1) * 8 * 64 * REC_H + (size_t)pb * REC_H + pj;
#pragma unroll
                    for (int q = 0; q < 8; ++q) dh += __ldcg(pp + (size_t)q * 64 * REC_H);
                }
            }
            {
                float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
                if (live) {
                    const float tc = fast_tanh(ct);
                    const float dc = dcrec + dh * g.w * (1.f - tc * tc);
                    o0 = dc * g.z * g.x * (1.f - g.x);
                    o1 = dc * cp * g.y * (1.f - g.y);
                    o2 = dc * g.x * (1.f - g.z * g.z);
                    o3 = dh * tc * g.w * (1.f - g.w);
                    dcrec = dc * g.y;
                }
                if (p_own) {
                    uint32_t h[2], l[2];
                    split2(o0, o1, h[0], l[0]);
                    split2(o2, o3, h[1], l[1]);
                    const size_t o = ((size_t)t * B + pb) * REC_G4 + 4 * pj;
                    *reinterpret_cast<uint2*>(P.DGhi + o) = make_uint2(h[0], h[1]);
                    *reinterpret_cast<uint2*>(P.DGlo + o) = make_uint2(l[0], l[1]);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(UM_PRODUCERS) : "memory");
            if (tid == 0) red_release_add(flagDG + (c >> 4), 1u);
            if (t == 0) break;                              // dh_{-1} is not needed

            // ================= GEMM role: partial[ki] of dgates_t . W_hh for output slice ji =================
            if (tid == 0) {
                const unsigned int target = 16u * (unsigned int)(step + 1);
                for (uint32_t spins = 0; ld_acquire_u32(flagDG + ki) < target; ++spins)
                    if (spins > (1u << 28)) __trap();
            }
            asm volatile("bar.sync 1, %0;" ::"n"(UM_PRODUCERS) : "memory");
            {
                const bf16* src_hi = P.DGhi + (size_t)t * B * REC_G4 + 256 * ki + a_j * 8;
                const bf16* src_lo = P.DGlo + (size_t)t * B * REC_G4 + 256 * ki + a_j * 8;
#pragma unroll
                for (int kb = 0; kb < RB_KB; ++kb) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int row = a_r0 + 32 * r;
                        const uint32_t dst = (uint32_t)(kb * RB_A_TILE + (row >> 3) * 1024 + (row & 7) * 128 + ((a_j ^ (row & 7)) << 4));
                        const bool ok = row < B;
                        const size_t so = ok ? (size_t)row * REC_G4 + kb * 64 : 0;
                        cp_async16(sA + dst, src_hi + so, ok);
                        if (want_lo) cp_async16(sA + RB_KB * RB_A_TILE + dst, src_lo + so, ok);
                    }
                    cp_async_commit();
                }
#pragma unroll
                for (int kb = 0; kb < RB_KB; ++kb) {
                    switch (RB_KB - 1 - kb) {
                        case 3: cp_async_wait<3>(); break; case 2: cp_async_wait<2>(); break;
                        case 1: cp_async_wait<1>(); break; default: cp_async_wait<0>(); break;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bars[kb]));
                }
            }
            mbar_wait(smem_u32(&bars[RB_KB]), (uint32_t)step & 1u);
            tc_fence_after();
            float acc[16];
            tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(16 * ehalf), acc);
            if (e_own) {
                float* dst = P.partial + ((size_t)(t & 1) * 8 + ki) * 64 * REC_H + (size_t)eb * REC_H + 32 * ji + 16 * ehalf;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    __stcg(reinterpret_cast<float4*>(dst) + q, make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]));
            }
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(UM_PRODUCERS) : "memory");
            if (tid == 0) red_release_add(flagDH + ji, 1u);
        }
    } else {
        // ---------------------------------------------------------------- MMA issuer (whole warp, elected lane issues)
        const uint32_t idesc = umma_idesc_bf16_mn(64, 32);
        const bool leader = elect_one();
        const uint32_t uA = __shfl_sync(0xffffffffu, sA, 0), uB = __shfl_sync(0xffffffffu, sB, 0);
        const uint32_t uT = __shfl_sync(0xffffffffu, tmem_base, 0);
        for (int step = 0; step < T - 1; ++step) {
            const uint32_t ph = (uint32_t)step & 1u;
#pragma unroll
            for (int kb = 0; kb < RB_KB; ++kb) {
                mbar_wait(smem_u32(&bars[kb]), ph);
                tc_fence_after();
                const uint64_t a_hi = umma_desc_sw128(uA + kb * RB_A_TILE), a_lo = umma_desc_sw128(uA + (RB_KB + kb) * RB_A_TILE);
                const uint64_t b_hi = umma_desc_sw128(uB + kb * RB_B_TILE), b_lo = umma_desc_sw128(uB + (RB_KB + kb) * RB_B_TILE);
#pragma unroll
                for (int k = 0; k < UM_BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)(k * 32 >> 4);
                    if (leader) {
                        uint32_t accum = (kb | k) ? 1u : 0u;
                        if (want_lo) {
                            umma_bf16(uT, a_lo + adv, b_hi + adv, idesc, accum);
                            umma_bf16(uT, a_hi + adv, b_lo + adv, idesc, 1u);
                            accum = 1u;
                        }
                        umma_bf16(uT, a_hi + adv, b_hi + adv, idesc, accum);
                    }
                }
            }
            if (leader) umma_commit(smem_u32(&bars[RB_KB]));
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == UM_PRODUCERS / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32) : "memory");
    }
}

static inline cudaError_t launch_rec_bwd(const RecBwdParams& P, cudaStream_t s) {
    static unsigned long long configured = 0;                       // bit per device ordinal
    {
        cudaError_t e = ensure_dynamic_smem(rec_bwd_kernel, RB_SMEM, &configured);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaMemsetAsync(P.flags, 0, 24 * sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&P};
    return cudaLaunchCooperativeKernel((const void*)rec_bwd_kernel, dim3(128), dim3(UM_THREADS), args, RB_SMEM, s);
}

}  // namespace r2d2
