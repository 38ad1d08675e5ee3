// "Window" convolution on tcgen05: stride-1 KH x KW convolutions over a dense pixel grid as sums of per-tap GEMMs
// that all read ONE staged activation window.
//
// The im2col formulation (umma2.cuh + SrcConvK) fetches every activation KH*KW times from L2 -- the conv kernels run
// at the L2->SM bandwidth ceiling (7.9-8.7 TB/s measured) with the tensor pipe 13-20 % busy.  Here a CTA stages, per
// 64-channel block, the rows [p0, p0 + 128 + halo) of the activation matrix X[pixel][channel] ONCE (contiguous rows,
// 128-byte swizzle on absolute smem address bits) and issues, for tap (ky,kx), MMAs whose A descriptor simply starts
// ky*GW + kx rows further down the same buffer (a row-shifted descriptor reads the right rows: tools/shift_probe.py).
// Outputs are produced for every pixel of the input grid; pixels whose window leaves the frame are junk and dropped
// by the epilogue (GW x GH grid -> OW x OH valid outputs).  The weights of all taps stay resident in shared memory
// (persistent CTAs loop over pixel tiles), so L2 traffic per 128-pixel tile is one window instead of KH*KW tiles
// plus the weight matrix.
//
// Warp roles (288 threads): warps 0-3 producers (cp.async of the window rows), warps 4-7 epilogue (TMEM -> bias/ReLU
// -> split store), warp 8 MMA issue.  Two TMEM accumulators alternate between consecutive tiles so that the epilogue
// of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "umma2.cuh"

namespace r2d2 {

template <int GW, int IC, int KH, int KW, int N, bool A_LO, bool B_LO>
struct WinCfg {
    static constexpr int kTaps = KH * KW;
    static constexpr int kKB = IC / 64;                               // 64-channel blocks per pixel
    static constexpr int kHalo = (KH - 1) * GW + (KW - 1);
    static constexpr int kWinRows = (128 + kHalo + 7) / 8 * 8;
    static constexpr int kWinBytes = kWinRows * 128;                  // one plane of one window
    static constexpr int kAStage = (A_LO ? 2 : 1) * kWinBytes;
    static constexpr int kBTile = N * 128;                            // [N][64] bf16
    static constexpr int kBBytes = (B_LO ? 2 : 1) * kTaps * kKB * kBTile;
    static constexpr int kStages = ((227 * 1024 - 2048 - kBBytes) / kAStage) >= 4 ? 4 : ((227 * 1024 - 2048 - kBBytes) / kAStage);
    static constexpr int kSmem = kBBytes + kStages * kAStage + 1024 + 256;
    static_assert(IC % 64 == 0 && N % 16 == 0 && N <= 128, "window conv shape");
    static_assert(kStages >= 2, "not enough shared memory for two window stages");
};

// Epilogue contract: store_row(p, acc-chunk) style functor with  void store16(long long p, int n, const float (&v)[16]) const
// where p is the GRID pixel index (frame * GW*GH + gy * GW + gx); the functor drops junk pixels itself.
template <int GW, int IC, int KH, int KW, int N, bool A_LO, bool B_LO, class Epi>
__global__ void __launch_bounds__(UM_THREADS, 1)
winconv_kernel(const bf16* __restrict__ Xhi, const bf16* __restrict__ Xlo, long long R /* total grid pixels */,
               const bf16* __restrict__ Whi, const bf16* __restrict__ Wlo /* [N][taps*IC], k = tap*IC + c */, const Epi ep) {
    using Cfg = WinCfg<GW, IC, KH, KW, N, A_LO, B_LO>;
    constexpr int S = Cfg::kStages, KB = Cfg::kKB, TAPS = Cfg::kTaps, KTOT = TAPS * IC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t sB = raw + pad;                                     // resident weights: [plane][tap][kb] tiles
    const uint32_t sA = sB + Cfg::kBBytes;                             // window ring: [stage][plane]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBBytes + S * Cfg::kAStage);
    // bars: full[S] | empty[S] | accf[2] | acce[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (R + 127) / 128;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(smem_u32(&bars[s]), 4); mbar_init(smem_u32(&bars[S + s]), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bars[2 * S + a]), 1); mbar_init(smem_u32(&bars[2 * S + 2 + a]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * (N < 32 ? 32 : N)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // resident weights (all threads help)
    for (int u = tid; u < TAPS * KB * N * 8; u += UM_THREADS) {
        const int j = u & 7, row = (u >> 3) % N, tile = (u >> 3) / N;          // tile = tap*KB + kb
        const uint32_t dst = (uint32_t)(tile * Cfg::kBTile + (row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
        const size_t src = (size_t)row * KTOT + (size_t)tile * 64 + j * 8;      // k = tap*IC + kb*64 + ...
        cp_async16(sB + dst, Whi + src, true);
        if (B_LO) cp_async16(sB + TAPS * KB * Cfg::kBTile + dst, Wlo + src, true);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr int ACC_COLS = N < 32 ? 32 : N;

    if (warp < 4) {
        // ------------------------------------------------------------------ producers: window rows -> smem
        long long it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long p0 = tile * 128;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = (int)(it % S);
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(smem_u32(&bars[S + s]), ph ^ 1u);
                const uint32_t st = sA + s * Cfg::kAStage;
                for (int u = tid; u < Cfg::kWinRows * 8; u += 128) {
                    const int row = u >> 3, j = u & 7;
                    const long long p = p0 + row;
                    const bool ok = p < R;
                    const size_t src = ok ? (size_t)p * IC + kb * 64 + j * 8 : 0;
                    const uint32_t dst = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
                    cp_async16(st + dst, Xhi + src, ok);
                    if (A_LO) cp_async16(st + Cfg::kWinBytes + dst, Xlo + src, ok);
                }
                cp_async_commit();
                cp_async_wait<0>();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[s]));
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------------ epilogue
        long long ti = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int a = (int)(ti & 1);
            mbar_wait(smem_u32(&bars[2 * S + a]), (uint32_t)(ti >> 1) & 1u);
            tc_fence_after();
            const long long p = tile * 128 + (warp & 3) * 32 + lane;
#pragma unroll
            for (int c = 0; c < N; c += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a * ACC_COLS + c), v);
                if (p < R) ep.store16(p, c, v);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[2 * S + 2 + a]));
        }
    } else {
        // ------------------------------------------------------------------ MMA issue (whole warp, elected lane)
        constexpr uint32_t idesc = umma_idesc_bf16(N);
        const bool leader = elect_one();
        const uint32_t uA = __shfl_sync(0xffffffffu, sA, 0), uB = __shfl_sync(0xffffffffu, sB, 0);
        const uint32_t uT = __shfl_sync(0xffffffffu, tmem_base, 0);
        long long it = 0, ti = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int a = (int)(ti & 1);
            mbar_wait(smem_u32(&bars[2 * S + 2 + a]), ((uint32_t)(ti >> 1) & 1u) ^ 1u);       // accumulator drained
            tc_fence_after();
            const uint32_t acc = uT + a * ACC_COLS;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = (int)(it % S);
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                mbar_wait(smem_u32(&bars[s]), ph);
                tc_fence_after();
                const uint32_t st = uA + s * Cfg::kAStage;
                if (leader) {
#pragma unroll
                    for (int t = 0; t < TAPS; ++t) {
                        const uint32_t shift = (uint32_t)((t / KW) * GW + (t % KW)) * 128u;          // row-shifted view of the window
                        const uint64_t a_hi = umma_desc_sw128(st + shift), a_lo = umma_desc_sw128(st + Cfg::kWinBytes + shift);
                        const uint32_t bt = uB + (t * KB + kb) * Cfg::kBTile;
                        const uint64_t b_hi = umma_desc_sw128(bt), b_lo = umma_desc_sw128(bt + TAPS * KB * Cfg::kBTile);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)(k * 2);
                            uint32_t accum = (kb | t | k) ? 1u : 0u;
                            if (A_LO) { umma_bf16(acc, a_lo + adv, b_hi + adv, idesc, accum); accum = 1u; }
                            if (B_LO) { umma_bf16(acc, a_hi + adv, b_lo + adv, idesc, accum); accum = 1u; }
                            umma_bf16(acc, a_hi + adv, b_hi + adv, idesc, accum);
                        }
                    }
                    umma_commit(smem_u32(&bars[S + s]));
                    if (kb == KB - 1) umma_commit(smem_u32(&bars[2 * S + a]));
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * ACC_COLS) : "memory");
    }
}

template <int GW, int IC, int KH, int KW, int N, bool A_LO, bool B_LO, class Epi>
static inline cudaError_t launch_winconv_inst(SplitC X, long long R, SplitC W, const Epi& ep, cudaStream_t s) {
    using Cfg = WinCfg<GW, IC, KH, KW, N, A_LO, B_LO>;
    auto kern = winconv_kernel<GW, IC, KH, KW, N, A_LO, B_LO, Epi>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const long long ntiles = (R + 127) / 128;
    const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    kern<<<grid, UM_THREADS, Cfg::kSmem, s>>>(X.hi, X.lo, R, W.hi, W.lo, ep);
    return cudaGetLastError();
}

// A_HAS_LO: whether X has a lo plane at all (u8 frames do not).  Precision policy as in launch_umma2 (weights are B).
template <int GW, int IC, int KH, int KW, int N, bool A_HAS_LO, class Epi>
static inline cudaError_t launch_winconv(SplitC X, long long R, SplitC W, const Epi& ep, cudaStream_t s) {
    if (g_fast_math == 1) return launch_winconv_inst<GW, IC, KH, KW, N, false, false>(X, R, W, ep, s);
    if constexpr (!A_HAS_LO) {
        return launch_winconv_inst<GW, IC, KH, KW, N, false, true>(X, R, W, ep, s);
    } else {
        if (g_fast_math == 2) return launch_winconv_inst<GW, IC, KH, KW, N, false, true>(X, R, W, ep, s);
        return launch_winconv_inst<GW, IC, KH, KW, N, true, true>(X, R, W, ep, s);
    }
}

}  // namespace r2d2
