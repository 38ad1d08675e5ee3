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
// Warp roles: warp 0 = producer (one lane issues ONE TMA load per window plane: a 2-D box of [window rows][64 channels] with
// the 128-byte swizzle; rows outside the activation matrix are zero-filled by the TMA unit), 4 or 8 epilogue warps (TMEM ->
// bias/ReLU -> split store), one MMA-issue warp.  Two TMEM accumulators alternate between consecutive tiles so that the epilogue
// of tile i overlaps the MMAs of tile i+1.
#pragma once
#include "umma3.cuh"

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
    static_assert(IC % 64 == 0 && (N == 32 || N == 64 || N == 128), "window conv shape");
    static_assert(kStages >= 2, "not enough shared memory for two window stages");
};

// Warp roles: warp 0 producer (TMA; round 1 needed four cp.async warps here and ran the LSU/L1 at 56-67 % of peak), then
// EW = 4 or 8 epilogue warps (warp & 3 = TMEM lane quadrant; with 8, two warps per quadrant split the columns), then the
// MMA warp.  EW = 8 is for epilogue-bound layers: conv1 ran 1,170 instructions per tile on ONE epilogue warp per
// scheduler at IPC 0.44 (ncu) -- 194 -> 156 us with two; the MMA-bound layers are faster with four (measured).
constexpr int WC_PRODUCERS = 32;
// Epilogue contract: struct Pre; void prefetch(long long p, int col0, Pre&) const  (operands of columns [col0, col0 + N/2));
// void store16(long long p, int n, const float (&v)[16], int col0, const Pre&) const
// where p is the GRID pixel index (frame * GW*GH + gy * GW + gx); the functor drops junk pixels itself.
template <int GW, int IC, int KH, int KW, int N, bool A_LO, bool B_LO, bool BACK, int EW, class Epi>
__global__ void __launch_bounds__(WC_PRODUCERS + 32 * EW + 32, 1)
winconv_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl /* X[pixel][channel] planes, box {64, window rows} */,
               long long R /* total grid pixels */, const bf16* __restrict__ Whi, const bf16* __restrict__ Wlo /* [N][taps*IC], k = tap*IC + c */,
               const Epi ep) {
    using Cfg = WinCfg<GW, IC, KH, KW, N, A_LO, B_LO>;
    constexpr int S = Cfg::kStages, KB = Cfg::kKB, TAPS = Cfg::kTaps, KTOT = TAPS * IC;
    constexpr int WC_THREADS = WC_PRODUCERS + 32 * EW + 32, WC_MMA_WARP = WC_PRODUCERS / 32 + EW;
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
        for (int s = 0; s < S; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[S + s]), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bars[2 * S + a]), 1); mbar_init(smem_u32(&bars[2 * S + 2 + a]), EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // FOLD (N <= 64, weights split): the lo plane of a weight tile sits right behind its hi plane, so ONE instruction with
    // N' = 2N computes  a_hi.b_hi | a_hi.b_lo  side by side; the epilogue adds the halves.  An SS-mode instruction costs
    // max(128 N/256, (128 + N) / 4) clk (tools/mma_rate.py): 64 clk for the folded 128 x 128 x 16 against 2 x 48 for two
    // 128 x 64 x 16.  Rotating accumulators instead buys nothing (same probe).  x2: tile double buffer.
    constexpr bool FOLD = B_LO && N <= 64;
    constexpr int ACC_COLS = (FOLD ? 2 * N : N) < 32 ? 32 : (FOLD ? 2 * N : N);
    constexpr int BT = (B_LO ? 2 : 1) * Cfg::kBTile;                    // bytes of one (tap, kb) weight tile: [hi][lo]
    if (warp == WC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // resident weights (all threads help)
    for (int u = tid; u < TAPS * KB * N * 8; u += WC_THREADS) {
        const int j = u & 7, row = (u >> 3) % N, tile = (u >> 3) / N;          // tile = tap*KB + kb
        const uint32_t dst = (uint32_t)(tile * BT + (row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
        const size_t src = (size_t)row * KTOT + (size_t)tile * 64 + j * 8;      // k = tap*IC + kb*64 + ...
        cp_async16(sB + dst, Whi + src, true);
        if (B_LO) cp_async16(sB + Cfg::kBTile + dst, Wlo + src, true);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < WC_PRODUCERS / 32) {
        // ------------------------------------------------------------------ producer: one TMA box per window plane
        if (lane == 0) {
            tma_prefetch_desc(&tmXh);
            if (A_LO) tma_prefetch_desc(&tmXl);
            long long it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int p_first = (int)(tile * 128) - (BACK ? Cfg::kHalo : 0);        // may be negative / run past R: zero fill
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = (int)(it % S);
                    const uint32_t ph = (uint32_t)(it / S) & 1u;
                    mbar_wait(smem_u32(&bars[S + s]), ph ^ 1u);
                    const uint32_t st = sA + s * Cfg::kAStage;
                    const uint32_t full = smem_u32(&bars[s]);
                    mbar_arrive_expect_tx(full, (uint32_t)Cfg::kAStage);
                    tma_load_2d(st, &tmXh, full, kb * 64, p_first);
                    if (A_LO) tma_load_2d(st + Cfg::kWinBytes, &tmXl, full, kb * 64, p_first);
                }
            }
        }
        __syncwarp();
    } else if (warp < WC_MMA_WARP) {
        // ------------------------------------------------------------------ epilogue (8 warps: lane quadrant x column half)
        constexpr int GRP = N * 4 / EW;                                         // columns per epilogue warp
        const int q = warp & 3, col0 = ((warp - WC_PRODUCERS / 32) >> 2) * GRP;
        long long ti = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int a = (int)(ti & 1);
            const long long p = tile * 128 + q * 32 + lane;
            typename Epi::Pre pre;                                              // operands the epilogue needs from global memory (ReLU masks):
            if (p < R) ep.prefetch(p, col0, pre);                               // requested BEFORE waiting for the accumulator
            mbar_wait(smem_u32(&bars[2 * S + a]), (uint32_t)(ti >> 1) & 1u);
            tc_fence_after();
            // all TMEM loads of this warp's columns are issued before one wait: their latencies overlap
            constexpr int SUB = GRP < 64 ? GRP : 64;                            // register budget: 64 columns at a time
#pragma unroll
            for (int c0 = 0; c0 < GRP; c0 += SUB) {
                uint32_t rv[SUB / 16][16], rw[FOLD ? SUB / 16 : 1][16];
                const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * ACC_COLS + col0 + c0);
#pragma unroll
                for (int g = 0; g < SUB / 16; ++g) {
                    tmem_ld16_issue(lane_base + g * 16, rv[g]);
                    if (FOLD) tmem_ld16_issue(lane_base + N + g * 16, rw[g]);
                }
#pragma unroll
                for (int g = 0; g < SUB / 16; ++g) {
                    tmem_ld_wait(rv[g]);
                    if (FOLD) tmem_ld_wait(rw[g]);
                }
#pragma unroll
                for (int g = 0; g < SUB / 16; ++g) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rv[g][i]) + (FOLD ? __uint_as_float(rw[g][i]) : 0.f);
                    if (p < R) ep.store16(p, col0 + c0 + g * 16, v, col0, pre);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[2 * S + 2 + a]));
        }
    } else {
        // ------------------------------------------------------------------ MMA issue (whole warp, elected lane)
        constexpr uint32_t idesc = umma_idesc_bf16(N), idesc2 = umma_idesc_bf16(2 * N);
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
                        const int off_t = (t / KW) * GW + (t % KW);
                        const uint32_t shift = (uint32_t)(BACK ? Cfg::kHalo - off_t : off_t) * 128u;  // row-shifted view of the window
                        const uint64_t a_hi = umma_desc_sw128(st + shift), a_lo = umma_desc_sw128(st + Cfg::kWinBytes + shift);
                        const uint32_t bt = uB + (t * KB + kb) * BT;
                        const uint64_t b_hi = umma_desc_sw128(bt), b_lo = umma_desc_sw128(bt + Cfg::kBTile);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)(k * 2);
                            uint32_t accum = (kb | t | k) ? 1u : 0u;
                            if (FOLD) {
                                umma_bf16(acc, a_hi + adv, b_hi + adv, idesc2, accum);            // [hi rows | lo rows] of the weight tile: N' = 2N
                                if (A_LO) umma_bf16(acc, a_lo + adv, b_hi + adv, idesc, 1u);
                            } else {
                                if (A_LO) { umma_bf16(acc, a_lo + adv, b_hi + adv, idesc, accum); accum = 1u; }
                                if (B_LO) { umma_bf16(acc, a_hi + adv, b_lo + adv, idesc, accum); accum = 1u; }
                                umma_bf16(acc, a_hi + adv, b_hi + adv, idesc, accum);
                            }
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
    if (warp == WC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * ACC_COLS) : "memory");
    }
}

template <int GW, int IC, int KH, int KW, int N, bool A_LO, bool B_LO, bool BACK, int EW, class Epi>
static inline cudaError_t launch_winconv_inst(SplitC X, long long R, SplitC W, const Epi& ep, cudaStream_t s) {
    using Cfg = WinCfg<GW, IC, KH, KW, N, A_LO, B_LO>;
    auto kern = winconv_kernel<GW, IC, KH, KW, N, A_LO, B_LO, BACK, EW, Epi>;
    static unsigned long long configured = 0;                       // bit per device ordinal
    {
        cudaError_t e = ensure_dynamic_smem(kern, Cfg::kSmem, &configured);
        if (e != cudaSuccess) return e;
    }
    const long long ntiles = (R + 127) / 128;
    const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    if (R >= (1ll << 31) - 256) return cudaErrorInvalidValue;                        // TMA coordinates are int32
    const CUtensorMap* mh = tmap_2d(X.hi, IC, (uint64_t)R, IC, 64, Cfg::kWinRows);
    const CUtensorMap* ml = A_LO ? tmap_2d(X.lo, IC, (uint64_t)R, IC, 64, Cfg::kWinRows) : mh;
    if (!mh || !ml) return cudaErrorInvalidValue;
    kern<<<grid, WC_PRODUCERS + 32 * EW + 32, Cfg::kSmem, s>>>(*mh, *ml, R, W.hi, W.lo, ep);
    return cudaGetLastError();
}

// A_HAS_LO: whether X has a lo plane at all (u8 frames do not).  Precision policy as in launch_umma2 (weights are B).
// BACK = true is the data gradient: out[q] = sum_taps X[q - off(tap)] . W[tap] with X the (zero-junk) gradient grid.
template <int GW, int IC, int KH, int KW, int N, bool A_HAS_LO, bool BACK = false, int EW = 4, class Epi>
static inline cudaError_t launch_winconv(SplitC X, long long R, SplitC W, const Epi& ep, cudaStream_t s) {
    if (g_fast_math == 1) return launch_winconv_inst<GW, IC, KH, KW, N, false, false, BACK, EW>(X, R, W, ep, s);
    if constexpr (!A_HAS_LO) {
        return launch_winconv_inst<GW, IC, KH, KW, N, false, true, BACK, EW>(X, R, W, ep, s);
    } else {
        if (g_fast_math == 2) return launch_winconv_inst<GW, IC, KH, KW, N, false, true, BACK, EW>(X, R, W, ep, s);
        return launch_winconv_inst<GW, IC, KH, KW, N, true, true, BACK, EW>(X, R, W, ep, s);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Window weight gradient:  dW[tap][c][n] = sum_p X[p + off(tap)][c] * G[p][n]
// X: activations on the input grid ([R][IC]); G: output gradient on the SAME grid ([R][NO], junk pixels are zero).
// Both operands are consumed MN-major (smem lines are pixels = the reduction index), so the staged bytes are the plain
// rows of X and G; tap t multiplies the X window shifted by off(t) lines.  M = IC (64, or 128 as two 64-channel atoms),
// N = NO, one TMEM accumulator per tap.  A work item is (pixel chunk, tap group); chunks are <= 4096 pixels because
// the TMEM accumulation truncates -- partials go to the split-K workspace [chunk][taps*IC][NO].
// PACK_G (NO = 32, X without lo plane): the hi and lo planes of a G row share one 128-byte line, a single N = 64 MMA
// yields X.G_hi | X.G_lo side by side and the epilogue adds the halves.
template <int GW, int IC, int KH, int KW, int TG, int NO, bool X_LO, bool G_LO, bool PACK_G, int KP, bool BIAS>
struct WinWgradCfg {
    static constexpr int kTaps = KH * KW;
    static_assert(TG == kTaps || TG == KW, "tap groups are the whole kernel or one kernel row");
    static constexpr int kGroups = kTaps / TG;
    static constexpr int kHG = (TG == kTaps) ? (KH - 1) * GW + (KW - 1) : (KW - 1);       // halo inside one group
    static constexpr int kAtoms = IC / 64;
    static constexpr int kWinRows = (KP + kHG + 7) / 8 * 8;
    static constexpr int kWinBytes = kWinRows * 128;
    static constexpr int kXBytes = (X_LO ? 2 : 1) * kAtoms * kWinBytes;
    static constexpr int kGPlane = KP * 128;
    static constexpr int kGBytes = (PACK_G ? 1 : (G_LO ? 2 : 1)) * kGPlane;
    static constexpr int kStage = kXBytes + kGBytes;
    static constexpr int kNMMA = PACK_G ? 64 : NO;
    // IC = 64 would leave half of the M = 128 datapath idle (an M = 64 MMA costs the same cycles): two taps share one
    // MMA instead -- atom 0 = the window at tap 2a, atom 1 = the SAME buffer off(2a+1) - off(2a) lines further (the
    // descriptor's atom stride is plain address arithmetic; tools/shift_probe.py).  Rows 64..127 of accumulator a = tap 2a+1.
    static constexpr bool kPair = IC == 64;
    static constexpr int kNAcc = kPair ? (TG + 1) / 2 : TG;
    static constexpr int kUsedCols = (kNAcc + (BIAS ? 1 : 0)) * kNMMA;          // BIAS: one more accumulator, ones^T . G
    static constexpr int kCols = kUsedCols <= 32 ? 32 : kUsedCols <= 64 ? 64 : kUsedCols <= 128 ? 128 : kUsedCols <= 256 ? 256 : 512;
    static constexpr int kOnesBytes = 4096;                                     // two 16-line atoms of bf16 1.0
    static constexpr int kStagesRaw = (227 * 1024 - 2048 - kOnesBytes) / kStage;
    // three stages when that lets two CTAs share an SM (prologue/epilogue of one hides behind the other), else up to five
    static constexpr bool kTwoCtas = kCols <= 256 && (2 * (3 * kStage + kOnesBytes + 1280) <= 227 * 1024);
    static constexpr int kStages = kTwoCtas ? 3 : (kStagesRaw > 5 ? 5 : kStagesRaw);
    static constexpr int kSmem = kStages * kStage + kOnesBytes + 1024 + 256;
    static_assert(IC % 64 == 0 && IC <= 128 && (NO == 32 || NO == 64) && KP % 16 == 0 && kUsedCols <= 512, "window wgrad shape");
    static_assert(!PACK_G || (NO == 32 && !X_LO), "PACK_G packs hi|lo of a 32-channel gradient row");
    static_assert(kStages >= 2, "not enough shared memory");
};

template <int GW, int IC, int KH, int KW, int TG, int NO, bool X_LO, bool G_LO, bool PACK_G, int KP, bool BIAS>
__global__ void __launch_bounds__(UM_THREADS, 1)
winwgrad_kernel(const bf16* __restrict__ Xhi, const bf16* __restrict__ Xlo, const bf16* __restrict__ Ghi, const bf16* __restrict__ Glo,
                long long R, int chunk /* pixels per item, multiple of KP */, float* __restrict__ ws, float* __restrict__ bias_ws /* [chunks][NO] */) {
    using Cfg = WinWgradCfg<GW, IC, KH, KW, TG, NO, X_LO, G_LO, PACK_G, KP, BIAS>;
    constexpr int S = Cfg::kStages, ATOMS = Cfg::kAtoms, NMMA = Cfg::kNMMA;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t sbase = raw + pad;
    const uint32_t sOnes = sbase + S * Cfg::kStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStage + Cfg::kOnesBytes);       // full[S] | empty[S] | acc
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int group = blockIdx.y;
    const int tap0 = group * TG;
    const int wbase = (tap0 / KW) * GW + (tap0 % KW);                           // first line of this group's window
    const long long c0 = (long long)blockIdx.x * chunk;
    const long long c1 = (c0 + chunk < R) ? c0 + chunk : R;
    const int nst = (int)((c1 - c0 + KP - 1) / KP);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(smem_u32(&bars[s]), 8); mbar_init(smem_u32(&bars[S + s]), 1); }
        mbar_init(smem_u32(&bars[2 * S]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const bool do_bias = BIAS && group == 0;
    if (BIAS) {
        for (int u = tid; u < Cfg::kOnesBytes / 4; u += UM_THREADS) reinterpret_cast<uint32_t*>(smem + S * Cfg::kStage)[u] = 0x3F803F80u;
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ------------------------------------------------------------------ producers
        for (int it = 0; it < nst; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            mbar_wait(smem_u32(&bars[S + s]), ph ^ 1u);
            const uint32_t st = sbase + s * Cfg::kStage;
            const long long p0 = c0 + (long long)it * KP;
            for (int u = tid; u < Cfg::kWinRows * 8 * ATOMS; u += UM_PRODUCERS) {
                const int j = u & 7, row = (u >> 3) % Cfg::kWinRows, atom = (u >> 3) / Cfg::kWinRows;
                const long long p = p0 + wbase + row;
                const bool ok = p < R;
                const size_t src = ok ? (size_t)p * IC + atom * 64 + j * 8 : 0;
                const uint32_t dst = (uint32_t)(atom * Cfg::kWinBytes + (row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
                cp_async16(st + dst, Xhi + src, ok);
                if (X_LO) cp_async16(st + ATOMS * Cfg::kWinBytes + dst, Xlo + src, ok);
            }
            const uint32_t sg = st + Cfg::kXBytes;
            if (PACK_G) {
                for (int u = tid; u < KP * 8; u += UM_PRODUCERS) {
                    const int j = u & 7, row = u >> 3;
                    const long long p = p0 + row;
                    const bool ok = p < c1;
                    const size_t src = ok ? (size_t)p * NO + (j & 3) * 8 : 0;
                    cp_async16(sg + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4)), (j < 4 ? Ghi : Glo) + src, ok);
                }
            } else {
                constexpr int CH = NO / 8;                                       // 16-byte chunks per gradient row
                for (int u = tid; u < KP * CH; u += UM_PRODUCERS) {
                    const int j = u % CH, row = u / CH;
                    const long long p = p0 + row;
                    const bool ok = p < c1;
                    const size_t src = ok ? (size_t)p * NO + j * 8 : 0;
                    const uint32_t dst = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4));
                    cp_async16(sg + dst, Ghi + src, ok);
                    if (G_LO) cp_async16(sg + Cfg::kGPlane + dst, Glo + src, ok);
                }
            }
            cp_async_commit();
            if (it > 0) {                                                       // hand over the PREVIOUS stage: two groups stay in flight per thread
                cp_async_wait<1>();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[(it - 1) % S]));
            }
        }
        cp_async_wait<0>();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[(nst - 1) % S]));
        // ------------------------------------------------------------------ epilogue: TMEM -> partial [chunk][taps*IC][NO]
        mbar_wait(smem_u32(&bars[2 * S]), 0);
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        float* out = ws + (size_t)blockIdx.x * (Cfg::kTaps * IC * NO);
        constexpr int NCH = NO / 16;                                            // 16-column chunks per accumulator
        for (int u = half; u < Cfg::kNAcc * NCH; u += 2) {
            const int a = u / NCH, c = (u % NCH) * 16;
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NMMA + c), v);
            if (PACK_G) {
                float w[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NMMA + 32 + c), w);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += w[i];
            }
            // accumulator row -> (tap, channel).  M = 128: row = 32q + lane; paired taps put tap 2a+1 in rows 64..127.
            // M = 64 (unpaired last tap): rows live in lanes 32q + (0..15)
            int tap, ch; bool ok = true;
            if (Cfg::kPair) {
                const bool paired = 2 * a + 1 < TG;
                if (paired) { tap = 2 * a + (q >> 1); ch = (q & 1) * 32 + lane; }
                else { tap = 2 * a; ch = q * 16 + lane; ok = lane < 16; }
            } else { tap = a; ch = q * 32 + lane; }
            if (ok) {
                float* o = out + ((size_t)(tap0 + tap) * IC + ch) * NO + c;
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
        if (do_bias && warp == 0) {                                              // row 0 of the ones^T . G accumulator = column sums
            for (int c = 0; c < NO; c += 16) {
                float v[16];
                tmem_ld16(tmem_base + (uint32_t)(Cfg::kNAcc * NMMA + c), v);
                if (PACK_G) {
                    float w[16];
                    tmem_ld16(tmem_base + (uint32_t)(Cfg::kNAcc * NMMA + 32 + c), w);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += w[i];
                }
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) bias_ws[(size_t)blockIdx.x * NO + c + i] = v[i];
                }
            }
        }
        tc_fence_before();
    } else {
        // ------------------------------------------------------------------ MMA issue
        constexpr uint32_t idesc_base = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NMMA >> 3) << 17)) | (1u << 15) | (1u << 16);
        constexpr uint32_t idesc = idesc_base | ((uint32_t)(IC >> 4) << 24);      // M = IC (bias row, unpaired taps)
        constexpr uint32_t idesc128 = idesc_base | (8u << 24);                    // M = 128
        const bool leader = elect_one();
        const uint32_t uS = __shfl_sync(0xffffffffu, sbase, 0), uT = __shfl_sync(0xffffffffu, tmem_base, 0);
        for (int it = 0; it < nst; ++it) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)(it / S) & 1u;
            mbar_wait(smem_u32(&bars[s]), ph);
            tc_fence_after();
            const uint32_t st = uS + s * Cfg::kStage;
            if (leader) {
                // loop order k -> term -> accumulator
                const uint64_t g_hi = umma_desc_sw128_mn(st + Cfg::kXBytes, 8192), g_lo = umma_desc_sw128_mn(st + Cfg::kXBytes + Cfg::kGPlane, 8192);
                const uint64_t ones = umma_desc_sw128_mn(sOnes, 2048);
#pragma unroll
                for (int k = 0; k < KP / 16; ++k) {
                    const uint64_t adv = (uint64_t)(k * 128);                    // 16 lines x 128 B, >> 4
                    const uint32_t first = (it | k) ? 1u : 0u;
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        if (term == 0 && !X_LO) continue;
                        if (term == 1 && !(G_LO && !PACK_G)) continue;
                        // accumulate flag: only the first issued term of the first k-step overwrites
                        const bool first_term = (term == 0) || (term == 1 && !X_LO) || (term == 2 && !X_LO && !(G_LO && !PACK_G));
                        const uint32_t accum = first_term ? first : 1u;
#pragma unroll
                        for (int a = 0; a < Cfg::kNAcc; ++a) {
                            const int t1 = tap0 + (Cfg::kPair ? 2 * a : a);     // (ky, kx); tap0 is a multiple of KW or 0
                            const bool paired = Cfg::kPair && (2 * a + 1 < TG);
                            const int off1 = (t1 / KW) * GW + (t1 % KW), off2 = ((t1 + 1) / KW) * GW + ((t1 + 1) % KW);
                            const uint32_t shift = (uint32_t)(off1 - wbase) * 128u;
                            const uint32_t lbo = paired ? (uint32_t)(off2 - off1) * 128u : (uint32_t)Cfg::kWinBytes;
                            const uint32_t id = (paired || IC == 128) ? idesc128 : idesc;
                            const uint64_t x = umma_desc_sw128_mn(st + (term == 0 ? ATOMS * Cfg::kWinBytes : 0) + shift, lbo);
                            umma_bf16(uT + a * NMMA, x + adv, (term == 1 ? g_lo : g_hi) + adv, id, accum);
                        }
                        if (do_bias && term != 0)                                // ones^T . G (hi, and lo when it is a separate plane)
                            umma_bf16(uT + Cfg::kNAcc * NMMA, ones, (term == 1 ? g_lo : g_hi) + adv, idesc, (term == 2 && G_LO && !PACK_G) ? 1u : first);
                    }
                }
                umma_commit(smem_u32(&bars[S + s]));
                if (it == nst - 1) umma_commit(smem_u32(&bars[2 * S]));
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kCols) : "memory");
    }
}

template <int GW, int IC, int KH, int KW, int TG, int NO, bool X_LO, bool G_LO, bool PACK_G, int KP, bool BIAS>
static inline cudaError_t launch_winwgrad_inst(SplitC X, SplitC G, long long R, int chunk, float* ws, float* bias_ws, cudaStream_t s) {
    using Cfg = WinWgradCfg<GW, IC, KH, KW, TG, NO, X_LO, G_LO, PACK_G, KP, BIAS>;
    auto kern = winwgrad_kernel<GW, IC, KH, KW, TG, NO, X_LO, G_LO, PACK_G, KP, BIAS>;
    static unsigned long long configured = 0;                       // bit per device ordinal
    {
        cudaError_t e = ensure_dynamic_smem(kern, Cfg::kSmem, &configured);
        if (e != cudaSuccess) return e;
    }
    const int nchunks = (int)((R + chunk - 1) / chunk);
    kern<<<dim3(nchunks, Cfg::kGroups), UM_THREADS, Cfg::kSmem, s>>>(X.hi, X.lo, G.hi, G.lo, R, chunk, ws, bias_ws);
    return cudaGetLastError();
}

// Precision policy of the other weight gradients (LO_NO_WEIGHT): strict = all three split products, otherwise hi.hi.
// BIAS: also emit per-chunk column sums of G (the bias gradient) into bias_ws[chunk][NO].
template <int GW, int IC, int KH, int KW, int TG, int NO, bool X_HAS_LO, int KP, bool BIAS>
static inline cudaError_t launch_winwgrad(SplitC X, SplitC G, long long R, int chunk, float* ws, float* bias_ws, cudaStream_t s) {
    if (g_fast_math != 0) return launch_winwgrad_inst<GW, IC, KH, KW, TG, NO, false, false, false, KP, BIAS>(X, G, R, chunk, ws, bias_ws, s);
    if constexpr (!X_HAS_LO && NO == 32) {
        return launch_winwgrad_inst<GW, IC, KH, KW, TG, NO, false, true, true, KP, BIAS>(X, G, R, chunk, ws, bias_ws, s);
    } else {
        return launch_winwgrad_inst<GW, IC, KH, KW, TG, NO, X_HAS_LO, true, false, KP, BIAS>(X, G, R, chunk, ws, bias_ws, s);
    }
}

}  // namespace r2d2
