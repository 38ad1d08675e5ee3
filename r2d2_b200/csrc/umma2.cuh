// tcgen05 gather-GEMM, v2 data path: operands live in HBM already split into bf16 hi/lo planes.
//
//   D[m][n] = sum_k A(m,k) * B(n,k)          CTA tile 128 x BN, 64 reduction indices per stage
//
// Every activation / gradient / packed weight that feeds a contraction is stored as two bf16
// tensors of identical layout (x = hi + lo, |x - hi - lo| <= 2^-17 |x|) by the kernel that
// produces it, so the producer warps of this kernel do no arithmetic: they copy 16-byte chunks
// (8 bf16) with cp.async straight into the canonical UMMA shared-memory layouts, computing only
// the gather address (im2col, dgrad, row maps ...) per chunk:
//
//   K-major operand  (kMN = false): smem line = operand row, 64 reduction indices (128 B) per line;
//                    a chunk is 8 consecutive reduction indices of one row.
//   MN-major operand (kMN = true) : smem line = reduction index, 64 operand rows (128 B) per line,
//                    64-row atoms side by side (LBO apart); a chunk is 8 consecutive operand rows
//                    at one reduction index.  This is how transposed operands (every wgrad) are fed
//                    without a transposing store: the source is contiguous along the operand rows.
// Both use the 128-byte swizzle (16-byte chunk index ^= line & 7) and 8-line groups 1024 B apart.
//
// Math: bf16x3 (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) or, with kFast, hi*hi only.
// An operand with kHasLo = false is exact in bf16 (u8 pixels): its lo plane is never touched.
//
// Pipeline: kStages stages; each producer thread issues its cp.async chunks for stage kt, commits the
// group, then waits for the group of stage kt-(kStages-1), fences generic->async proxy and arrives
// (one arrive per warp) on full[]; one thread issues the MMAs; tcgen05.commit arms empty[].
#pragma once
#include "umma.cuh"

namespace r2d2 {

using bf16 = __nv_bfloat16;

struct SplitC { const bf16* hi; const bf16* lo; };   // read-only split tensor
struct SplitW { bf16* hi; bf16* lo; };               // writable split tensor

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;                      // src-size 0 -> 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// MN-major SWIZZLE_128B descriptor: LBO = distance between 64-row atoms, SBO = 1024 B (8 reduction lines)
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_major(int n, bool a_mn, bool b_mn) {
    return umma_idesc_bf16(n) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
}

// hi/lo split of a float pair / quad for epilogues
__device__ __forceinline__ void split_store8(bf16* hi, bf16* lo, size_t off, const float* v) {   // 8 floats -> 16 B + 16 B
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}
// 16 floats -> one 32-byte sector per plane (256-bit stores: a lane fills whole sectors even when neighbouring lanes
// write rows far apart).  off must be a multiple of 16 elements.
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
                 "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void split_store16(bf16* hi, bf16* lo, size_t off, const float* v) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
    st_global_256(hi + off, h);
    st_global_256(lo + off, l);
}
__device__ __forceinline__ float split_load(const bf16* hi, const bf16* lo, size_t off) {
    return __bfloat162float(hi[off]) + __bfloat162float(lo[off]);
}
__device__ __forceinline__ void split_load8(const bf16* hi, const bf16* lo, size_t off, float* v) {
    const uint4 h = *reinterpret_cast<const uint4*>(hi + off);
    const uint4 l = *reinterpret_cast<const uint4*>(lo + off);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
        v[2 * j + 1] = __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lw[j] & 0xFFFF0000u);
    }
}

template <int BN, bool A_LO, bool B_LO> struct Umma2Cfg {
    static constexpr int kTileA = UM_BM * UM_BK * 2;
    static constexpr int kTileB = BN * UM_BK * 2;
    static constexpr int kStageBytes = (A_LO ? 2 : 1) * kTileA + (B_LO ? 2 : 1) * kTileB;
#ifndef R2D2_STAGE_BUDGET_KB
#define R2D2_STAGE_BUDGET_KB 100
#endif
    static constexpr int kBudget = R2D2_STAGE_BUDGET_KB * 1024;
    static constexpr int kStages = (kStageBytes * 4 <= kBudget) ? 4 : ((kStageBytes * 3 <= kBudget) ? 3 : 2);
    static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
    static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

// Per-thread staging plan of one operand tile (ROWS operand rows x 64 reduction indices).  A thread's chunks keep
// the same operand rows for the whole K loop, so the row part of the gather address (frame / pixel decomposition,
// row-map lookup, bounds) is computed ONCE into `ctx`; per k-block only the reduction-index part is added.
template <int ROWS, class Src> struct OperandStager {
    static constexpr int kUnits = ROWS * 8;
    static constexpr int kIters = (kUnits + UM_PRODUCERS - 1) / UM_PRODUCERS;
    typename Src::RCtx ctx[kIters];
    uint32_t dst[kIters];
    int kofs[kIters];                // reduction-index offset of the chunk inside a stage

    __device__ __forceinline__ void init(const Src& src, int row0, int tid) {
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int u = tid + it * UM_PRODUCERS;
            const int j = u & 7;
            if constexpr (!Src::kMN) {
                const int line = u >> 3;                    // operand row within the tile
                ctx[it] = src.rctx(row0 + line);
                kofs[it] = j * 8;
                dst[it] = (uint32_t)((line >> 3) * 1024 + (line & 7) * 128 + ((j ^ (line & 7)) << 4));
            } else {
                const int line = (u >> 3) & 63, atom = u >> 9;   // reduction index within the stage, 64-row atom
                ctx[it] = src.rctx(row0 + atom * 64 + j * 8);
                kofs[it] = line;
                dst[it] = (uint32_t)(atom * 8192 + (line >> 3) * 1024 + (line & 7) * 128 + ((j ^ (line & 7)) << 4));
            }
        }
    }
    __device__ __forceinline__ void issue(const Src& src, int kbase, int k_end, uint32_t hi_tile, uint32_t lo_tile, bool want_lo,
                                          int tid) const {
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            if (kUnits % UM_PRODUCERS != 0 && tid + it * UM_PRODUCERS >= kUnits) break;
            const int k = kbase + kofs[it];
            const long long off = (k < k_end) ? src.chunk(ctx[it], k) : -1;
            const bool ok = off >= 0;
            const long long o = ok ? off : 0;
            cp_async16(hi_tile + dst[it], src.hi + o, ok);
            if (Src::kHasLo && want_lo) cp_async16(lo_tile + dst[it], src.lo + o, ok);
        }
    }
};


template <int BN, bool WANT_A_LO, bool WANT_B_LO, class ASrcT, class BSrcT, class EpiT>
__global__ void __launch_bounds__(UM_THREADS, R2D2_STAGE_BUDGET_KB <= 104 ? 2 : 1)
umma2_kernel(const ASrcT a_, const BSrcT b_, const EpiT ep_, int K, int k_per_split) {
    using ASrc = typename std::remove_cv<typename std::remove_reference<decltype(sel_z(a_, 0))>::type>::type;
    using BSrc = typename std::remove_cv<typename std::remove_reference<decltype(sel_z(b_, 0))>::type>::type;
    constexpr bool A_LO = ASrc::kHasLo && WANT_A_LO, B_LO = BSrc::kHasLo && WANT_B_LO;
    using Cfg = Umma2Cfg<BN, A_LO, B_LO>;
    constexpr int kStages = Cfg::kStages;
    static_assert(BN == 16 || BN == 32 || BN == 64 || BN == 128 || BN == 256, "UMMA N / TMEM columns");
    static_assert(!BSrc::kMN || BN % 64 == 0, "MN-major B needs whole 64-row atoms");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t smem_base = raw + pad;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z;
    const ASrc& as = sel_z(a_, z);
    const BSrc& bs = sel_z(b_, z);
    const auto& ep = sel_z(ep_, z);
    const int m0 = blockIdx.y * UM_BM, n0 = blockIdx.x * BN;
    const int k_begin = IsPair<ASrcT>::value ? 0 : z * k_per_split;
    const int k_end = IsPair<ASrcT>::value ? K : min(K, k_begin + k_per_split);
    const int nk = (k_end - k_begin + UM_BK - 1) / UM_BK;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&bars[s]), UM_PRODUCERS / 32);
            mbar_init(smem_u32(&bars[kStages + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * kStages]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == UM_PRODUCERS / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(Cfg::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // stage layout: [A hi][A lo?][B hi][B lo?]
    constexpr int kOffALo = Cfg::kTileA;
    constexpr int kOffBHi = (A_LO ? 2 : 1) * Cfg::kTileA;
    constexpr int kOffBLo = kOffBHi + Cfg::kTileB;

    if (warp < UM_PRODUCERS / 32) {
        // ------------------------------------------------------------------ producers (cp.async)
        constexpr int D = kStages - 1;                    // prefetch distance in stages
        OperandStager<UM_BM, ASrc> sa;
        OperandStager<BN, BSrc> sb;
        sa.init(as, m0, tid);
        sb.init(bs, n0, tid);
        for (int kt = 0; kt < nk + D; ++kt) {
            if (kt < nk) {
                const int s = kt % kStages;
                const uint32_t ph = (uint32_t)(kt / kStages) & 1u;
                mbar_wait(smem_u32(&bars[kStages + s]), ph ^ 1u);
                const uint32_t st = smem_base + s * Cfg::kStageBytes;
                const int kbase = k_begin + kt * UM_BK;
                sa.issue(as, kbase, k_end, st, st + kOffALo, A_LO, tid);
                sb.issue(bs, kbase, k_end, st + kOffBHi, st + kOffBLo, B_LO, tid);
            }
            cp_async_commit();
            if (kt >= D) {
                cp_async_wait<D>();                        // the group of stage kt-D has landed
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[(kt - D) % kStages]));
            }
        }
        // ------------------------------------------------------------------ epilogue
        if (nk > 0) {
            mbar_wait(smem_u32(&bars[2 * kStages]), 0);
            tc_fence_after();
        }
        const int row = (warp & 3) * 32 + lane;
        constexpr int kHalf = BN >= 32 ? BN / 2 : BN;
        const int cbeg = (BN >= 32 && warp >= 4) ? kHalf : 0;
        const int cend = (BN >= 32) ? cbeg + kHalf : (warp >= 4 ? 0 : BN);
        for (int c = cbeg; c < cend; c += 16) {
            float v[16];
            if (nk > 0) tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c, v);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
            ep.store16(m0 + row, n0 + c, v, z);
        }
        tc_fence_before();
    } else {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop (warp-uniform descriptor arithmetic -> uniform registers feed UTCHMMA); only the
        // elected lane issues the MMAs and commits.
        {
            constexpr uint32_t idesc = umma_idesc_bf16_major(BN, ASrc::kMN, BSrc::kMN);
            constexpr uint32_t kLboA = 64 * 128, kLboB = 64 * 128;
            const bool leader = elect_one();
            // bases broadcast through a shuffle: everything derived from them stays warp-uniform for the compiler
            const uint32_t u_smem = __shfl_sync(0xffffffffu, smem_base, 0), u_tmem = __shfl_sync(0xffffffffu, tmem_base, 0);
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % kStages;
                const uint32_t ph = (uint32_t)(kt / kStages) & 1u;
                mbar_wait(smem_u32(&bars[s]), ph);
                tc_fence_after();
                const uint32_t st = u_smem + s * Cfg::kStageBytes;
                auto mk_a = [&](uint32_t addr) { return ASrc::kMN ? umma_desc_sw128_mn(addr, kLboA) : umma_desc_sw128(addr); };
                auto mk_b = [&](uint32_t addr) { return BSrc::kMN ? umma_desc_sw128_mn(addr, kLboB) : umma_desc_sw128(addr); };
                const uint64_t a_hi = mk_a(st), a_lo = mk_a(st + kOffALo), b_hi = mk_b(st + kOffBHi), b_lo = mk_b(st + kOffBLo);
                // per K=16 step: K-major +32 B inside the 128 B line; MN-major +16 lines = 2048 B
                constexpr uint64_t kAdvA = (ASrc::kMN ? 2048 : 32) >> 4, kAdvB = (BSrc::kMN ? 2048 : 32) >> 4;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < UM_BK / 16; ++k) {
                        const uint64_t da = kAdvA * k, db = kAdvB * k;
                        uint32_t acc = (kt | k) ? 1u : 0u;
                        if (A_LO) { umma_bf16(u_tmem, a_lo + da, b_hi + db, idesc, acc); acc = 1u; }
                        if (B_LO) { umma_bf16(u_tmem, a_hi + da, b_lo + db, idesc, acc); acc = 1u; }
                        umma_bf16(u_tmem, a_hi + da, b_hi + db, idesc, acc);
                    }
                    umma_commit(smem_u32(&bars[kStages + s]));
                }
                __syncwarp();
            }
            if (nk > 0 && leader) umma_commit(smem_u32(&bars[2 * kStages]));
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == UM_PRODUCERS / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
    }
}

// Precision policy.  g_fast_math: 0 = strict (every split operand contributes hi and lo: bf16x3 products),
// 1 = fast (hi planes only: plain bf16 products), 2 = balanced: hi+lo only for WEIGHT operands of the encoder
// contractions (activations / gradients go in as single bf16; their rounding averages out over K >= 256 terms),
// while the recurrence, input projection and dueling head stay strict.  A call site states what its operands are:
enum LoPolicy { LO_STRICT = 0,   // keep both lo planes unless the mode is fast
                LO_WEIGHT_B = 1, // B is a weight matrix, A an activation/gradient tensor
                LO_NO_WEIGHT = 2 /* neither operand is a weight (weight-gradient contractions) */ };
extern int g_fast_math;

template <int BN, bool AL, bool BL, class ASrc, class BSrc, class Epi>
static inline cudaError_t launch_umma2_inst(const ASrc& a, const BSrc& b, const Epi& ep, dim3 grid, int K, int k_per_split, cudaStream_t s) {
    using A0 = typename std::remove_cv<typename std::remove_reference<decltype(sel_z(a, 0))>::type>::type;
    using B0 = typename std::remove_cv<typename std::remove_reference<decltype(sel_z(b, 0))>::type>::type;
    using Cfg = Umma2Cfg<BN, A0::kHasLo && AL, B0::kHasLo && BL>;
    auto kern = umma2_kernel<BN, AL, BL, ASrc, BSrc, Epi>;
    static unsigned long long configured = 0;                       // bit per device ordinal
    {
        cudaError_t e = ensure_dynamic_smem(kern, Cfg::kSmem, &configured);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, UM_THREADS, Cfg::kSmem, s>>>(a, b, ep, K, k_per_split);
    return cudaGetLastError();
}

template <int BN, int POL, class ASrc, class BSrc, class Epi>
static inline cudaError_t launch_umma2_grid(const ASrc& a, const BSrc& b, const Epi& ep, dim3 grid, int K, int k_per_split, cudaStream_t s) {
    if (g_fast_math == 1) return launch_umma2_inst<BN, false, false>(a, b, ep, grid, K, k_per_split, s);
    if (g_fast_math == 2 && POL == LO_WEIGHT_B) return launch_umma2_inst<BN, false, true>(a, b, ep, grid, K, k_per_split, s);
    if (g_fast_math == 2 && POL == LO_NO_WEIGHT) return launch_umma2_inst<BN, false, false>(a, b, ep, grid, K, k_per_split, s);
    return launch_umma2_inst<BN, true, true>(a, b, ep, grid, K, k_per_split, s);
}

template <int BN, int POL = LO_STRICT, class ASrc, class BSrc, class Epi>
static inline cudaError_t launch_umma2(const ASrc& a, const BSrc& b, const Epi& ep, int M, int N, int K, int splits, cudaStream_t s) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    int k_per_split = (K + splits - 1) / splits;
    k_per_split = (k_per_split + UM_BK - 1) / UM_BK * UM_BK;
    dim3 grid((N + BN - 1) / BN, (M + UM_BM - 1) / UM_BM, splits);
    return launch_umma2_grid<BN, POL>(a, b, ep, grid, K, k_per_split, s);
}

// ---------------------------------------------------------------------------------------------
// chunk sources.  chunk(row, k) returns the element offset (into both planes) of the 8-element
// chunk starting at (row, k) -- along k for K-major sources, along rows for MN-major sources --
// or -1 when it lies outside the operand (zero fill).
// ---------------------------------------------------------------------------------------------
struct SrcMatK {   // X(row, k) = p[row*ld + k]
    static constexpr bool kMN = false, kHasLo = true;
    using RCtx = long long;
    const bf16* hi; const bf16* lo; int rows, K; long long ld;
    __device__ __forceinline__ RCtx rctx(int row) const { return row < rows ? row * ld : -1; }
    __device__ __forceinline__ long long chunk(RCtx c, int k) const { return (c >= 0 && k < K) ? c + k : -1; }
};
struct SrcMatMN {  // X(row, k) = p[k*ld + row]   (transposed use of a [K][rows] matrix)
    static constexpr bool kMN = true, kHasLo = true;
    using RCtx = long long;
    const bf16* hi; const bf16* lo; int rows, K; long long ld;
    __device__ __forceinline__ RCtx rctx(int row) const { return row < rows ? row : -1; }
    __device__ __forceinline__ long long chunk(RCtx c, int k) const { return (c >= 0 && k < K) ? k * ld + c : -1; }
};
struct SrcRowGatherK {   // X(r, k) = p[src[r]*ld + k]
    static constexpr bool kMN = false, kHasLo = true;
    using RCtx = long long;
    const bf16* hi; const bf16* lo; const int* src; int rows, K; long long ld;
    __device__ __forceinline__ RCtx rctx(int row) const {
        if (row >= rows) return -1;
        const int s = __ldg(src + row);
        return s >= 0 ? s * ld : -1;
    }
    __device__ __forceinline__ long long chunk(RCtx c, int k) const { return (c >= 0 && k < K) ? c + k : -1; }
};
struct SrcRowGatherMN {  // X(j, r) = p[src[r]*ld + j]
    static constexpr bool kMN = true, kHasLo = true;
    using RCtx = long long;
    const bf16* hi; const bf16* lo; const int* src; int rows, K; long long ld;   // rows: width limit (j); K: gathered rows
    __device__ __forceinline__ RCtx rctx(int row) const { return row < rows ? row : -1; }
    __device__ __forceinline__ long long chunk(RCtx c, int k) const {
        if (c < 0 || k >= K) return -1;
        const int s = __ldg(src + k);
        return s >= 0 ? s * ld + c : -1;
    }
};
// NHWC im2col:  pixel m = (frame, oy, ox),  kernel index kk = (ky, kx, c); IC % 8 == 0.  offset = pix(m) + tap(kk)
template <int IH, int IW, int IC, int OH, int OW, int KH, int KW, int S>
struct ConvGeom {
    __device__ static __forceinline__ long long pix(int m, int nframes) {
        if (m >= nframes * OH * OW) return -1;
        const int f = m / (OH * OW), p = m - f * (OH * OW), oy = p / OW, ox = p - oy * OW;
        return (((long long)f * IH + S * oy) * IW + S * ox) * IC;
    }
    __device__ static __forceinline__ long long tap(int kk) {
        if (kk >= KH * KW * IC) return -1;
        const int t = kk / IC, c = kk - t * IC, ky = t / KW, kx = t - ky * KW;
        return (long long)(ky * IW + kx) * IC + c;
    }
};
template <int IH, int IW, int IC, int OH, int OW, int KH, int KW, int S, bool HASLO = true>
struct SrcConvK {    // K-major: row = pixel, k = kernel index
    static constexpr bool kMN = false, kHasLo = HASLO;
    using G = ConvGeom<IH, IW, IC, OH, OW, KH, KW, S>;
    using RCtx = long long;
    const bf16* hi; const bf16* lo; int nframes;
    __device__ __forceinline__ RCtx rctx(int m) const { return G::pix(m, nframes); }
    __device__ __forceinline__ long long chunk(RCtx c, int k) const {
        const long long t = G::tap(k);
        return (c >= 0 && t >= 0) ? c + t : -1;
    }
};
template <int IH, int IW, int IC, int OH, int OW, int KH, int KW, int S, bool HASLO = true>
struct SrcConvMN {   // MN-major (wgrad): row = kernel index (8 consecutive channels), k = pixel
    static constexpr bool kMN = true, kHasLo = HASLO;
    using G = ConvGeom<IH, IW, IC, OH, OW, KH, KW, S>;
    using RCtx = long long;
    const bf16* hi; const bf16* lo; int nframes;
    __device__ __forceinline__ RCtx rctx(int kk) const { return G::tap(kk); }
    __device__ __forceinline__ long long chunk(RCtx c, int m) const {
        const long long p = G::pix(m, nframes);
        return (c >= 0 && p >= 0) ? c + p : -1;
    }
};
// ---------------------------------------------------------------------------------------------
// epilogues (16 consecutive columns of one row)
// ---------------------------------------------------------------------------------------------
struct Epi2Partial {          // split-K partial, fp32
    float* ws; int M, N;
    __device__ __forceinline__ void store16(int m, int n, const float (&v)[16], int z) const {
        if (m >= M) return;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            if (n + j < N) *reinterpret_cast<float4*>(ws + ((size_t)z * M + m) * N + n + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
};
struct Epi2F32 {              // plain fp32 store (test path / XP): out[m*ld+n] = v*scale + bias[n]
    float* out; const float* bias; int M, N; long long ld; float scale;
    __device__ __forceinline__ void store16(int m, int n, const float (&v)[16], int) const {
        if (m >= M) return;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            if (n + j >= N) break;
            float4 r;
            r.x = v[j] * scale + (bias ? __ldg(bias + n + j) : 0.f);
            r.y = v[j + 1] * scale + (bias ? __ldg(bias + n + j + 1) : 0.f);
            r.z = v[j + 2] * scale + (bias ? __ldg(bias + n + j + 2) : 0.f);
            r.w = v[j + 3] * scale + (bias ? __ldg(bias + n + j + 3) : 0.f);
            *reinterpret_cast<float4*>(out + m * ld + n + j) = r;
        }
    }
};
template <bool kRelu>
struct Epi2BiasSplit {        // out(split)[m*ld+n] = act(v*scale + bias[n]);  N % 8 == 0
    SplitW out; const float* bias; int M, N; long long ld; float scale;
    __device__ __forceinline__ void store16(int m, int n, const float (&v)[16], int) const {
        if (m >= M) return;
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            if (n + j >= N) break;
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float x = v[j + i] * scale + (bias ? __ldg(bias + n + j + i) : 0.f);
                r[i] = kRelu ? fmaxf(x, 0.f) : x;
            }
            split_store8(out.hi, out.lo, (size_t)(m * ld + n + j), r);
        }
    }
};

}  // namespace r2d2
