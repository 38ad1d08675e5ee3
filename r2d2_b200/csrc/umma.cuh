// tcgen05 (5th-gen tensor core) main loop for the gather-GEMM template of gemm.cuh.
//
//   C[m][n] = sum_k A(m,k) * B(n,k)     CTA tile 128 x BN, K in blocks of 64
//
// Operands are fp32 (or u8 frames) in HBM and are fetched through the same ALoad/BLoad functors
// as the FFMA baseline.  Producer warps split every value into bf16 hi + bf16 lo
// (x = hi + lo + O(2^-17 x)), and write both halves into K-major, 128-byte-swizzled smem tiles
// (the canonical UMMA layout: 8-row x 128-byte atoms, SBO = 1024 B).  One elected thread issues
// tcgen05.mma kind::f16 with the accumulator in TMEM:
//     TERMS = 3 : hi*hi + hi*lo + lo*hi   (error ~2^-16 relative per product: the parity mode)
//     TERMS = 1 : hi*hi                   (plain bf16: the fast mode)
// Pipeline: kStages smem stages, full[] barriers armed by the producer warps (after
// fence.proxy.async so the tensor core sees the generic-proxy stores), empty[] barriers armed by
// tcgen05.commit.  Epilogue: all 8 producer warps read the 128 x BN fp32 accumulator back with
// tcgen05.ld (warp w owns TMEM lanes 32*(w%4).., warps 4-7 the upper half of the columns) and hand
// 4 consecutive columns at a time to the Epi functor.
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "gemm.cuh"

namespace r2d2 {

constexpr int UM_BM = 128;
constexpr int UM_BK = 64;
constexpr int UM_PRODUCERS = 256;            // 8 producer/epilogue warps
constexpr int UM_THREADS = UM_PRODUCERS + 32;  // + 1 MMA/TMEM warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded spin: a protocol bug traps (launch error) instead of hanging the GPU
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
// warp-uniform leader election (elect.sync): keeps the enclosing code warp-convergent for the compiler
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) | [32,46) SBO >> 4 (1024 B)
//   [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// issue only; the registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// the wait names the registers as in/out operands so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// split 2 floats into packed bf16x2 hi and lo
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __low2float(h), x1 - __high2float(h));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// byte offset of (row, k) inside a K-major SW128 bf16 tile (k multiple of 4 -> 8-byte aligned)
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + ((k & 7) << 1));
}

template <class F> struct Pair {   // two operand sets selected by blockIdx.z (online / target network)
    F f0, f1;
};
template <class F> __device__ __forceinline__ const F& sel_z(const F& f, int) { return f; }
template <class F> __device__ __forceinline__ const F& sel_z(const Pair<F>& f, int z) { return z ? f.f1 : f.f0; }
template <class F> struct IsPair { static constexpr bool value = false; };
template <class F> struct IsPair<Pair<F>> { static constexpr bool value = true; };

template <int BN, int TERMS> struct UmmaCfg {
    static constexpr int kTileA = UM_BM * UM_BK * 2;                     // bytes of one bf16 tile of A
    static constexpr int kTileB = BN * UM_BK * 2;
    static constexpr int kStageBytes = (TERMS == 3 ? 2 : 1) * (kTileA + kTileB);
    static constexpr int kStages = (kStageBytes * 4 <= 160 * 1024) ? 4 : ((kStageBytes * 3 <= 200 * 1024) ? 3 : 2);
    static constexpr int kSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

// ---- producer: fetch one k-block of an operand into registers, then split+store to smem
template <int ROWS, class Load> struct Stager {
    static constexpr bool kK = Load::kKMajor;
    // K-major: unit = (row, 8 k)        -> ROWS*8 units ; M-major: unit = (4 rows, 4 k) -> ROWS/4*16 units
    static constexpr int kUnits = kK ? ROWS * 8 : (ROWS / 4) * 16;
    static constexpr int kIters = (kUnits + UM_PRODUCERS - 1) / UM_PRODUCERS;
    static constexpr int kRegs = kK ? 8 : 16;
    float r[kIters][kRegs];

    __device__ __forceinline__ void fetch(const Load& ld, int row0, int kbase, int k_end, int tid) {
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int u = tid + it * UM_PRODUCERS;
            if (kUnits % UM_PRODUCERS != 0 && u >= kUnits) break;
            if constexpr (kK) {
                const int row = u >> 3, k = kbase + ((u & 7) << 3);
                float a[4], b[4];
                if (k < k_end) ld.load4(row0 + row, k, a); else zero4(a);
                if (k + 4 < k_end) ld.load4(row0 + row, k + 4, b); else zero4(b);
#pragma unroll
                for (int j = 0; j < 4; ++j) { r[it][j] = a[j]; r[it][4 + j] = b[j]; }
            } else {
                const int rq = u % (ROWS / 4), kq = u / (ROWS / 4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int k = kbase + kq * 4 + kk;
                    float a[4];
                    if (k < k_end) ld.load4(row0 + rq * 4, k, a); else zero4(a);
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[it][j * 4 + kk] = a[j];     // r[row j][k kk]
                }
            }
        }
    }
    template <int TERMS>
    __device__ __forceinline__ void store(uint8_t* hi_tile, uint8_t* lo_tile, int tid) const {
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int u = tid + it * UM_PRODUCERS;
            if (kUnits % UM_PRODUCERS != 0 && u >= kUnits) break;
            if constexpr (kK) {
                const int row = u >> 3, k = (u & 7) << 3;
                uint32_t h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split2(r[it][2 * j], r[it][2 * j + 1], h[j], l[j]);
                const uint32_t off = sw128_off(row, k);
                *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
                if (TERMS == 3) *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(l[0], l[1], l[2], l[3]);
            } else {
                const int rq = u % (ROWS / 4), kq = u / (ROWS / 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t h0, l0, h1, l1;
                    split2(r[it][j * 4 + 0], r[it][j * 4 + 1], h0, l0);
                    split2(r[it][j * 4 + 2], r[it][j * 4 + 3], h1, l1);
                    const uint32_t off = sw128_off(rq * 4 + j, kq * 4);
                    *reinterpret_cast<uint2*>(hi_tile + off) = make_uint2(h0, h1);
                    if (TERMS == 3) *reinterpret_cast<uint2*>(lo_tile + off) = make_uint2(l0, l1);
                }
            }
        }
    }
};

template <int BN, int TERMS, class ALoadT, class BLoadT, class EpiT>
__global__ void __launch_bounds__(UM_THREADS, 1)
umma_gemm_kernel(const ALoadT al_, const BLoadT bl_, const EpiT ep_, int K, int k_per_split) {
    using Cfg = UmmaCfg<BN, TERMS>;
    constexpr int kStages = Cfg::kStages;
    static_assert(BN == 16 || BN == 32 || BN == 64 || BN == 128 || BN == 256, "UMMA N / TMEM columns");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    // bars[0..kStages) full, [kStages..2kStages) empty, [2kStages] accumulator ready, then the TMEM slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.z;
    const auto& al = sel_z(al_, z);
    const auto& bl = sel_z(bl_, z);
    const auto& ep = sel_z(ep_, z);
    const int m0 = blockIdx.y * UM_BM, n0 = blockIdx.x * BN;
    const int k_begin = IsPair<ALoadT>::value ? 0 : z * k_per_split;
    const int k_end = IsPair<ALoadT>::value ? K : min(K, k_begin + k_per_split);
    const int nk = (k_end - k_begin + UM_BK - 1) / UM_BK;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&bars[s]), UM_PRODUCERS / 32);   // one arrive per producer warp
            mbar_init(smem_u32(&bars[kStages + s]), 1);         // tcgen05.commit
        }
        mbar_init(smem_u32(&bars[2 * kStages]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == UM_PRODUCERS / 32) {                            // TMEM allocation by the MMA warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(Cfg::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < UM_PRODUCERS / 32) {
        // ------------------------------------------------------------------ producers
        Stager<UM_BM, typename std::remove_cv<typename std::remove_reference<decltype(al)>::type>::type> sa;
        Stager<BN, typename std::remove_cv<typename std::remove_reference<decltype(bl)>::type>::type> sb;
        if (nk > 0) { sa.fetch(al, m0, k_begin, k_end, tid); sb.fetch(bl, n0, k_begin, k_end, tid); }
        for (int kt = 0; kt < nk; ++kt) {
            const int s = kt % kStages;
            const uint32_t ph = (uint32_t)(kt / kStages) & 1u;
            mbar_wait(smem_u32(&bars[kStages + s]), ph ^ 1u);            // slot free
            uint8_t* st = smem + s * Cfg::kStageBytes;
            uint8_t* a_hi = st;
            uint8_t* b_hi = st + Cfg::kTileA;
            uint8_t* a_lo = st + Cfg::kTileA + Cfg::kTileB;
            uint8_t* b_lo = a_lo + Cfg::kTileA;
            sa.template store<TERMS>(a_hi, a_lo, tid);
            sb.template store<TERMS>(b_hi, b_lo, tid);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[s]));
            if (kt + 1 < nk) {
                sa.fetch(al, m0, k_begin + (kt + 1) * UM_BK, k_end, tid);
                sb.fetch(bl, n0, k_begin + (kt + 1) * UM_BK, k_end, tid);
            }
        }
        // ------------------------------------------------------------------ epilogue
        if (nk > 0) {
            mbar_wait(smem_u32(&bars[2 * kStages]), 0);
            tc_fence_after();
        }
        const int row = (warp & 3) * 32 + lane;
        constexpr int kHalf = BN >= 32 ? BN / 2 : BN;            // warps 4-7 take the upper column half
        const int cbeg = (BN >= 32 && warp >= 4) ? kHalf : 0;
        const int cend = (BN >= 32) ? cbeg + kHalf : (warp >= 4 ? 0 : BN);
        for (int c = cbeg; c < cend; c += 16) {
            float v[16];
            if (nk > 0) tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c, v);
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4) ep.store4(m0 + row, n0 + c + j, &v[j], z);
        }
        tc_fence_before();
    } else {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BN);
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % kStages;
                const uint32_t ph = (uint32_t)(kt / kStages) & 1u;
                mbar_wait(smem_u32(&bars[s]), ph);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * Cfg::kStageBytes);
                const uint64_t a_hi = umma_desc_sw128(st), b_hi = umma_desc_sw128(st + Cfg::kTileA);
                const uint64_t a_lo = umma_desc_sw128(st + Cfg::kTileA + Cfg::kTileB);
                const uint64_t b_lo = umma_desc_sw128(st + 2 * Cfg::kTileA + Cfg::kTileB);
#pragma unroll
                for (int k = 0; k < UM_BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)(k * 32 >> 4);           // +32 B per K=16 step inside the 128 B atom
                    if (TERMS == 3) {
                        umma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, (kt | k) ? 1u : 0u);
                        umma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                        umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
                    } else {
                        umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (kt | k) ? 1u : 0u);
                    }
                }
                umma_commit(smem_u32(&bars[kStages + s]));                  // frees the smem slot when the MMAs retire
            }
            if (nk > 0) umma_commit(smem_u32(&bars[2 * kStages]));           // accumulator complete
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == UM_PRODUCERS / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
    }
}

template <int BN, int TERMS, class ALoad, class BLoad, class Epi>
static inline cudaError_t launch_umma(const ALoad& al, const BLoad& bl, const Epi& ep, int M, int N, int K, int splits,
                                      cudaStream_t s) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    using Cfg = UmmaCfg<BN, TERMS>;
    auto kern = umma_gemm_kernel<BN, TERMS, ALoad, BLoad, Epi>;
    static unsigned long long configured = 0;                       // bit per device ordinal
    {
        cudaError_t e = ensure_dynamic_smem(kern, Cfg::kSmem, &configured);
        if (e != cudaSuccess) return e;
    }
    int k_per_split = (K + splits - 1) / splits;
    k_per_split = (k_per_split + UM_BK - 1) / UM_BK * UM_BK;
    dim3 grid((N + BN - 1) / BN, (M + UM_BM - 1) / UM_BM, splits);
    kern<<<grid, UM_THREADS, Cfg::kSmem, s>>>(al, bl, ep, K, k_per_split);
    return cudaGetLastError();
}

}  // namespace r2d2
