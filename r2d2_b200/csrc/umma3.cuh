// tcgen05 GEMM, v3 data path: CTA PAIRS (cta_group::2) fed by TMA.
//
//   D[m][n] = sum_k A(m,k) * B(n,k)          pair tile 256 x 256, 64 reduction indices per stage
//
// Why pairs.  A bf16x3 product needs both bf16 planes (hi, lo) of both operands: 4 bytes per element, and three
// MMAs per k-step.  A single-CTA 128 x 128 tile (umma2.cuh) then needs 64 KB of operands per 768 tensor-clocks =
// 85 B/clk/SM, twice what an SM can pull from L2 (~40 B/clk measured): those GEMMs ran at 22-38 % tensor-pipe
// activity.  With cta_group::2 the two SMs of a TPC compute ONE 256 x 256 tile: each CTA stages only its own 128
// rows of A and its own 128 rows of B (64 KB per stage, as before) but the pair issues M = 256, N = 256 MMAs
// (1536 clocks per stage on each SM): 42 B/clk/SM -- operand traffic per FLOP is halved.
//
// Operands are plain matrices already split into bf16 hi/lo planes (see umma2.cuh), described by 2-D tensor maps:
//   K-major  operand: matrix [rows][ld], box {64 k, 128 rows}          -> 128 smem lines of 128 B  (one TMA per plane)
//   MN-major operand: matrix [K][ld] (rows contiguous), box {64 rows, 64 k} -> two 64-row atoms 8 KB apart (two TMAs)
// both with the 128-byte swizzle the UMMA descriptors of umma2.cuh expect; out-of-range rows / reduction indices
// are zero-filled by the TMA unit, so ragged M, N, K need no predicates.
//
// Roles (320 threads per CTA, both CTAs of the pair run the same code):
//   warp 0   one lane issues the TMA loads of its CTA's half; every load completes on the LEADER CTA's full[] barrier
//   warp 1   TMEM allocation (cta_group::2); in the leader CTA it issues the MMAs for the pair and commits
//            (multicast to both CTAs) on empty[] / accumulator-full barriers
//   warps 2-9  epilogue: TMEM -> registers -> Epi functor (the same functors as umma2.cuh); two 256-column
//            accumulators alternate so the epilogue of tile i overlaps the main loop of tile i+1 (persistent pairs).
#pragma once
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "umma2.cuh"

namespace r2d2 {

constexpr int U3_EPI_WARPS = 8;
constexpr int U3_THREADS = 64 + 32 * U3_EPI_WARPS;
constexpr uint32_t U3_PEER_MASK = 0xFEFFFFFFu;         // clears the CTA-rank bit of a shared::cluster address -> CTA 0 of the pair
constexpr int U3_PLANE = 128 * 64 * 2;                // one bf16 plane of one operand tile: 16 KB

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t mbar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t mbar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {          // arrives on the barrier at this offset in BOTH CTAs
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

// cluster / tensor-memory helpers shared with recurrence2.cuh
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "r"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {      // K-major, 64-byte rows, 8-row groups 512 B apart
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <bool A_LO, bool B_LO> struct Umma3Cfg {
    static constexpr int kOffALo = U3_PLANE;
    static constexpr int kOffBHi = (A_LO ? 2 : 1) * U3_PLANE;
    static constexpr int kOffBLo = kOffBHi + U3_PLANE;
    static constexpr int kStage = ((A_LO ? 2 : 1) + (B_LO ? 2 : 1)) * U3_PLANE;
    static constexpr int kStages = (200 * 1024 / kStage) > 6 ? 6 : (200 * 1024 / kStage);     // 3 (strict) .. 6
    static constexpr int kSmem = kStages * kStage + 1024 + 256;
};

template <bool A_MN, bool B_MN, bool A_LO, bool B_LO, class EpiT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(U3_THREADS, 1)
umma3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmBh,
             const __grid_constant__ CUtensorMap tmBl, const EpiT ep, int K, int k_per_split, int m_tiles, int n_tiles, int splits) {
    using Cfg = Umma3Cfg<A_LO, B_LO>;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t smem_base = raw + pad;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStage);      // full[S] | empty[S] | accf[2] | acce[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int tiles_mn = m_tiles * n_tiles, ntiles = tiles_mn * splits;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[S + s]), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bars[2 * S + a]), 1); mbar_init(smem_u32(&bars[2 * S + 2 + a]), 2 * U3_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmBh);
        if (A_LO) tma_prefetch_desc(&tmAl);
        if (B_LO) tma_prefetch_desc(&tmBl);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (one lane)
        if (lane == 0) {
            long long it = 0;
            for (int tile = pair; tile < ntiles; tile += npairs) {
                const int z = tile / tiles_mn, r = tile - z * tiles_mn, mt = r / n_tiles, nt = r - mt * n_tiles;
                const int m0 = mt * 256 + (int)rank * 128, nb0 = nt * 256 + (int)rank * 128;
                const int k_begin = z * k_per_split, k_end = min(K, k_begin + k_per_split);
                const int nk = (k_end - k_begin + 63) >> 6;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = (int)(it % S);
                    const uint32_t ph = (uint32_t)(it / S) & 1u;
                    mbar_wait(smem_u32(&bars[S + s]), ph ^ 1u);
                    const uint32_t st = smem_base + s * Cfg::kStage;
                    const uint32_t full = smem_u32(&bars[s]) & U3_PEER_MASK;            // the LEADER's barrier counts both halves
                    if (rank == 0) mbar_arrive_expect_tx(smem_u32(&bars[s]), 2u * Cfg::kStage);
                    const int k0 = k_begin + kb * 64;
                    if constexpr (!A_MN) {
                        tma_load_2d_pair(st, &tmAh, full, k0, m0);
                        if (A_LO) tma_load_2d_pair(st + Cfg::kOffALo, &tmAl, full, k0, m0);
                    } else {
#pragma unroll
                        for (int at = 0; at < 2; ++at) {
                            tma_load_2d_pair(st + at * 8192, &tmAh, full, m0 + 64 * at, k0);
                            if (A_LO) tma_load_2d_pair(st + Cfg::kOffALo + at * 8192, &tmAl, full, m0 + 64 * at, k0);
                        }
                    }
                    if constexpr (!B_MN) {
                        tma_load_2d_pair(st + Cfg::kOffBHi, &tmBh, full, k0, nb0);
                        if (B_LO) tma_load_2d_pair(st + Cfg::kOffBLo, &tmBl, full, k0, nb0);
                    } else {
#pragma unroll
                        for (int at = 0; at < 2; ++at) {
                            tma_load_2d_pair(st + Cfg::kOffBHi + at * 8192, &tmBh, full, nb0 + 64 * at, k0);
                            if (B_LO) tma_load_2d_pair(st + Cfg::kOffBLo + at * 8192, &tmBl, full, nb0 + 64 * at, k0);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issue (leader CTA; whole warp loops, one lane issues)
        if (rank == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u) |
                                       ((256u >> 3) << 17) | ((256u >> 4) << 24);
            const bool leader = elect_one();
            const uint32_t u_smem = __shfl_sync(0xffffffffu, smem_base, 0), u_tmem = __shfl_sync(0xffffffffu, tmem_base, 0);
            long long it = 0;
            int ti = 0;
            for (int tile = pair; tile < ntiles; tile += npairs, ++ti) {
                const int z = tile / tiles_mn;
                const int k_begin = z * k_per_split, k_end = min(K, k_begin + k_per_split);
                const int nk = (k_end - k_begin + 63) >> 6;
                const int a = ti & 1;
                mbar_wait(smem_u32(&bars[2 * S + 2 + a]), (((uint32_t)ti >> 1) & 1u) ^ 1u);     // both CTAs drained this accumulator
                tc_fence_after();
                const uint32_t acc = u_tmem + (uint32_t)(a * 256);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = (int)(it % S);
                    const uint32_t ph = (uint32_t)(it / S) & 1u;
                    mbar_wait(smem_u32(&bars[s]), ph);
                    tc_fence_after();
                    const uint32_t st = u_smem + s * Cfg::kStage;
                    auto mk_a = [&](uint32_t addr) { return A_MN ? umma_desc_sw128_mn(addr, 8192) : umma_desc_sw128(addr); };
                    auto mk_b = [&](uint32_t addr) { return B_MN ? umma_desc_sw128_mn(addr, 8192) : umma_desc_sw128(addr); };
                    const uint64_t a_hi = mk_a(st), a_lo = mk_a(st + Cfg::kOffALo), b_hi = mk_b(st + Cfg::kOffBHi), b_lo = mk_b(st + Cfg::kOffBLo);
                    constexpr uint64_t kAdvA = (A_MN ? 2048 : 32) >> 4, kAdvB = (B_MN ? 2048 : 32) >> 4;
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = kAdvA * k, db = kAdvB * k;
                            uint32_t accum = (kb | k) ? 1u : 0u;
                            if (A_LO) { umma_bf16_pair(acc, a_lo + da, b_hi + db, idesc, accum); accum = 1u; }
                            if (B_LO) { umma_bf16_pair(acc, a_hi + da, b_lo + db, idesc, accum); accum = 1u; }
                            umma_bf16_pair(acc, a_hi + da, b_hi + db, idesc, accum);
                        }
                        umma_commit_pair(smem_u32(&bars[S + s]));                        // stage s is free in both CTAs
                        if (kb == nk - 1) umma_commit_pair(smem_u32(&bars[2 * S + a]));  // accumulator complete in both CTAs
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue: 8 warps = 4 lane quadrants x 2 column halves
        const int q = warp & 3, half = (warp - 2) >> 2;
        int ti = 0;
        for (int tile = pair; tile < ntiles; tile += npairs, ++ti) {
            const int z = tile / tiles_mn, r = tile - z * tiles_mn, mt = r / n_tiles, nt = r - mt * n_tiles;
            const int row = mt * 256 + (int)rank * 128 + q * 32 + lane, n0 = nt * 256 + half * 128;
            const int a = ti & 1;
            mbar_wait(smem_u32(&bars[2 * S + a]), ((uint32_t)ti >> 1) & 1u);
            tc_fence_after();
            const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + half * 128);
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 64) {
                uint32_t rv[4][16];
#pragma unroll
                for (int g = 0; g < 4; ++g) tmem_ld16_issue(lane_base + c0 + g * 16, rv[g]);
#pragma unroll
                for (int g = 0; g < 4; ++g) tmem_ld_wait(rv[g]);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rv[g][i]);
                    ep.store16(row, n0 + c0 + g * 16, v, z);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(smem_u32(&bars[2 * S + 2 + a]) & U3_PEER_MASK);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // neither CTA frees TMEM (or exits) while the peer's MMAs / arrivals may still target it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (PFN_tmapEncodeTiled)p;
    }();
    return fn;
}

// 2-D bf16 tensor [outer][ld] of which [outer][inner] is valid; box {box_inner, box_outer}; 128-byte swizzle; OOB -> zeros.
// Cached per (pointer, shape, box): the workspaces of a net handle never move.
static inline const CUtensorMap* tmap_2d(const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
    using Key = std::tuple<const void*, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t>;
    static std::map<Key, CUtensorMap> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const Key key{base, inner, outer, ld, box_inner, box_outer};
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return nullptr;
    CUtensorMap m;
    const cuuint64_t dims[2] = {inner, outer};
    const cuuint64_t strides[1] = {ld * sizeof(bf16)};
    const cuuint32_t box[2] = {box_inner, box_outer};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nullptr;
    return &cache.emplace(key, m).first->second;
}
// plain split matrix operand: K-major = [rows][ld] (K valid columns), MN-major = [K][ld] (rows valid columns)
struct Mat3 { const bf16* hi; const bf16* lo; int rows, K; long long ld; };

template <bool A_MN, bool B_MN, bool A_LO, bool B_LO, class Epi>
static inline cudaError_t launch_umma3_inst(const Mat3& A, const Mat3& B, const Epi& ep, int M, int N, int K, int splits, cudaStream_t s) {
    using Cfg = Umma3Cfg<A_LO, B_LO>;
    auto kern = umma3_kernel<A_MN, B_MN, A_LO, B_LO, Epi>;
    static unsigned long long configured = 0;
    {
        cudaError_t e = ensure_dynamic_smem(kern, Cfg::kSmem, &configured);
        if (e != cudaSuccess) return e;
    }
    auto mk = [&](const bf16* p, const Mat3& X, bool mn) {
        return mn ? tmap_2d(p, (uint64_t)X.rows, (uint64_t)X.K, (uint64_t)X.ld, 64, 64) : tmap_2d(p, (uint64_t)X.K, (uint64_t)X.rows, (uint64_t)X.ld, 64, 128);
    };
    const CUtensorMap* ah = mk(A.hi, A, A_MN);
    const CUtensorMap* al = A_LO ? mk(A.lo, A, A_MN) : ah;
    const CUtensorMap* bh = mk(B.hi, B, B_MN);
    const CUtensorMap* bl = B_LO ? mk(B.lo, B, B_MN) : bh;
    if (!ah || !al || !bh || !bl) return cudaErrorInvalidValue;
    int k_per_split = (K + splits - 1) / splits;
    k_per_split = (k_per_split + 63) / 64 * 64;
    splits = (K + k_per_split - 1) / k_per_split;                       // every split owns at least one k-block
    const int m_tiles = (M + 255) / 256, n_tiles = (N + 255) / 256;
    const int ntiles = m_tiles * n_tiles * splits;
    const int pairs = ntiles < kNumSMs / 2 ? ntiles : kNumSMs / 2;
    kern<<<2 * pairs, U3_THREADS, Cfg::kSmem, s>>>(*ah, *al, *bh, *bl, ep, K, k_per_split, m_tiles, n_tiles, splits);
    return cudaGetLastError();
}

// number of split-K partials launch_umma3 will actually write for (K, splits)
static inline int umma3_effective_splits(int K, int splits) {
    int k_per_split = (K + splits - 1) / splits;
    k_per_split = (k_per_split + 63) / 64 * 64;
    return (K + k_per_split - 1) / k_per_split;
}

template <bool A_MN, bool B_MN, int POL = LO_STRICT, class Epi>
static inline cudaError_t launch_umma3(const Mat3& A, const Mat3& B, const Epi& ep, int M, int N, int K, int splits, cudaStream_t s) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    if (g_fast_math == 1) return launch_umma3_inst<A_MN, B_MN, false, false>(A, B, ep, M, N, K, splits, s);
    if (g_fast_math == 2 && POL == LO_WEIGHT_B) return launch_umma3_inst<A_MN, B_MN, false, true>(A, B, ep, M, N, K, splits, s);
    if (g_fast_math == 2 && POL == LO_NO_WEIGHT) return launch_umma3_inst<A_MN, B_MN, false, false>(A, B, ep, M, N, K, splits, s);
    return launch_umma3_inst<A_MN, B_MN, true, true>(A, B, ep, M, N, K, splits, s);
}

}  // namespace r2d2
