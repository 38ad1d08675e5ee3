// Test entry points: plain-matrix instances of the GEMM kernels (umma2.cuh single CTA / cp.async, umma3.cuh CTA pairs / TMA)
// so descriptors, swizzle, pipeline and TMEM epilogue can be validated against a host fp64 product independently of the
// network, plus the hardware-semantics probes the kernels rely on.
#include "umma3.cuh"

namespace r2d2 {
int g_fast_math = 0;
extern int g_config_epoch;
}  // namespace r2d2

using namespace r2d2;

extern "C" {

/* Precision mode of the tensor-core path: 0 = strict (bf16x3 split products everywhere, default), 1 = fast (plain
 * bf16 products), 2 = balanced (hi+lo only for weight operands of the encoder contractions; recurrence, input
 * projection and head stay strict).  Returns the previous mode. */
int r2d2_set_fast_math(int mode) {
    ++g_config_epoch;
    int prev = g_fast_math;
    if (mode >= 0 && mode <= 2) g_fast_math = mode;
    return prev;
}

/* v2 test entry: operands given as bf16 hi/lo planes.  a_major/b_major as in r2d2_debug_gemm (1 = [K][rows]
 * storage, fed through MN-major descriptors).  C fp32 [splits][M][N]. */
int r2d2_debug_gemm2(int ubn, int a_major, int b_major, int M, int N, int K, const void* a_hi, const void* a_lo,
                     const void* b_hi, const void* b_lo, float* C, int splits, void* stream) {
    R2D2_REQUIRE(a_hi && a_lo && b_hi && b_lo && C && M > 0 && N > 0 && K > 0 && splits >= 1, "bad arguments");
    R2D2_REQUIRE(N % 8 == 0 && (a_major == 0 ? K % 8 == 0 : M % 8 == 0) && (b_major == 0 ? K % 8 == 0 : N % 8 == 0), "alignment");
    cudaStream_t s = as_stream(stream);
    const bf16 *ah = (const bf16*)a_hi, *al = (const bf16*)a_lo, *bh = (const bf16*)b_hi, *bl = (const bf16*)b_lo;
    Epi2Partial ep{C, M, N};
    cudaError_t e = cudaErrorInvalidValue;
#define R2D2_DBG2(UBN, AS, BS) e = launch_umma2<UBN>(AS, BS, ep, M, N, K, splits, s)
#define R2D2_DBG2_ALL(UBN)                                                                                      \
    if (a_major == 0 && b_major == 0) R2D2_DBG2(UBN, (SrcMatK{ah, al, M, K, K}), (SrcMatK{bh, bl, N, K, K}));     \
    else if (a_major == 1 && b_major == 0) R2D2_DBG2(UBN, (SrcMatMN{ah, al, M, K, M}), (SrcMatK{bh, bl, N, K, K}));
#define R2D2_DBG2_MNB(UBN)                                                                                      \
    if (a_major == 0 && b_major == 1) R2D2_DBG2(UBN, (SrcMatK{ah, al, M, K, K}), (SrcMatMN{bh, bl, N, K, N}));    \
    else if (a_major == 1 && b_major == 1) R2D2_DBG2(UBN, (SrcMatMN{ah, al, M, K, M}), (SrcMatMN{bh, bl, N, K, N}));
    switch (ubn) {
        case 16: R2D2_DBG2_ALL(16) break;
        case 32: R2D2_DBG2_ALL(32) break;
        case 64: R2D2_DBG2_ALL(64) R2D2_DBG2_MNB(64) break;
        case 128: R2D2_DBG2_ALL(128) R2D2_DBG2_MNB(128) break;
        default: R2D2_DBG2_ALL(256) R2D2_DBG2_MNB(256) break;
    }
#undef R2D2_DBG2
#undef R2D2_DBG2_ALL
#undef R2D2_DBG2_MNB
    R2D2_CUDA_CHECK(e);
    return R2D2_OK;
}

/* v3 test entry (CTA-pair, TMA-fed kernel of umma3.cuh): operands as bf16 hi/lo planes, majors as in r2d2_debug_gemm2.
 * C fp32 [splits][M][N], zero-initialised by the caller (fewer partials may be written than requested). */
int r2d2_debug_gemm3(int a_major, int b_major, int M, int N, int K, const void* a_hi, const void* a_lo, const void* b_hi,
                     const void* b_lo, float* C, int splits, void* stream) {
    R2D2_REQUIRE(a_hi && a_lo && b_hi && b_lo && C && M > 0 && N > 0 && K > 0 && splits >= 1, "bad arguments");
    R2D2_REQUIRE(N % 8 == 0 && (a_major == 0 ? K % 8 == 0 : M % 8 == 0) && (b_major == 0 ? K % 8 == 0 : N % 8 == 0), "alignment");
    cudaStream_t s = as_stream(stream);
    const Mat3 A{(const bf16*)a_hi, (const bf16*)a_lo, M, K, a_major ? M : K}, B{(const bf16*)b_hi, (const bf16*)b_lo, N, K, b_major ? N : K};
    Epi2Partial ep{C, M, N};
    cudaError_t e;
    if (!a_major && !b_major) e = launch_umma3<false, false>(A, B, ep, M, N, K, splits, s);
    else if (a_major && !b_major) e = launch_umma3<true, false>(A, B, ep, M, N, K, splits, s);
    else if (!a_major && b_major) e = launch_umma3<false, true>(A, B, ep, M, N, K, splits, s);
    else e = launch_umma3<true, true>(A, B, ep, M, N, K, splits, s);
    R2D2_CUDA_CHECK(e);
    return R2D2_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Experiment: can a K-major SWIZZLE_128B operand be addressed at a ROW-SHIFTED start (start + 128*s bytes) so that
// several im2col taps read overlapping windows of ONE staged activation buffer?  mode 0: base_offset field = 0,
// mode 1: base_offset = s & 7.  Data is stored with the absolute-address swizzle (chunk ^= physical_row & 7).
// D[128][32] = A[s .. s+128)[64] . B[32][64]^T, bf16 hi only.
// ------------------------------------------------------------------------------------------------
namespace r2d2 {
__global__ void __launch_bounds__(128) shift_probe_kernel(const bf16* __restrict__ A /*[144][64]*/, const bf16* __restrict__ Bm /*[32][64]*/,
                                                          float* __restrict__ D /*[128][32]*/, int shift, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t sA = raw + pad, sB = sA + 144 * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 144 * 128 + 32 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int u = tid; u < 144 * 8; u += 128) {
        const int row = u >> 3, j = u & 7;
        *reinterpret_cast<uint4*>(smem + (row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4)) =
            *reinterpret_cast<const uint4*>(A + row * 64 + j * 8);
    }
    for (int u = tid; u < 32 * 8; u += 128) {
        const int row = u >> 3, j = u & 7;
        *reinterpret_cast<uint4*>(smem + 144 * 128 + (row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4)) =
            *reinterpret_cast<const uint4*>(Bm + row * 64 + j * 8);
    }
    if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t start = sA + 128u * (uint32_t)shift;
        const uint64_t bdesc = umma_desc_sw128(sB);
        if (mode >= 3) {
            // MN-major, M = 128 as two OVERLAPPING 64-column atoms: atom 1 starts (mode - 2) lines after atom 0 (LBO = 128 B x lines)
            const uint64_t adesc = umma_desc_sw128_mn(start, 128u * (uint32_t)(mode - 2));
            const uint32_t idesc = ((1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24)) | (1u << 15);
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
        } else if (mode == 2) {
            // A read MN-major: smem rows are the REDUCTION index (64 + shift lines of 64 M-elements), M = 64
            const uint64_t adesc = umma_desc_sw128_mn(start, 8192);
            const uint32_t idesc = ((1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((64u >> 4) << 24)) | (1u << 15);
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
        } else {
            uint64_t adesc = umma_desc_sw128(start);
            if (mode == 1) adesc |= (uint64_t)(shift & 7) << 49;                  // base_offset, bits [49,52)
            const uint32_t idesc = umma_idesc_bf16(32);
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
        }
        umma_commit(smem_u32(bar));
    }
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    for (int c = 0; c < 32; c += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        for (int i = 0; i < 16; ++i) D[(warp * 32 + (tid & 31)) * 32 + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}
}  // namespace r2d2

extern "C" int r2d2_debug_shift_probe(const void* A, const void* B, float* D, int shift, int mode, void* stream) {
    R2D2_REQUIRE(A && B && D && shift >= 0 && shift <= 16 && mode >= 0 && mode <= 24, "bad arguments");
    const int smem = 144 * 128 + 32 * 128 + 1024 + 64;
    r2d2::shift_probe_kernel<<<1, 128, smem, r2d2::as_stream(stream)>>>((const r2d2::bf16*)A, (const r2d2::bf16*)B, D, shift, mode);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

// ------------------------------------------------------------------------------------------------
// tcgen05.mma issue-rate probe (SS mode, operands in shared memory): one CTA issues `reps` x 4 MMAs of M x N x 16
// back to back and reports clock64 cycles.  mode bit 0: alternate two accumulators; bit 1: A consumed MN-major.
namespace r2d2 {
__global__ void __launch_bounds__(128) mma_rate_kernel(int M, int N, int reps, int mode, long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t sA = raw + pad, sB = sA + 16384;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int u = tid; u < (16384 + 32768) / 4; u += 128) reinterpret_cast<uint32_t*>(smem)[u] = 0x3C003C00u;   // small finite bf16
    if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 0) {
        const bool leader = elect_one();
        const bool amn = mode & 2;
        const uint32_t idesc = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24)) | (amn ? (1u << 15) : 0u);
        const uint64_t a0 = amn ? umma_desc_sw128_mn(sA, 8192) : umma_desc_sw128(sA), b0 = umma_desc_sw128(sB);
        const uint64_t astep = amn ? 128 : 2;
        const uint32_t acc1 = (mode & 1) ? (uint32_t)N : 0u;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (leader) umma_bf16(tmem + ((k & 1) ? acc1 : 0u), a0 + astep * k, b0 + 2 * k, idesc, 1u);
        }
        if (leader) umma_commit(smem_u32(bar));
        __syncwarp();
        mbar_wait(smem_u32(bar), 0);
        long long t1 = clock64();
        if (leader && blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory"); }
}
}  // namespace r2d2

extern "C" int r2d2_debug_mma_rate(int M, int N, int reps, int mode, int ctas, long long* cycles, void* stream) {
    R2D2_REQUIRE((M == 64 || M == 128) && N >= 16 && N <= 256 && N % 16 == 0 && reps > 0 && ctas > 0 && cycles, "bad arguments");
    const int smem = 16384 + 32768 + 1024 + 64;
    static unsigned long long configured = 0;
    R2D2_CUDA_CHECK(r2d2::ensure_dynamic_smem(r2d2::mma_rate_kernel, smem, &configured));
    r2d2::mma_rate_kernel<<<ctas, 128, smem, r2d2::as_stream(stream)>>>(M, N, reps, mode, cycles);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

// ------------------------------------------------------------------------------------------------
// Hardware probe for the cluster recurrence (recurrence2.cuh): tcgen05.mma with the A operand read from TENSOR MEMORY
// (written there with tcgen05.st, two bf16 reduction indices per 32-bit column, even index in the low half) and the B
// operand in shared memory as K-major SWIZZLE_64B tiles of [16 rows][32 k].  D[128][16] = A[128][64] . B[16][64]^T.
namespace r2d2 {
__global__ void __launch_bounds__(128) ts_probe_kernel(const bf16* __restrict__ A /*[128][64]*/, const bf16* __restrict__ Bm /*[16][64]*/,
                                                       float* __restrict__ D /*[128][16]*/) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t sB = raw + pad;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2048);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int u = tid; u < 16 * 8; u += 128) {                 // 16 rows x 8 chunks of 8 k
        const int r = u >> 3, ch = u & 7, tile = ch >> 2, c4 = ch & 3;
        *reinterpret_cast<uint4*>(smem + tile * 1024 + (r >> 3) * 512 + (r & 7) * 64 + ((c4 ^ ((r >> 1) & 3)) << 4)) =
            *reinterpret_cast<const uint4*>(Bm + r * 64 + ch * 8);
    }
    if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    {   // A row `tid` -> TMEM lane tid, columns 0..31 (k = 2c, 2c+1)
        const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)tid * 64);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t r[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = arow[h * 16 + i];
            tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * 16), r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        for (int s = 0; s < 4; ++s)
            umma_bf16_ts(tmem + 32, tmem + (uint32_t)(8 * s), umma_desc_sw64(sB + (s >> 1) * 1024 + (s & 1) * 32), idesc, s ? 1u : 0u);
        umma_commit(smem_u32(bar));
    }
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 32u, v);
    for (int i = 0; i < 16; ++i) D[tid * 16 + i] = v[i];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}
}  // namespace r2d2

extern "C" int r2d2_debug_ts_probe(const void* A, const void* B, float* D, void* stream) {
    R2D2_REQUIRE(A && B && D, "bad arguments");
    r2d2::ts_probe_kernel<<<1, 128, 2048 + 1024 + 64, r2d2::as_stream(stream)>>>((const r2d2::bf16*)A, (const r2d2::bf16*)B, D);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}
