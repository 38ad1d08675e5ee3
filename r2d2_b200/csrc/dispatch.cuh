// Backend switch for the gather-GEMM: the tcgen05 path is the product; the fp32 FFMA path is kept
// as the on-device numerical reference the tests compare it against (r2d2_set_gemm_backend).
#pragma once
#include "umma.cuh"

namespace r2d2 {

enum GemmBackend { GEMM_FFMA = 0, GEMM_UMMA_BF16X3 = 1, GEMM_UMMA_BF16 = 2 };
extern int g_gemm_backend;

template <int FBM, int FBN, int UBN, class A, class B, class E>
static inline cudaError_t run_gemm(const A& a, const B& b, const E& e, int M, int N, int K, int splits, cudaStream_t s) {
    switch (g_gemm_backend) {
        case GEMM_FFMA: return launch_gemm<FBM, FBN, 16>(a, b, e, M, N, K, splits, s);
        case GEMM_UMMA_BF16X3: return launch_umma<UBN, 3>(a, b, e, M, N, K, splits, s);
        default: return launch_umma<UBN, 1>(a, b, e, M, N, K, splits, s);
    }
}

// the same contraction for two independent operand sets (online / target) in one launch
template <int FBM, int FBN, int UBN, class A, class B, class E>
static inline cudaError_t run_gemm_pair(const A& a0, const B& b0, const E& e0, const A& a1, const B& b1, const E& e1, int M, int N,
                                        int K, cudaStream_t s) {
    if (g_gemm_backend == GEMM_FFMA) {
        cudaError_t err = launch_gemm<FBM, FBN, 16>(a0, b0, e0, M, N, K, 1, s);
        if (err != cudaSuccess) return err;
        return launch_gemm<FBM, FBN, 16>(a1, b1, e1, M, N, K, 1, s);
    }
    Pair<A> pa{a0, a1};
    Pair<B> pb{b0, b1};
    Pair<E> pe{e0, e1};
    if (g_gemm_backend == GEMM_UMMA_BF16X3) return launch_umma<UBN, 3>(pa, pb, pe, M, N, K, 2, s);
    return launch_umma<UBN, 1>(pa, pb, pe, M, N, K, 2, s);
}

}  // namespace r2d2
