// Gather-GEMM building block:  C[m][n] = sum_k A(m,k) * B(n,k),  epilogue(m, n, acc).
//
// Every contraction of the R2D2 network pass (implicit-GEMM convolutions, FC, LSTM input
// projection / recurrence, dueling head, and all their dgrad/wgrad transposes) is an instance of
// this template with three functors:
//   ALoad / BLoad : how a 4-wide slice of the operand is fetched from HBM (im2col gather, strided
//                   matrix, u8 frames, row-gather through an index map ...).  kKMajor says which way
//                   the 4 consecutive elements run: along k (K-major source) or along the row index
//                   (row-major-in-m source, i.e. a transposed operand).
//   Epi           : what happens to 4 consecutive-n accumulators (bias, ReLU, masks, scatter,
//                   layout permutation, LSTM cell update, split-K partial store).
// This file holds the fp32 CUDA-core (FFMA) main loop: exact fp32 products, the numerical
// baseline every faster path is validated against.
#pragma once
#include "common.cuh"

namespace r2d2 {

// ----------------------------------------------------------------------------------------------
// main loop
// ----------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__((BM / 8) * (BN / 8) >= 64 ? (BM / 8) * (BN / 8) : 64)
gemm_ffma_kernel(const ALoad al, const BLoad bl, const Epi ep, int K, int k_per_split) {
    constexpr int TX = BN / 8, TY = BM / 8, NT = TX * TY;
    static_assert(NT >= 32 && BK % 4 == 0 && BM % 8 == 0 && BN % 8 == 0, "tile config");
    constexpr int LDA = BM + 4, LDB = BN + 4;
    constexpr int A_VECS = BM * BK / 4, B_VECS = BN * BK / 4;
    constexpr int A_IT = (A_VECS + NT - 1) / NT, B_IT = (B_VECS + NT - 1) / NT;
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];

    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kz = blockIdx.z;
    const int k_begin = kz * k_per_split;
    const int k_end = min(K, k_begin + k_per_split);
    const int nk = (k_end - k_begin + BK - 1) / BK;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[A_IT][4], rb[B_IT][4];

    auto fetch = [&](int kt) {
        const int kbase = k_begin + kt * BK;
#pragma unroll
        for (int it = 0; it < A_IT; ++it) {
            const int v = tid + it * NT;
            if (A_VECS % NT == 0 || v < A_VECS) {
                if constexpr (ALoad::kKMajor) {
                    const int row = v / (BK / 4), kq = v % (BK / 4);
                    const int k = kbase + kq * 4;
                    if (k < k_end) al.load4(m0 + row, k, ra[it]);
                    else { ra[it][0] = ra[it][1] = ra[it][2] = ra[it][3] = 0.f; }
                } else {
                    const int kk = v / (BM / 4), mq = v % (BM / 4);
                    const int k = kbase + kk;
                    if (k < k_end) al.load4(m0 + mq * 4, k, ra[it]);
                    else { ra[it][0] = ra[it][1] = ra[it][2] = ra[it][3] = 0.f; }
                }
            }
        }
#pragma unroll
        for (int it = 0; it < B_IT; ++it) {
            const int v = tid + it * NT;
            if (B_VECS % NT == 0 || v < B_VECS) {
                if constexpr (BLoad::kKMajor) {
                    const int row = v / (BK / 4), kq = v % (BK / 4);
                    const int k = kbase + kq * 4;
                    if (k < k_end) bl.load4(n0 + row, k, rb[it]);
                    else { rb[it][0] = rb[it][1] = rb[it][2] = rb[it][3] = 0.f; }
                } else {
                    const int kk = v / (BN / 4), nq = v % (BN / 4);
                    const int k = kbase + kk;
                    if (k < k_end) bl.load4(n0 + nq * 4, k, rb[it]);
                    else { rb[it][0] = rb[it][1] = rb[it][2] = rb[it][3] = 0.f; }
                }
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int it = 0; it < A_IT; ++it) {
            const int v = tid + it * NT;
            if (A_VECS % NT == 0 || v < A_VECS) {
                if constexpr (ALoad::kKMajor) {
                    const int row = v / (BK / 4), kq = v % (BK / 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) As[buf][kq * 4 + j][row] = ra[it][j];
                } else {
                    const int kk = v / (BM / 4), mq = v % (BM / 4);
                    *reinterpret_cast<float4*>(&As[buf][kk][mq * 4]) = make_float4(ra[it][0], ra[it][1], ra[it][2], ra[it][3]);
                }
            }
        }
#pragma unroll
        for (int it = 0; it < B_IT; ++it) {
            const int v = tid + it * NT;
            if (B_VECS % NT == 0 || v < B_VECS) {
                if constexpr (BLoad::kKMajor) {
                    const int row = v / (BK / 4), kq = v % (BK / 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) Bs[buf][kq * 4 + j][row] = rb[it][j];
                } else {
                    const int kk = v / (BN / 4), nq = v % (BN / 4);
                    *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = make_float4(rb[it][0], rb[it][1], rb[it][2], rb[it][3]);
                }
            }
        }
    };

    if (nk > 0) {
        fetch(0);
        stash(0);
    }
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) fetch(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][BM / 2 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][BN / 2 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) stash(buf ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
        ep.store4(m, n0 + tx * 4, &acc[i][0], kz);
        ep.store4(m, n0 + BN / 2 + tx * 4, &acc[i][4], kz);
    }
}

template <int BM, int BN, int BK, class ALoad, class BLoad, class Epi>
static inline cudaError_t launch_gemm(const ALoad& al, const BLoad& bl, const Epi& ep, int M, int N, int K,
                                      int splits, cudaStream_t s) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    constexpr int NT = (BM / 8) * (BN / 8);
    int k_per_split = (K + splits - 1) / splits;
    k_per_split = (k_per_split + BK - 1) / BK * BK;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
    gemm_ffma_kernel<BM, BN, BK, ALoad, BLoad, Epi><<<grid, NT, 0, s>>>(al, bl, ep, K, k_per_split);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------
// operand loaders
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void zero4(float (&v)[4]) { v[0] = v[1] = v[2] = v[3] = 0.f; }
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

// X(row, k) = p[row*ld + k]            (4 consecutive k)
struct MatK {
    static constexpr bool kKMajor = true;
    const float* p; int rows, K, ld;
    __device__ __forceinline__ void load4(int row, int k, float (&v)[4]) const {
        if (row < rows && k < K) ld4(p + (size_t)row * ld + k, v); else zero4(v);
    }
};
// X(row, k) = p[k*ld + row]            (4 consecutive rows) -- a transposed operand
struct MatM {
    static constexpr bool kKMajor = false;
    const float* p; int rows, K, ld;
    __device__ __forceinline__ void load4(int row, int k, float (&v)[4]) const {
        if (row < rows && k < K) ld4(p + (size_t)k * ld + row, v); else zero4(v);
    }
};
// X(r, k) = p[src[r]*ld + k], src[r] < 0 -> 0      (row gather through an index map)
struct RowGatherK {
    static constexpr bool kKMajor = true;
    const float* p; const int* src; int rows, K, ld;
    __device__ __forceinline__ void load4(int row, int k, float (&v)[4]) const {
        int s = (row < rows) ? __ldg(src + row) : -1;
        if (s >= 0 && k < K) ld4(p + (size_t)s * ld + k, v); else zero4(v);
    }
};
// X(j, r) = p[src[r]*ld + j]            (transposed use of a row-gathered matrix)
struct RowGatherM {
    static constexpr bool kKMajor = false;
    const float* p; const int* src; int rows, K, ld;   // rows = width limit (j), K = number of gathered rows
    __device__ __forceinline__ void load4(int row, int k, float (&v)[4]) const {
        int s = (k < K) ? __ldg(src + k) : -1;
        if (s >= 0 && row < rows) ld4(p + (size_t)s * ld + row, v); else zero4(v);
    }
};

// conv1 im2col over u8 frames: m = (frame, oy, ox) on the 20x20 output grid, k = (c, ky, kx), 8x8 stride 4.
// Frames are (C, 84, 84) u8; 4 consecutive kx are one aligned 32-bit load.
struct Conv1FrameK {
    static constexpr bool kKMajor = true;
    const uint8_t* obs; int C, nframes;
    __device__ __forceinline__ void load4(int m, int k, float (&v)[4]) const {
        if (m >= nframes * 400 || k >= C * 64) { zero4(v); return; }
        const int f = m / 400, p = m - f * 400, oy = p / 20, ox = p - oy * 20;
        const int c = k >> 6, ky = (k >> 3) & 7, kx = k & 7;
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(
            obs + ((size_t)f * C + c) * 7056 + (4 * oy + ky) * 84 + 4 * ox + kx));
        v[0] = (float)(w & 255u); v[1] = (float)((w >> 8) & 255u);
        v[2] = (float)((w >> 16) & 255u); v[3] = (float)(w >> 24);
    }
};
// the same gather used as the transposed operand of conv1's wgrad: row = k index (4 consecutive kx), k = m
struct Conv1FrameM {
    static constexpr bool kKMajor = false;
    const uint8_t* obs; int C, nframes;
    __device__ __forceinline__ void load4(int row, int k, float (&v)[4]) const {
        Conv1FrameK g{obs, C, nframes};
        g.load4(k, row, v);
    }
};

// NHWC im2col:  m = (frame, oy, ox),  k = (ky, kx, c)   (4 consecutive channels)
template <int IH, int IW, int IC, int OH, int OW, int KH, int KW, int S>
struct ConvNHWC_K {
    static constexpr bool kKMajor = true;
    const float* in; int nframes;
    __device__ __forceinline__ void load4(int m, int k, float (&v)[4]) const {
        if (m >= nframes * OH * OW || k >= KH * KW * IC) { zero4(v); return; }
        const int f = m / (OH * OW), p = m - f * (OH * OW), oy = p / OW, ox = p - oy * OW;
        const int tap = k / IC, c = k - tap * IC, ky = tap / KW, kx = tap - ky * KW;
        ld4(in + (((size_t)f * IH + S * oy + ky) * IW + S * ox + kx) * IC + c, v);
    }
};
template <int IH, int IW, int IC, int OH, int OW, int KH, int KW, int S>
struct ConvNHWC_M {   // transposed use (wgrad B operand): row = k index, k = m
    static constexpr bool kKMajor = false;
    const float* in; int nframes;
    __device__ __forceinline__ void load4(int row, int k, float (&v)[4]) const {
        ConvNHWC_K<IH, IW, IC, OH, OW, KH, KW, S> g{in, nframes};
        g.load4(k, row, v);
    }
};
// dgrad gather of a stride-1 or parity-decomposed stride-2 conv:
//   m = (frame, y', x') on a GH x GW grid,  k = (jy, jx, c_out),  value = dout[f][y'-jy][x'-jx][c_out]
template <int GH, int GW, int OH, int OW, int OC, int JH, int JW>
struct ConvDgradK {
    static constexpr bool kKMajor = true;
    const float* dout; int nframes;
    __device__ __forceinline__ void load4(int m, int k, float (&v)[4]) const {
        if (m >= nframes * GH * GW || k >= JH * JW * OC) { zero4(v); return; }
        const int f = m / (GH * GW), p = m - f * (GH * GW), y = p / GW, x = p - y * GW;
        const int tap = k / OC, c = k - tap * OC, jy = tap / JW, jx = tap - jy * JW;
        const int oy = y - jy, ox = x - jx;
        if (oy < 0 || oy >= OH || ox < 0 || ox >= OW) { zero4(v); return; }
        ld4(dout + (((size_t)f * OH + oy) * OW + ox) * OC + c, v);
    }
};

// ----------------------------------------------------------------------------------------------
// epilogues
// ----------------------------------------------------------------------------------------------
// out[m*ld + n] = act(acc*scale + bias[n])
template <bool kRelu>
struct EpiBias {
    float* out; const float* bias; int M, N, ld; float scale;
    __device__ __forceinline__ void store4(int m, int n, const float* a, int) const {
        if (m >= M || n >= N) return;
        float4 r;
        float* rr = reinterpret_cast<float*>(&r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x = a[j] * scale + (bias ? __ldg(bias + n + j) : 0.f);
            rr[j] = kRelu ? fmaxf(x, 0.f) : x;
        }
        *reinterpret_cast<float4*>(out + (size_t)m * ld + n) = r;
    }
};
// split-K partial: ws[z][m][n] = acc
struct EpiPartial {
    float* ws; int M, N;
    __device__ __forceinline__ void store4(int m, int n, const float* a, int z) const {
        if (m >= M || n >= N) return;
        *reinterpret_cast<float4*>(ws + ((size_t)z * M + m) * N + n) = make_float4(a[0], a[1], a[2], a[3]);
    }
};
// dgrad with ReLU mask taken from the saved post-activation tensor laid out like the output
struct EpiMasked {
    float* out; const float* act; int M, N, ld;
    __device__ __forceinline__ void store4(int m, int n, const float* a, int) const {
        if (m >= M || n >= N) return;
        const float4 y = *reinterpret_cast<const float4*>(act + (size_t)m * ld + n);
        *reinterpret_cast<float4*>(out + (size_t)m * ld + n) =
            make_float4(y.x > 0.f ? a[0] : 0.f, y.y > 0.f ? a[1] : 0.f, y.z > 0.f ? a[2] : 0.f, y.w > 0.f ? a[3] : 0.f);
    }
};

}  // namespace r2d2
