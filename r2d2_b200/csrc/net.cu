// K1 / K1b: sequence-unroll forward and BPTT backward of the R2D2 network
// (model.py:27-150 of the reference: conv encoder -> LSTM -> dueling head).
//
// One `r2d2_net` handle owns the packed weights, saved activations and backward scratch of the
// online (slot 0) and target (slot 1) networks for a fixed (B, T, C, A) batch shape.
//
// Data layout in HBM (per slot, NF = B*T frames, frame index f = b*T + t as in the batch):
//   act1 [NF][20][20][32]  act2 [NF][9][9][64]  act3 [NF][7][7][64]      NHWC fp32, post-ReLU
//   U    [T][B][KU]        LSTM input rows, time-major: latent(512) | one-hot last action(A) |
//                          last reward | zero pad to KU = roundup(512+A+1, 16)
//   XP   [T][B][4H]        input projection incl. both biases, gate-interleaved (col = 4*j + gate)
//   Hs, Cs [T][B][H]       hidden / cell state after step t (frozen past a sequence's length)
//   Gs   [T][B][4H]        post-nonlinearity gates (i,f,g,o interleaved), saved for BPTT
// Weights are kept by the caller in the reference's state_dict layout (one flat fp32 buffer, see
// r2d2_net_param_layout); `pack` re-lays them out for NHWC im2col order / gate interleave.
//
// The online pass is run ONCE for b+l+f steps: the reference's pass 1 (calculate_q_, no grad) and
// pass 3 (calculate_q, grad) share every hidden state up to b+l-1, so Q at the learning positions
// and at the n-step-shifted positions are two row gathers of the same unroll.
#include <map>

#include "dispatch.cuh"

namespace r2d2 {

constexpr int H = 512;          // config.hidden_dim (model.py:28); the kernels are specialised for it
constexpr int G4 = 4 * H;
constexpr int LATENT = 512;
constexpr int FLAT3 = 3136;     // 7*7*64
constexpr int NPARAM = 20;
constexpr int kRecSplits = 8;   // split-K of the BPTT recurrence GEMM (K = 2048) across CTAs

enum ParamId {
    P_C1W, P_C1B, P_C2W, P_C2B, P_C3W, P_C3B, P_FCW, P_FCB, P_WIH, P_WHH, P_BIH, P_BHH,
    P_A0W, P_A0B, P_A2W, P_A2B, P_V0W, P_V0B, P_V2W, P_V2B
};

static void param_sizes(int A, int C, int64_t* n) {
    n[P_C1W] = 32ll * C * 64; n[P_C1B] = 32; n[P_C2W] = 64 * 512; n[P_C2B] = 64; n[P_C3W] = 64 * 576; n[P_C3B] = 64;
    n[P_FCW] = 512ll * FLAT3; n[P_FCB] = 512; n[P_WIH] = (int64_t)G4 * (LATENT + A + 1); n[P_WHH] = (int64_t)G4 * H;
    n[P_BIH] = G4; n[P_BHH] = G4; n[P_A0W] = H * H; n[P_A0B] = H; n[P_A2W] = (int64_t)A * H; n[P_A2B] = A;
    n[P_V0W] = H * H; n[P_V0B] = H; n[P_V2W] = H; n[P_V2B] = 1;
}

struct Packed {
    float *W2p, *W3p, *Wfcp, *Wih_p, *Whh_p, *bias_p, *Wh0, *bh0, *W3d, *W2d;  // W2d: [4][32][256]
};
struct Acts {
    float *act1, *act2, *act3, *U, *XP, *Hs, *Cs, *Gs, *hid;   // hid: [2*Rmax][1024]
};

}  // namespace r2d2

struct r2d2_net {
    int B, T, C, A, Lmax, F, KU, NF, Rmax, KIH;
    int64_t off[r2d2::NPARAM + 1];
    r2d2::Packed pk[2];
    r2d2::Acts ac[2];
    // per-batch metadata (device)
    int *row_src, *len_full, *len_learn, *d_rows;     // row_src: [2*Rmax]  (q rows | shifted rows)
    // backward scratch
    float *dH, *DG, *dhrec, *dcrec, *dlat, *dpre3, *dpre2, *dpre1, *dhid, *dout16, *ws, *colws;
    size_t ws_floats;
    // last forward inputs (device pointers, caller-owned, must stay valid until backward)
    const uint8_t* obs; const float* hidden;
};

namespace r2d2 {

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void pack_kernel(const float* __restrict__ p, const int64_t* __restrict__ off, Packed pk, int A, int KU) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int KIH = LATENT + A + 1;
    // conv2: [n][c][ky][kx] -> [n][(ky*4+kx)*32 + c]
    if (i < 64 * 512) {
        const int n = i / 512, k = i % 512, tap = k >> 5, c = k & 31;
        pk.W2p[i] = p[off[P_C2W] + n * 512 + c * 16 + tap];
        // dgrad, 4 parity classes: W2d[cls][c_in][(jy*2+jx)*64 + c_out] = W2[c_out][c_in][py+2jy][px+2jx]
        const int cls = i / (32 * 256), r = i % (32 * 256), ci = r / 256, kk = r % 256, j = kk >> 6, co = kk & 63;
        const int py = cls >> 1, px = cls & 1, jy = j >> 1, jx = j & 1;
        pk.W2d[i] = p[off[P_C2W] + co * 512 + ci * 16 + (py + 2 * jy) * 4 + (px + 2 * jx)];
    }
    if (i < 64 * 576) {
        const int n = i / 576, k = i % 576, tap = k >> 6, c = k & 63;
        pk.W3p[i] = p[off[P_C3W] + n * 576 + c * 9 + tap];
        // dgrad: W3d[c_in][(ky*3+kx)*64 + c_out] = W3[c_out][c_in][ky][kx]
        pk.W3d[i] = p[off[P_C3W] + c * 576 + n * 9 + tap];
    }
    if (i < 512ll * FLAT3) {                      // fc: col c*49+hw -> hw*64+c
        const int n = i / FLAT3, k = i % FLAT3, hw = k >> 6, c = k & 63;
        pk.Wfcp[i] = p[off[P_FCW] + (int64_t)n * FLAT3 + c * 49 + hw];
    }
    if (i < (int64_t)G4 * KU) {                   // W_ih: gate-interleaved rows, zero-padded cols
        const int np = i / KU, k = i % KU, g = np & 3, j = np >> 2;
        pk.Wih_p[i] = (k < KIH) ? p[off[P_WIH] + (int64_t)(g * H + j) * KIH + k] : 0.f;
    }
    if (i < (int64_t)G4 * H) {
        const int np = i / H, k = i % H, g = np & 3, j = np >> 2;
        pk.Whh_p[i] = p[off[P_WHH] + (int64_t)(g * H + j) * H + k];
    }
    if (i < G4) {
        const int g = i & 3, j = i >> 2;
        pk.bias_p[i] = p[off[P_BIH] + g * H + j] + p[off[P_BHH] + g * H + j];
    }
    if (i < 2 * H * H) {                          // head layer 0: advantage | value stacked
        pk.Wh0[i] = (i < H * H) ? p[off[P_A0W] + i] : p[off[P_V0W] + i - H * H];
    }
    if (i < 2 * H) pk.bh0[i] = (i < H) ? p[off[P_A0B] + i] : p[off[P_V0B] + i - H];
}

// row maps + sequence lengths.  One CTA.  model.py:102-111 (shifted rows) and model.py:143.
__global__ void prep_rows_kernel(const uint8_t* __restrict__ burn, const uint8_t* __restrict__ learn,
                                 const uint8_t* __restrict__ fwd, int B, int F, int Rmax, int* __restrict__ row_src,
                                 int* __restrict__ len_full, int* __restrict__ len_learn, int* __restrict__ d_rows) {
    extern __shared__ int s_off[];
    for (int i = threadIdx.x; i < 2 * Rmax; i += blockDim.x) row_src[i] = -1;
    if (threadIdx.x == 0) {
        int run = 0;
        for (int n = 0; n < B; ++n) { s_off[n] = run; run += learn[n]; }
        s_off[B] = run;
        *d_rows = run;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < B; n += blockDim.x) {
        const int b = burn[n], l = learn[n], f = fwd[n];
        len_full[n] = b + l + f;
        len_learn[n] = b + l;
        for (int i = 0; i < l; ++i) {
            row_src[s_off[n] + i] = (b + i) * B + n;
            row_src[Rmax + s_off[n] + i] = min(b + F + i, b + l + f - 1) * B + n;
        }
    }
}

// U side columns: one-hot last action, last reward, zero pad  (model.py:92)
__global__ void side_columns_kernel(float* __restrict__ U, const uint8_t* __restrict__ last_action,
                                    const float* __restrict__ last_reward, int B, int T, int A, int KU) {
    const int row = blockIdx.x;                 // time-major row t*B + b
    const int t = row / B, b = row % B;
    const int f = b * T + t;
    for (int k = LATENT + threadIdx.x; k < KU; k += blockDim.x) {
        float v = 0.f;
        if (k < LATENT + A) v = last_action[(size_t)f * A + (k - LATENT)] ? 1.f : 0.f;
        else if (k == LATENT + A) v = last_reward[f];
        U[(size_t)row * KU + k] = v;
    }
}

// dueling output layer: one warp per row.  model.py:115-117
__global__ void head_out_kernel(const float* __restrict__ hid, const float* __restrict__ Wa2,
                                const float* __restrict__ ba2, const float* __restrict__ Wv2,
                                const float* __restrict__ bv2, int rows_cap, int A, float* __restrict__ q_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows_cap) return;
    const float* ha = hid + (size_t)warp * 2 * H;
    const float* hv = ha + H;
    float adv[32];
    float sum = 0.f;
    for (int a = 0; a < A; ++a) {
        float s = 0.f;
        for (int k = lane; k < H; k += 32) s = fmaf(ha[k], __ldg(Wa2 + a * H + k), s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        adv[a] = s + ba2[a];
        sum += adv[a];
    }
    float v = 0.f;
    for (int k = lane; k < H; k += 32) v = fmaf(hv[k], __ldg(Wv2 + k), v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    v += bv2[0];
    const float mean = sum / (float)A;
    if (lane == 0)
        for (int a = 0; a < A; ++a) q_out[(size_t)warp * A + a] = v + adv[a] - mean;
}

// backward of the dueling output layer: dq -> (dadv | dval) and masked d(hidden layer)
__global__ void head_out_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ hid,
                                    const float* __restrict__ Wa2, const float* __restrict__ Wv2,
                                    const int* __restrict__ d_rows, float grad_scale_unused, int rows_cap, int A,
                                    float* __restrict__ dout16, float* __restrict__ dhid) {
    const int r = blockIdx.x;
    __shared__ float s_d[17];
    const bool live = r < *d_rows;
    if (threadIdx.x < 16) {
        float v = 0.f;
        if (live) {
            float tot = 0.f;
            for (int a = 0; a < A; ++a) tot += dq[(size_t)r * A + a];
            const int a = threadIdx.x;
            if (a < A) v = dq[(size_t)r * A + a] - tot / (float)A;     // d adv (through the mean)
            else if (a == A) v = tot;                                   // d val
        }
        s_d[threadIdx.x] = v;
        dout16[(size_t)r * 16 + threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * H; j += blockDim.x) {
        float g = 0.f;
        if (live && hid[(size_t)r * 2 * H + j] > 0.f) {
            if (j < H) { for (int a = 0; a < A; ++a) g = fmaf(s_d[a], __ldg(Wa2 + a * H + j), g); }
            else g = s_d[A] * __ldg(Wv2 + j - H);
        }
        dhid[(size_t)r * 2 * H + j] = g;
    }
}

// LSTM cell epilogue fused into the recurrent GEMM (gate-interleaved columns: n = 4*j + gate)
struct EpiLstmCell {
    const float* xp_t; const float* c_prev; const float* h_prev; int ld_prev;
    float* h_out; float* c_out; float* gates_out; const int* len; int t, B;
    __device__ __forceinline__ void store4(int m, int n, const float* a, int) const {
        if (m >= B || n >= G4) return;
        const float4 x = *reinterpret_cast<const float4*>(xp_t + (size_t)m * G4 + n);
        const int j = n >> 2;
        const float gi = 1.f / (1.f + expf(-(a[0] + x.x)));
        const float gf = 1.f / (1.f + expf(-(a[1] + x.y)));
        const float gg = tanhf(a[2] + x.z);
        const float go = 1.f / (1.f + expf(-(a[3] + x.w)));
        const float cp = c_prev[(size_t)m * ld_prev + j];
        const float cn = gf * cp + gi * gg;
        const float hn = go * tanhf(cn);
        const bool live = t < len[m];
        if (gates_out) *reinterpret_cast<float4*>(gates_out + (size_t)m * G4 + n) = make_float4(gi, gf, gg, go);
        c_out[(size_t)m * H + j] = live ? cn : cp;
        h_out[(size_t)m * H + j] = live ? hn : h_prev[(size_t)m * ld_prev + j];
    }
};

// BPTT pointwise step: d(pre-activation gates) at time t
__global__ void lstm_bwd_pointwise_kernel(const float* __restrict__ dH_t, const float* __restrict__ dhrec, int nparts,
                                          float* __restrict__ dcrec, const float* __restrict__ G_t,
                                          const float* __restrict__ C_t, const float* __restrict__ C_prev, int ld_cprev,
                                          const int* __restrict__ len_learn, int t, int B, float* __restrict__ DG_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, j = i % H;
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < len_learn[b]) {
        const float4 g = *reinterpret_cast<const float4*>(G_t + (size_t)b * G4 + 4 * j);
        const float gi = g.x, gf = g.y, gg = g.z, go = g.w;
        const float tc = tanhf(C_t[i]);
        float dh = dH_t[i];
        for (int p = 0; p < nparts; ++p) dh += dhrec[(size_t)p * B * H + i];      // split-K partials of dgates.W_hh
        const float dc = dcrec[i] + dh * go * (1.f - tc * tc);
        const float cp = C_prev[(size_t)b * ld_cprev + j];
        out.x = dc * gg * gi * (1.f - gi);
        out.y = dc * cp * gf * (1.f - gf);
        out.z = dc * gi * (1.f - gg * gg);
        out.w = dh * tc * go * (1.f - go);
        dcrec[i] = dc * gf;
    }
    *reinterpret_cast<float4*>(DG_t + (size_t)b * G4 + 4 * j) = out;
}

// previous-hidden operand of the W_hh wgrad: X(j, row=(t,b)) = t>0 ? Hs[t-1][b][j] : h0[b][j]
struct HprevM {
    static constexpr bool kKMajor = false;
    const float* Hs; const float* hidden; int B, rows;   // hidden: [B][2][H]
    __device__ __forceinline__ void load4(int row, int k, float (&v)[4]) const {
        if (k >= rows || row >= H) { zero4(v); return; }
        const int t = k / B, b = k - t * B;
        const float* src = (t > 0) ? Hs + ((size_t)(t - 1) * B + b) * H : hidden + (size_t)b * 2 * H;
        ld4(src + row, v);
    }
};

// FC epilogue: latent -> U (time-major rows), ReLU
struct EpiLatent {
    float* U; const float* bias; int B, T, KU;
    __device__ __forceinline__ void store4(int m, int n, const float* a, int) const {
        if (m >= B * T || n >= LATENT) return;
        const int b = m / T, t = m - b * T;
        float4 r = make_float4(fmaxf(a[0] + bias[n], 0.f), fmaxf(a[1] + bias[n + 1], 0.f), fmaxf(a[2] + bias[n + 2], 0.f),
                               fmaxf(a[3] + bias[n + 3], 0.f));
        *reinterpret_cast<float4*>(U + ((size_t)t * B + b) * KU + n) = r;
    }
};
// d latent epilogue: rows are time-major (t,b); mask by latent>0; write frame-major for the encoder backward
struct EpiDLatent {
    float* dlat; const float* U; int B, T, KU;
    __device__ __forceinline__ void store4(int m, int n, const float* a, int) const {
        if (m >= B * T || n >= LATENT) return;
        const int t = m / B, b = m - t * B;
        const float4 y = *reinterpret_cast<const float4*>(U + (size_t)m * KU + n);
        *reinterpret_cast<float4*>(dlat + ((size_t)b * T + t) * LATENT + n) =
            make_float4(y.x > 0.f ? a[0] : 0.f, y.y > 0.f ? a[1] : 0.f, y.z > 0.f ? a[2] : 0.f, y.w > 0.f ? a[3] : 0.f);
    }
};
// conv2 dgrad epilogue for parity class (py,px): rows (f, y', x') on the 10x10 grid -> act1 position (2y'+py, 2x'+px)
struct EpiDgradS2 {
    float* out; const float* act; int nframes, py, px;
    __device__ __forceinline__ void store4(int m, int n, const float* a, int) const {
        if (m >= nframes * 100 || n >= 32) return;
        const int f = m / 100, p = m - f * 100, yq = p / 10, xq = p - yq * 10;
        const size_t o = (((size_t)f * 20 + 2 * yq + py) * 20 + 2 * xq + px) * 32 + n;
        const float4 y = *reinterpret_cast<const float4*>(act + o);
        *reinterpret_cast<float4*>(out + o) =
            make_float4(y.x > 0.f ? a[0] : 0.f, y.y > 0.f ? a[1] : 0.f, y.z > 0.f ? a[2] : 0.f, y.w > 0.f ? a[3] : 0.f);
    }
};
// scatter rows of d(hidden rows) to dH[t][b] through the row map
struct EpiScatterRows {
    float* dH; const int* src; const int* d_rows; int Rmax;
    __device__ __forceinline__ void store4(int m, int n, const float* a, int) const {
        if (m >= Rmax || n >= H || m >= *d_rows) return;
        const int s = src[m];
        if (s < 0) return;
        *reinterpret_cast<float4*>(dH + (size_t)s * H + n) = make_float4(a[0], a[1], a[2], a[3]);
    }
};

// split-K reduce + routing of a weight gradient into the reference's parameter layout
enum RouteKind { R_C1, R_C2, R_C3, R_FC, R_WIH, R_WHH, R_H0, R_H2 };
__global__ void reduce_route_kernel(const float* __restrict__ ws, int splits, int M, int N, int kind, float* __restrict__ g,
                                    const int64_t* __restrict__ off, int A, int C, float scale) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * N) return;
    const int m = i / N, n = i % N;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(size_t)z * M * N + i];
    s *= scale;
    const int KIH = LATENT + A + 1;
    switch (kind) {
        case R_C1: g[off[P_C1W] + (int64_t)m * 64 * C + n] = s; break;
        case R_C2: { const int tap = n >> 5, c = n & 31; g[off[P_C2W] + m * 512 + c * 16 + tap] = s; } break;
        case R_C3: { const int tap = n >> 6, c = n & 63; g[off[P_C3W] + m * 576 + c * 9 + tap] = s; } break;
        case R_FC: { const int hw = n >> 6, c = n & 63; g[off[P_FCW] + (int64_t)m * FLAT3 + c * 49 + hw] = s; } break;
        case R_WIH: { const int row = (m & 3) * H + (m >> 2); if (n < KIH) g[off[P_WIH] + (int64_t)row * KIH + n] = s; } break;
        case R_WHH: { const int row = (m & 3) * H + (m >> 2); g[off[P_WHH] + (int64_t)row * H + n] = s; } break;
        case R_H0: { if (m < H) g[off[P_A0W] + (int64_t)m * H + n] = s; else g[off[P_V0W] + (int64_t)(m - H) * H + n] = s; } break;
        case R_H2: { if (m < A && n < H) g[off[P_A2W] + m * H + n] = s; else if (m == A && n >= H) g[off[P_V2W] + n - H] = s; } break;
    }
}

// deterministic column sums: partial[p][n] = sum over row chunk p ; then final routing
__global__ void colsum_partial_kernel(const float* __restrict__ X, int M, int N, int chunk, float* __restrict__ part) {
    __shared__ float s[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x, p = blockIdx.y;
    const int r0 = p * chunk, r1 = min(M, r0 + chunk);
    float acc = 0.f;
    if (col < N)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) acc += X[(size_t)r * N + col];
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && col < N) {
        float t = 0.f;
        for (int y = 0; y < 8; ++y) t += s[y][threadIdx.x];
        part[(size_t)p * N + col] = t;
    }
}
enum BiasKind { B_PLAIN, B_LSTM, B_H0, B_H2 };
__global__ void colsum_final_kernel(const float* __restrict__ part, int P, int N, int kind, float* __restrict__ g,
                                    int64_t o0, int64_t o1, int A) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += part[(size_t)p * N + n];
    switch (kind) {
        case B_PLAIN: g[o0 + n] = s; break;
        case B_LSTM: { const int row = (n & 3) * H + (n >> 2); g[o0 + row] = s; g[o1 + row] = s; } break;   // b_ih and b_hh
        case B_H0: { if (n < H) g[o0 + n] = s; else g[o1 + n - H] = s; } break;
        case B_H2: { if (n < A) g[o0 + n] = s; else if (n == A) g[o1] = s; } break;
    }
}

static cudaError_t colsum(const float* X, int M, int N, int kind, float* g, int64_t o0, int64_t o1, int A, float* colws,
                          cudaStream_t s) {
    const int P = 128;
    const int chunk = (M + P - 1) / P;
    dim3 grid((N + 31) / 32, P), block(32, 8);
    colsum_partial_kernel<<<grid, block, 0, s>>>(X, M, N, chunk, colws);
    colsum_final_kernel<<<(N + 127) / 128, 128, 0, s>>>(colws, P, N, kind, g, o0, o1, A);
    return cudaGetLastError();
}

template <int BM, int BN, int UBN, class AL, class BL>
static cudaError_t wgrad(const AL& al, const BL& bl, int M, int N, int K, int splits, int kind, r2d2_net* net, float* grads,
                         const int64_t* d_off, float scale, cudaStream_t s) {
    if ((size_t)splits * M * N > net->ws_floats) return cudaErrorInvalidValue;
    EpiPartial ep{net->ws, M, N};
    cudaError_t e = run_gemm<BM, BN, UBN>(al, bl, ep, M, N, K, splits, s);
    if (e != cudaSuccess) return e;
    const int64_t tot = (int64_t)M * N;
    reduce_route_kernel<<<cdiv(tot, 256), 256, 0, s>>>(net->ws, splits, M, N, kind, grads, d_off, net->A, net->C, scale);
    return cudaGetLastError();
}

}  // namespace r2d2

using namespace r2d2;


extern "C" {

int r2d2_net_param_layout(int A, int C, int64_t* offsets_out /* [21] */) {
    R2D2_REQUIRE(offsets_out && A >= 1 && A <= 32 && C >= 1 && C <= 16, "bad arguments");
    int64_t n[NPARAM];
    param_sizes(A, C, n);
    int64_t o = 0;
    for (int i = 0; i < NPARAM; ++i) { offsets_out[i] = o; o += (n[i] + 3) / 4 * 4; }   // 16-byte aligned tensors
    offsets_out[NPARAM] = o;
    return R2D2_OK;
}

}  // extern "C"

// device copy of the parameter offsets, kept in a side table keyed by handle
static std::map<r2d2_net*, int64_t*> g_doff;

static int alloc_f(float** p, size_t n) {
    R2D2_CUDA_CHECK(cudaMalloc(p, n * sizeof(float)));
    R2D2_CUDA_CHECK(cudaMemset(*p, 0, n * sizeof(float)));
    return R2D2_OK;
}

extern "C" {

int r2d2_net_create(int B, int T, int C, int A, int Lmax, int max_forward, r2d2_net** out) {
    R2D2_REQUIRE(out && B >= 1 && B <= 4096 && T >= 1 && T <= 255 && C >= 1 && C <= 16 && A >= 1 && A <= 15 && Lmax >= 1 &&
                     Lmax <= T && max_forward >= 0,
                 "bad shape");
    r2d2_net* n = new r2d2_net();
    memset(n, 0, sizeof(*n));
    n->B = B; n->T = T; n->C = C; n->A = A; n->Lmax = Lmax; n->F = max_forward;
    n->KIH = LATENT + A + 1;
    n->KU = (n->KIH + 15) / 16 * 16;
    n->NF = B * T;
    n->Rmax = (B * Lmax + 3) / 4 * 4;
    r2d2_net_param_layout(A, C, n->off);
    int64_t* d_off = nullptr;
    R2D2_CUDA_CHECK(cudaMalloc(&d_off, sizeof(n->off)));
    R2D2_CUDA_CHECK(cudaMemcpy(d_off, n->off, sizeof(n->off), cudaMemcpyHostToDevice));
    g_doff[n] = d_off;
    const size_t NF = n->NF, TB = (size_t)T * B;
    int rc = 0;
    for (int k = 0; k < 2 && !rc; ++k) {
        Packed& p = n->pk[k];
        rc |= alloc_f(&p.W2p, 64 * 512); rc |= alloc_f(&p.W3p, 64 * 576); rc |= alloc_f(&p.Wfcp, 512ull * FLAT3);
        rc |= alloc_f(&p.Wih_p, (size_t)G4 * n->KU); rc |= alloc_f(&p.Whh_p, (size_t)G4 * H); rc |= alloc_f(&p.bias_p, G4);
        rc |= alloc_f(&p.Wh0, 2 * H * H); rc |= alloc_f(&p.bh0, 2 * H); rc |= alloc_f(&p.W3d, 64 * 576);
        rc |= alloc_f(&p.W2d, 4 * 32 * 256);
        Acts& a = n->ac[k];
        rc |= alloc_f(&a.act1, NF * 12800); rc |= alloc_f(&a.act2, NF * 5184); rc |= alloc_f(&a.act3, NF * FLAT3);
        rc |= alloc_f(&a.U, TB * n->KU); rc |= alloc_f(&a.XP, TB * G4); rc |= alloc_f(&a.Hs, TB * H); rc |= alloc_f(&a.Cs, TB * H);
        rc |= alloc_f(&a.Gs, k == 0 ? TB * G4 : 4); rc |= alloc_f(&a.hid, (size_t)2 * n->Rmax * 2 * H);
    }
    if (rc) return rc;
    R2D2_CUDA_CHECK(cudaMalloc(&n->row_src, 2 * n->Rmax * sizeof(int)));
    R2D2_CUDA_CHECK(cudaMalloc(&n->len_full, B * sizeof(int)));
    R2D2_CUDA_CHECK(cudaMalloc(&n->len_learn, B * sizeof(int)));
    R2D2_CUDA_CHECK(cudaMalloc(&n->d_rows, sizeof(int)));
    rc |= alloc_f(&n->dH, TB * H); rc |= alloc_f(&n->DG, TB * G4); rc |= alloc_f(&n->dhrec, (size_t)kRecSplits * B * H);
    rc |= alloc_f(&n->dcrec, (size_t)B * H); rc |= alloc_f(&n->dlat, NF * LATENT); rc |= alloc_f(&n->dpre3, NF * FLAT3);
    rc |= alloc_f(&n->dpre2, NF * 5184); rc |= alloc_f(&n->dpre1, NF * 12800); rc |= alloc_f(&n->dhid, (size_t)n->Rmax * 2 * H);
    rc |= alloc_f(&n->dout16, (size_t)n->Rmax * 16);
    n->ws_floats = 32ull << 20;                        // 128 MB split-K workspace
    rc |= alloc_f(&n->ws, n->ws_floats); rc |= alloc_f(&n->colws, 128 * 4096);
    if (rc) return rc;
    R2D2_CUDA_CHECK(cudaDeviceSynchronize());
    *out = n;
    return R2D2_OK;
}

int r2d2_net_destroy(r2d2_net* n) {
    if (!n) return R2D2_OK;
    for (int k = 0; k < 2; ++k) {
        Packed& p = n->pk[k];
        float* ps[] = {p.W2p, p.W3p, p.Wfcp, p.Wih_p, p.Whh_p, p.bias_p, p.Wh0, p.bh0, p.W3d, p.W2d};
        for (float* x : ps) cudaFree(x);
        Acts& a = n->ac[k];
        float* as[] = {a.act1, a.act2, a.act3, a.U, a.XP, a.Hs, a.Cs, a.Gs, a.hid};
        for (float* x : as) cudaFree(x);
    }
    float* fs[] = {n->dH, n->DG, n->dhrec, n->dcrec, n->dlat, n->dpre3, n->dpre2, n->dpre1, n->dhid, n->dout16, n->ws, n->colws};
    for (float* x : fs) cudaFree(x);
    cudaFree(n->row_src); cudaFree(n->len_full); cudaFree(n->len_learn); cudaFree(n->d_rows);
    cudaFree(g_doff[n]);
    g_doff.erase(n);
    delete n;
    return R2D2_OK;
}

int r2d2_net_rows_capacity(const r2d2_net* n) { return n ? n->Rmax : -1; }

/* re-lay out the caller's flat parameter buffer (reference state_dict layout) for slot `which` */
int r2d2_net_pack(r2d2_net* n, int which, const float* params, void* stream) {
    R2D2_REQUIRE(n && (which == 0 || which == 1) && params, "bad arguments");
    const int64_t work = 512ll * FLAT3 > (int64_t)G4 * n->KU ? 512ll * FLAT3 : (int64_t)G4 * n->KU;
    pack_kernel<<<cdiv(work, 256), 256, 0, as_stream(stream)>>>(params, g_doff[n], n->pk[which], n->A, n->KU);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

struct FwdArgs {
    const float* params; const uint8_t* obs; const uint8_t* last_action; const float* last_reward; const float* hidden;
};

// encoder + input projection of one slot (model.py:39-49,92); 1/255 of worker.py:342 folded into the conv1 epilogue
static int net_encode(r2d2_net* n, int which, const FwdArgs& fa, cudaStream_t s) {
    const int B = n->B, T = n->T, C = n->C, A = n->A, KU = n->KU, NF = n->NF;
    Packed& pk = n->pk[which];
    Acts& ac = n->ac[which];
    const int64_t* off = n->off;
    const float* params = fa.params;
    side_columns_kernel<<<T * B, 32, 0, s>>>(ac.U, fa.last_action, fa.last_reward, B, T, A, KU);
    R2D2_LAUNCH_CHECK();
    {
        Conv1FrameK a{fa.obs, C, NF};
        MatK b{params + off[P_C1W], 32, C * 64, C * 64};
        EpiBias<true> e{ac.act1, params + off[P_C1B], NF * 400, 32, 32, 1.f / 255.f};
        R2D2_CUDA_CHECK((run_gemm<128, 32, 32>(a, b, e, NF * 400, 32, C * 64, 1, s)));
    }
    {
        ConvNHWC_K<20, 20, 32, 9, 9, 4, 4, 2> a{ac.act1, NF};
        MatK b{pk.W2p, 64, 512, 512};
        EpiBias<true> e{ac.act2, params + off[P_C2B], NF * 81, 64, 64, 1.f};
        R2D2_CUDA_CHECK((run_gemm<128, 64, 64>(a, b, e, NF * 81, 64, 512, 1, s)));
    }
    {
        ConvNHWC_K<9, 9, 64, 7, 7, 3, 3, 1> a{ac.act2, NF};
        MatK b{pk.W3p, 64, 576, 576};
        EpiBias<true> e{ac.act3, params + off[P_C3B], NF * 49, 64, 64, 1.f};
        R2D2_CUDA_CHECK((run_gemm<128, 64, 64>(a, b, e, NF * 49, 64, 576, 1, s)));
    }
    {
        MatK a{ac.act3, NF, FLAT3, FLAT3};
        MatK b{pk.Wfcp, LATENT, FLAT3, FLAT3};
        EpiLatent e{ac.U, params + off[P_FCB], B, T, KU};
        R2D2_CUDA_CHECK((run_gemm<128, 128, 128>(a, b, e, NF, LATENT, FLAT3, 1, s)));
    }
    {   // LSTM input projection for all steps at once (hoisted out of the recurrence)
        MatK a{ac.U, T * B, KU, KU};
        MatK b{pk.Wih_p, G4, KU, KU};
        EpiBias<false> e{ac.XP, pk.bias_p, T * B, G4, G4, 1.f};
        R2D2_CUDA_CHECK((run_gemm<128, 128, 128>(a, b, e, T * B, G4, KU, 1, s)));
    }
    return R2D2_OK;
}

struct StepOps { MatK a; MatK b; EpiLstmCell e; };
static StepOps lstm_step_ops(r2d2_net* n, int which, const float* hidden, int t) {
    const int B = n->B;
    Packed& pk = n->pk[which];
    Acts& ac = n->ac[which];
    const float* hp = t ? ac.Hs + (size_t)(t - 1) * B * H : hidden;
    const float* cp = t ? ac.Cs + (size_t)(t - 1) * B * H : hidden + H;
    const int ldp = t ? H : 2 * H;
    return StepOps{MatK{hp, B, H, ldp}, MatK{pk.Whh_p, G4, H, H},
                   EpiLstmCell{ac.XP + (size_t)t * B * G4, cp, hp, ldp, ac.Hs + (size_t)t * B * H, ac.Cs + (size_t)t * B * H,
                               which == 0 ? ac.Gs + (size_t)t * B * G4 : nullptr, n->len_full, t, B}};
}

// recurrence (model.py:95-100 / 134-141): sequences advance while t < b+l+f.  which = 0/1: one slot; 2: both slots
// in the same launches (two independent recurrences hide each other's per-step latency).
static int net_recurrence(r2d2_net* n, int which, const float* hidden, cudaStream_t s) {
    const int B = n->B, T = n->T;
    for (int t = 0; t < T; ++t) {
        if (which == 2) {
            StepOps o0 = lstm_step_ops(n, 0, hidden, t), o1 = lstm_step_ops(n, 1, hidden, t);
            R2D2_CUDA_CHECK((run_gemm_pair<64, 64, 64>(o0.a, o0.b, o0.e, o1.a, o1.b, o1.e, B, G4, H, s)));
        } else {
            StepOps o = lstm_step_ops(n, which, hidden, t);
            R2D2_CUDA_CHECK((run_gemm<64, 64, 64>(o.a, o.b, o.e, B, G4, H, 1, s)));
        }
    }
    return R2D2_OK;
}

// dueling head on the gathered rows (model.py:102-117, 143-148)
static int net_heads(r2d2_net* n, int which, const float* params, float* q_learn_out, float* q_shift_out, cudaStream_t s) {
    const int A = n->A, Rmax = n->Rmax;
    Packed& pk = n->pk[which];
    Acts& ac = n->ac[which];
    const int64_t* off = n->off;
    const int nsets = 2;
    {
        RowGatherK a{ac.Hs, n->row_src, nsets * Rmax, H, H};
        MatK b{pk.Wh0, 2 * H, H, H};
        EpiBias<true> e{ac.hid, pk.bh0, nsets * Rmax, 2 * H, 2 * H, 1.f};
        R2D2_CUDA_CHECK((run_gemm<128, 128, 128>(a, b, e, nsets * Rmax, 2 * H, H, 1, s)));
    }
    if (q_learn_out)
        head_out_kernel<<<cdiv((int64_t)Rmax * 32, 256), 256, 0, s>>>(ac.hid, params + off[P_A2W], params + off[P_A2B],
                                                                     params + off[P_V2W], params + off[P_V2B], Rmax, A, q_learn_out);
    if (q_shift_out)
        head_out_kernel<<<cdiv((int64_t)Rmax * 32, 256), 256, 0, s>>>(ac.hid + (size_t)Rmax * 2 * H, params + off[P_A2W],
                                                                     params + off[P_A2B], params + off[P_V2W],
                                                                     params + off[P_V2B], Rmax, A, q_shift_out);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

static int net_prep(r2d2_net* n, const uint8_t* burn, const uint8_t* learn, const uint8_t* fwd, cudaStream_t s) {
    prep_rows_kernel<<<1, 256, (n->B + 1) * sizeof(int), s>>>(burn, learn, fwd, n->B, n->F, n->Rmax, n->row_src, n->len_full,
                                                             n->len_learn, n->d_rows);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

/* Forward unroll of slot `which` (0 online, 1 target).  model.py:81-150.
 *   obs u8 [B][T][C][84][84]; last_action u8/bool [B][T][A]; last_reward f32 [B][T];
 *   hidden f32 [B][2][H] ([b][0]=h0, [b][1]=c0 -- the Block.hidden layout, worker.py:198);
 *   burn/learn/fwd u8 [B].
 *   q_learn_out  [Rmax][A] : Q at the learning positions (calculate_q rows), may be NULL
 *   q_shift_out  [Rmax][A] : Q at the n-step shifted positions (calculate_q_ rows), may be NULL
 * Rows are sequence-major; only the first sum(learn) rows are meaningful. */
int r2d2_net_forward(r2d2_net* n, int which, const float* params, const uint8_t* obs, const uint8_t* last_action,
                     const float* last_reward, const float* hidden, const uint8_t* burn, const uint8_t* learn,
                     const uint8_t* fwd, float* q_learn_out, float* q_shift_out, void* stream) {
    R2D2_REQUIRE(n && (which == 0 || which == 1) && params && obs && last_action && last_reward && hidden && burn && learn && fwd,
                 "bad arguments");
    cudaStream_t s = as_stream(stream);
    if (which == 0) { n->obs = obs; n->hidden = hidden; }
    FwdArgs fa{params, obs, last_action, last_reward, hidden};
    int rc = net_prep(n, burn, learn, fwd, s);
    if (!rc) rc = net_encode(n, which, fa, s);
    if (!rc) rc = net_recurrence(n, which, hidden, s);
    if (!rc) rc = net_heads(n, which, params, q_learn_out, q_shift_out, s);
    return rc;
}

/* The learner's three Q tensors in one call (worker.py:346,347,352): online and target unrolls on the same
 * batch with the two recurrences advanced together.  q_learn_out / qn_online_out from the online parameters,
 * qn_target_out from the target parameters. */
int r2d2_net_forward_pair(r2d2_net* n, const float* params_online, const float* params_target, const uint8_t* obs,
                          const uint8_t* last_action, const float* last_reward, const float* hidden, const uint8_t* burn,
                          const uint8_t* learn, const uint8_t* fwd, float* q_learn_out, float* qn_online_out,
                          float* qn_target_out, void* stream) {
    R2D2_REQUIRE(n && params_online && params_target && obs && last_action && last_reward && hidden && burn && learn && fwd &&
                     q_learn_out && qn_online_out && qn_target_out,
                 "bad arguments");
    cudaStream_t s = as_stream(stream);
    n->obs = obs; n->hidden = hidden;
    FwdArgs f0{params_online, obs, last_action, last_reward, hidden}, f1{params_target, obs, last_action, last_reward, hidden};
    int rc = net_prep(n, burn, learn, fwd, s);
    if (!rc) rc = net_encode(n, 0, f0, s);
    if (!rc) rc = net_encode(n, 1, f1, s);
    if (!rc) rc = net_recurrence(n, 2, hidden, s);
    if (!rc) rc = net_heads(n, 0, params_online, q_learn_out, qn_online_out, s);
    if (!rc) rc = net_heads(n, 1, params_target, nullptr, qn_target_out, s);
    return rc;
}

/* BPTT + encoder backward of the ONLINE slot (loss.backward(), worker.py:363).
 *   dq [Rmax][A] = d loss / d q_learn (rows >= sum(learn) ignored); grads: flat buffer in the
 *   parameter layout, fully overwritten. */
int r2d2_net_backward(r2d2_net* n, const float* params, const float* dq, float* grads, void* stream) {
    R2D2_REQUIRE(n && params && dq && grads && n->obs, "bad arguments / forward(online) not run");
    cudaStream_t s = as_stream(stream);
    const int B = n->B, T = n->T, C = n->C, A = n->A, KU = n->KU, NF = n->NF, Rmax = n->Rmax;
    Packed& pk = n->pk[0];
    Acts& ac = n->ac[0];
    const int64_t* off = n->off;
    const int64_t* d_off = g_doff[n];
    const size_t TB = (size_t)T * B;

    // ---- head
    head_out_bwd_kernel<<<Rmax, 256, 0, s>>>(dq, ac.hid, params + off[P_A2W], params + off[P_V2W], n->d_rows, 1.f, Rmax, A,
                                            n->dout16, n->dhid);
    R2D2_LAUNCH_CHECK();
    {   // layer-2 weights: [A+1 (pad 16)] x [1024] = dout16^T . hid
        MatM a{n->dout16, 16, Rmax, 16};
        MatM b{ac.hid, 2 * H, Rmax, 2 * H};
        R2D2_CUDA_CHECK((wgrad<64, 64, 128>(a, b, 16, 2 * H, Rmax, 4, R_H2, n, grads, d_off, 1.f, s)));
        R2D2_CUDA_CHECK(colsum(n->dout16, Rmax, 16, B_H2, grads, off[P_A2B], off[P_V2B], A, n->colws, s));
    }
    {   // layer-0 weights: [1024] x [512] = dhid^T . Hsel
        MatM a{n->dhid, 2 * H, Rmax, 2 * H};
        RowGatherM b{ac.Hs, n->row_src, H, Rmax, H};
        R2D2_CUDA_CHECK((wgrad<128, 128, 128>(a, b, 2 * H, H, Rmax, 4, R_H0, n, grads, d_off, 1.f, s)));
        R2D2_CUDA_CHECK(colsum(n->dhid, Rmax, 2 * H, B_H0, grads, off[P_A0B], off[P_V0B], A, n->colws, s));
    }
    R2D2_CUDA_CHECK(cudaMemsetAsync(n->dH, 0, TB * H * sizeof(float), s));
    {   // d hidden rows -> dH[t][b]
        MatK a{n->dhid, Rmax, 2 * H, 2 * H};
        MatM b{pk.Wh0, H, 2 * H, H};
        EpiScatterRows e{n->dH, n->row_src, n->d_rows, Rmax};
        R2D2_CUDA_CHECK((run_gemm<128, 128, 128>(a, b, e, Rmax, H, 2 * H, 1, s)));
    }
    // ---- BPTT through all b+l steps, burn-in included (no detach anywhere in model.py:122-150)
    R2D2_CUDA_CHECK(cudaMemsetAsync(n->dcrec, 0, (size_t)B * H * sizeof(float), s));
    for (int t = T - 1; t >= 0; --t) {
        const float* cprev = t ? ac.Cs + (size_t)(t - 1) * B * H : n->hidden + H;
        lstm_bwd_pointwise_kernel<<<cdiv(B * H, 256), 256, 0, s>>>(n->dH + (size_t)t * B * H, n->dhrec,
                                                                  t == T - 1 ? 0 : kRecSplits, n->dcrec,
                                                                  ac.Gs + (size_t)t * B * G4, ac.Cs + (size_t)t * B * H, cprev,
                                                                  t ? H : 2 * H, n->len_learn, t, B, n->DG + (size_t)t * B * G4);
        if (t > 0) {
            MatK a{n->DG + (size_t)t * B * G4, B, G4, G4};
            MatM b{pk.Whh_p, H, G4, H};
            EpiPartial e{n->dhrec, B, H};
            R2D2_CUDA_CHECK((run_gemm<64, 64, 64>(a, b, e, B, H, G4, kRecSplits, s)));
        }
    }
    R2D2_LAUNCH_CHECK();
    {   // recurrent weight gradients over all (t,b) rows
        MatM a{n->DG, G4, (int)TB, G4};
        HprevM bh{ac.Hs, n->hidden, B, (int)TB};
        R2D2_CUDA_CHECK((wgrad<128, 128, 128>(a, bh, G4, H, (int)TB, 4, R_WHH, n, grads, d_off, 1.f, s)));
        MatM bu{ac.U, KU, (int)TB, KU};
        R2D2_CUDA_CHECK((wgrad<128, 128, 128>(a, bu, G4, KU, (int)TB, 4, R_WIH, n, grads, d_off, 1.f, s)));
        R2D2_CUDA_CHECK(colsum(n->DG, (int)TB, G4, B_LSTM, grads, off[P_BIH], off[P_BHH], A, n->colws, s));
    }
    {   // d latent (ReLU-masked), frame-major
        MatK a{n->DG, (int)TB, G4, G4};
        MatM b{pk.Wih_p, LATENT, G4, KU};
        EpiDLatent e{n->dlat, ac.U, B, T, KU};
        R2D2_CUDA_CHECK((run_gemm<128, 128, 128>(a, b, e, (int)TB, LATENT, G4, 1, s)));
    }
    // ---- encoder backward
    {
        MatM a{n->dlat, LATENT, NF, LATENT};
        MatM b{ac.act3, FLAT3, NF, FLAT3};
        R2D2_CUDA_CHECK((wgrad<128, 128, 128>(a, b, LATENT, FLAT3, NF, 2, R_FC, n, grads, d_off, 1.f, s)));
        R2D2_CUDA_CHECK(colsum(n->dlat, NF, LATENT, B_PLAIN, grads, off[P_FCB], 0, A, n->colws, s));
        MatK a2{n->dlat, NF, LATENT, LATENT};
        MatM b2{pk.Wfcp, FLAT3, LATENT, FLAT3};
        EpiMasked e{n->dpre3, ac.act3, NF, FLAT3, FLAT3};
        R2D2_CUDA_CHECK((run_gemm<128, 128, 128>(a2, b2, e, NF, FLAT3, LATENT, 1, s)));
    }
    {   // conv3
        MatM a{n->dpre3, 64, NF * 49, 64};
        ConvNHWC_M<9, 9, 64, 7, 7, 3, 3, 1> b{ac.act2, NF};
        R2D2_CUDA_CHECK((wgrad<64, 64, 64>(a, b, 64, 576, NF * 49, 32, R_C3, n, grads, d_off, 1.f, s)));
        R2D2_CUDA_CHECK(colsum(n->dpre3, NF * 49, 64, B_PLAIN, grads, off[P_C3B], 0, A, n->colws, s));
        ConvDgradK<9, 9, 7, 7, 64, 3, 3> a2{n->dpre3, NF};
        MatK b2{pk.W3d, 64, 576, 576};
        EpiMasked e{n->dpre2, ac.act2, NF * 81, 64, 64};
        R2D2_CUDA_CHECK((run_gemm<128, 64, 64>(a2, b2, e, NF * 81, 64, 576, 1, s)));
    }
    {   // conv2
        MatM a{n->dpre2, 64, NF * 81, 64};
        ConvNHWC_M<20, 20, 32, 9, 9, 4, 4, 2> b{ac.act1, NF};
        R2D2_CUDA_CHECK((wgrad<64, 64, 64>(a, b, 64, 512, NF * 81, 36, R_C2, n, grads, d_off, 1.f, s)));
        R2D2_CUDA_CHECK(colsum(n->dpre2, NF * 81, 64, B_PLAIN, grads, off[P_C2B], 0, A, n->colws, s));
        for (int cls = 0; cls < 4; ++cls) {      // stride-2 dgrad as four stride-1 problems (output parity classes)
            ConvDgradK<10, 10, 9, 9, 64, 2, 2> a2{n->dpre2, NF};
            MatK b2{pk.W2d + cls * 32 * 256, 32, 256, 256};
            EpiDgradS2 e{n->dpre1, ac.act1, NF, cls >> 1, cls & 1};
            R2D2_CUDA_CHECK((run_gemm<128, 32, 32>(a2, b2, e, NF * 100, 32, 256, 1, s)));
        }
    }
    {   // conv1 (weights only; frames need no gradient)
        MatM a{n->dpre1, 32, NF * 400, 32};
        Conv1FrameM b{n->obs, C, NF};
        R2D2_CUDA_CHECK((wgrad<32, 64, 64>(a, b, 32, C * 64, NF * 400, 148, R_C1, n, grads, d_off, 1.f / 255.f, s)));
        R2D2_CUDA_CHECK(colsum(n->dpre1, NF * 400, 32, B_PLAIN, grads, off[P_C1B], 0, A, n->colws, s));
    }
    return R2D2_OK;
}

/* test/debug access to intermediates: name in {"U","XP","Hs","Cs","Gs","act1","act2","act3","hid","dH","DG","dlat",
 * "dpre1","dpre2","dpre3","row_src","rows"} */
void* r2d2_net_debug_ptr(r2d2_net* n, int which, const char* name) {
    if (!n || !name) return nullptr;
    Acts& a = n->ac[which & 1];
    struct { const char* k; void* v; } tab[] = {
        {"U", a.U}, {"XP", a.XP}, {"Hs", a.Hs}, {"Cs", a.Cs}, {"Gs", a.Gs}, {"act1", a.act1}, {"act2", a.act2}, {"act3", a.act3},
        {"hid", a.hid}, {"dH", n->dH}, {"DG", n->DG}, {"dlat", n->dlat}, {"dpre1", n->dpre1}, {"dpre2", n->dpre2},
        {"dpre3", n->dpre3}, {"row_src", n->row_src}, {"rows", n->d_rows}};
    for (auto& e : tab)
        if (!strcmp(e.k, name)) return e.v;
    return nullptr;
}
int r2d2_net_ku(const r2d2_net* n) { return n ? n->KU : -1; }

}  // extern "C"
