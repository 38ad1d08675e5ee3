// K1 / K1b: sequence-unroll forward and BPTT backward of the R2D2 network
// (model.py:27-150 of the reference: conv encoder -> LSTM -> dueling head) on tcgen05 tensor cores.
//
// One `r2d2_net` handle owns the packed weights, saved activations and backward scratch of the
// online (slot 0) and target (slot 1) networks for a fixed (B, T, C, A) batch shape.
//
// Data layout in HBM.  "split" = two bf16 tensors (hi, lo) of identical layout with x = hi + lo
// (|err| <= 2^-17 |x|): every tensor that feeds a contraction is stored that way by its producer so
// the GEMM (umma2.cuh) and window-convolution (winconv.cuh) kernels stage operands with pure 16-byte cp.async copies.
//   s2d   [NF][21][21][16C] bf16   frames after space-to-depth by 4 (u8 pixels are exact in bf16):
//                                   conv1 (8x8 stride 4) becomes a 2x2 stride-1 conv over 16C channels
//   act1  [NF*100][128] split      conv1 output (20x20x32) stored space-to-depth by 2: row (f, y/2, x/2), column
//                                   (y&1, x&1, c) -- conv2 (4x4 stride 2) becomes a 2x2 stride-1 conv over 128 channels
//   act2  [NF*81][64]  act3 [NF][3136]     split, NHWC, post-ReLU  (f = b*T + t)
//   dpre3 [NF*81][64]  dpre2 [NF*100][64]  dpre1g [NF*441][32]   split pre-activation gradients, each on its layer's
//                                   INPUT grid (9x9, 10x10, 21x21); grid pixels that are no conv output stay zero, so the
//                                   conv layers' data / weight / bias gradients run as window convolutions (winconv.cuh)
//   U     [T*B][KU] split   LSTM input rows, time-major: latent(512) | one-hot last action(A) |
//                            last reward | zero pad, KU = roundup(512+A+1, 16)
//   XP    [T*B][4H] fp32    input projection incl. both biases, gate-interleaved (col = 4*j + gate)
//   Hs    [T*B][H]  split   hidden state after step t (frozen past a sequence's length)
//   Cs    [T*B][H]  fp32,  Gs [T*B][4H] fp32 (post-nonlinearity gates, slot 0 only, for BPTT)
//   hid   [2*Rmax][1024] split   dueling hidden layer (advantage | value) of the gathered rows
// Weights stay with the caller in the reference's state_dict layout (one flat fp32 buffer, see
// r2d2_net_param_layout); `pack` re-lays them out (NHWC / s2d im2col order, gate interleave) as split.
//
// The online pass is run ONCE for b+l+f steps: the reference's pass 1 (calculate_q_, no grad) and
// pass 3 (calculate_q, grad) share every hidden state up to b+l-1, so Q at the learning positions
// and at the n-step-shifted positions are two row gathers of the same unroll.
#include <algorithm>
#include <map>

#include "recurrence2.cuh"
#include "winconv.cuh"

namespace r2d2 {

constexpr int H = 512;          // config.hidden_dim (model.py:28); the kernels are specialised for it
constexpr int G4 = 4 * H;
constexpr int LATENT = 512;
constexpr int FLAT3 = 3136;     // 7*7*64
constexpr int NPARAM = 20;
constexpr int kDW = 32;          // slots per row of d(dueling output): A advantage gradients + 1 value gradient, A <= 31
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
    SplitW W1s, W2p, W3p, Wfcp, Wih_p, Whh_p, WhhT_p, Wh0, W3d, W2q;   // W2q: [128][256] conv2 dgrad; WhhT_p: [512][2048] transpose
    float *bias_p, *bh0;
};
struct Acts {
    SplitW act1, act2, act3, U, hid;
    SplitW HsX;          // [(T+1)*B][H]: block 0 holds the stored state h0, block t+1 the state after step t
    SplitW Hs;           // = HsX + B*H   (state after step t at row t*B + b)
    float *XP, *Cs, *Gs;
};
static inline SplitC ro(const SplitW& w) { return SplitC{w.hi, w.lo}; }

// One pending gradient reduction: `splits` partial tensors [M][N] (weights: routed into the reference layout by `kind`;
// biases: M = 1, kind = kFinBias + BiasKind).  All reductions of a backward pass are deferred into ONE launch per group
// (dense layers, conv layers) of finalize_grads_kernel instead of one or two small launches per parameter tensor.
constexpr int kFinMax = 20, kFinBias = 100;
struct FinSeg { const float* part; int splits, M, N, kind, sl, block0; float scale; long long o0, o1; };
struct FinArgs { FinSeg seg[kFinMax]; int nseg, nblocks; };

}  // namespace r2d2

struct r2d2_net {
    int B, T, C, A, Lmax, F, KU, NF, Rmax, KIH;
    int64_t off[r2d2::NPARAM + 1];
    r2d2::Packed pk[2];
    r2d2::Acts ac[2];
    r2d2::bf16* s2d;                 // shared by both slots: the staging buffer the next forward / backward reads (= s2d_buf[s2d_idx])
    r2d2::bf16* s2d_buf[2];          // [1] is allocated on first use: batch i+1 is gathered into it while update i still reads the other one
    int s2d_idx;
    r2d2::SplitW W1both;             // conv1 weights of both slots stacked [64][64C]: one launch shares the frame tile
    int *row_src, *len_full, *len_learn, *d_rows;     // row_src: [2*Rmax]  (q rows | shifted rows)
    unsigned int* rec_bar;                            // [2] step counters of the persistent recurrence
    // backward scratch
    r2d2::SplitW dhid, DG, dlat, dpre3, dpre2, dpre1g;   // dpre*: pre-activation grads on the layer's INPUT grid (9x9x64, 10x10x64, 21x21x32), junk pixels stay 0
    float *dH, *dhrec, *dcrec, *dout16, *ws, *colws, *rec_partial;
    size_t ws_floats, colws_floats, ws_used, colws_used;     // ws / colws are bump-allocated per backward pass (one region per pending reduction)
    r2d2::FinArgs pend;                                       // reductions launched by the next flush
    void* dense_grads_event;         // optional cudaEvent_t recorded in r2d2_net_backward once every non-conv gradient is final
    const float* hidden;             // last forward's stored state (caller-owned, alive until backward)
};

namespace r2d2 {

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void put_split(const SplitW& w, size_t i, float x) {
    const bf16 h = __float2bfloat16_rn(x);
    w.hi[i] = h;
    w.lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

__global__ void pack_kernel(const float* __restrict__ p, const int64_t* __restrict__ off, Packed pk, SplitW w1both, int which, int A,
                            int C, int KU) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int KIH = LATENT + A + 1;
    if (i < 32ll * 64 * C) {                     // conv1 in space-to-depth order: k = (dy*2+dx)*16C + c*16 + r*4 + q
        const int K1 = 64 * C, n = i / K1, k = i % K1, tap = k / (16 * C), ch = k % (16 * C);
        const int dy = tap >> 1, dx = tap & 1, c = ch >> 4, r = (ch >> 2) & 3, q = ch & 3;
        const float w = p[off[P_C1W] + (int64_t)n * K1 + c * 64 + (4 * dy + r) * 8 + 4 * dx + q];
        put_split(pk.W1s, i, w);
        put_split(w1both, (size_t)which * 32 * K1 + i, w);
    }
    if (i < 64 * 512) {                          // conv2 as a 2x2 stride-1 conv over act1 in space-to-depth-by-2 form:
        {                                        // k = (dy*2+dx)*128 + (ry*2+rx)*32 + c  <->  W2[n][c][2dy+ry][2dx+rx]
            const int n = i / 512, k = i % 512, tap = k >> 7, sub = (k >> 5) & 3, c = k & 31;
            const int ky = 2 * (tap >> 1) + (sub >> 1), kx = 2 * (tap & 1) + (sub & 1);
            put_split(pk.W2p, i, p[off[P_C2W] + n * 512 + c * 16 + ky * 4 + kx]);
        }
        // dgrad, 4 parity classes (= s2d-by-2 sub-pixels): W2q[cls*32 + c_in][(jy*2+jx)*64 + c_out] = W2[c_out][c_in][py+2jy][px+2jx]
        const int cls = i / (32 * 256), r = i % (32 * 256), ci = r / 256, kk = r % 256, j = kk >> 6, co = kk & 63;
        const int py = cls >> 1, px = cls & 1, jy = j >> 1, jx = j & 1;
        const float wd = p[off[P_C2W] + co * 512 + ci * 16 + (py + 2 * jy) * 4 + (px + 2 * jx)];
        put_split(pk.W2q, (size_t)(cls * 32 + ci) * 256 + kk, wd);      // all four parity classes side by side (N = 128 = s2d-by-2 channel)
    }
    if (i < 64 * 576) {
        const int n = i / 576, k = i % 576, tap = k >> 6, c = k & 63;
        put_split(pk.W3p, i, p[off[P_C3W] + n * 576 + c * 9 + tap]);
        put_split(pk.W3d, i, p[off[P_C3W] + c * 576 + n * 9 + tap]);   // W3d[c_in][(ky*3+kx)*64 + c_out]
    }
    if (i < 512ll * FLAT3) {                      // fc: col c*49+hw -> hw*64+c
        const int n = i / FLAT3, k = i % FLAT3, hw = k >> 6, c = k & 63;
        put_split(pk.Wfcp, i, p[off[P_FCW] + (int64_t)n * FLAT3 + c * 49 + hw]);
    }
    if (i < (int64_t)G4 * KU) {                   // W_ih: gate-interleaved rows, zero-padded cols
        const int np = i / KU, k = i % KU, g = np & 3, j = np >> 2;
        put_split(pk.Wih_p, i, (k < KIH) ? p[off[P_WIH] + (int64_t)(g * H + j) * KIH + k] : 0.f);
    }
    if (i < (int64_t)G4 * H) {
        const int np = i / H, k = i % H, g = np & 3, j = np >> 2;
        const float w = p[off[P_WHH] + (int64_t)(g * H + j) * H + k];
        put_split(pk.Whh_p, i, w);
        put_split(pk.WhhT_p, (size_t)k * G4 + np, w);          // transposed copy for the persistent BPTT kernel
    }
    if (i < G4) {
        const int g = i & 3, j = i >> 2;
        pk.bias_p[i] = p[off[P_BIH] + g * H + j] + p[off[P_BHH] + g * H + j];
    }
    if (i < 2 * H * H) put_split(pk.Wh0, i, (i < H * H) ? p[off[P_A0W] + i] : p[off[P_V0W] + i - H * H]);
    if (i < 2 * H) pk.bh0[i] = (i < H) ? p[off[P_A0B] + i] : p[off[P_V0B] + i - H];
}

// u8 frames (C,84,84) -> space-to-depth bf16 [f][Y][X][c*16 + r*4 + q],  pixel (c, 4Y+r, 4X+q).
// item = (f, c, Y, X): four coalesced 32-bit reads -> one 32-byte sector of bf16
__global__ void s2d_kernel(const uint8_t* __restrict__ obs, bf16* __restrict__ s2d, int C, int64_t total /* NF*C*441 */) {
    // two items per thread: eight independent loads in flight, one 256-bit store per item
    const int64_t i0 = (blockIdx.x * (int64_t)blockDim.x) * 2 + threadIdx.x;
    uint32_t w[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int64_t i = i0 + u * blockDim.x;
        if (i >= total) { w[u][0] = w[u][1] = w[u][2] = w[u][3] = 0u; continue; }
        const int X = i % 21, Y = (i / 21) % 21, c = (i / 441) % C;
        const int64_t f = i / (441 * (int64_t)C);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(obs) + ((f * C + c) * 84 + 4 * Y) * 21 + X;
#pragma unroll
        for (int r = 0; r < 4; ++r) w[u][r] = __ldg(src + r * 21);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int64_t i = i0 + u * blockDim.x;
        if (i >= total) return;
        const int X = i % 21, Y = (i / 21) % 21, c = (i / 441) % C;
        const int64_t f = i / (441 * (int64_t)C);
        uint32_t o[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const __nv_bfloat162 p0 = __floats2bfloat162_rn((float)(w[u][r] & 255u), (float)((w[u][r] >> 8) & 255u));
            const __nv_bfloat162 p1 = __floats2bfloat162_rn((float)((w[u][r] >> 16) & 255u), (float)(w[u][r] >> 24));
            o[2 * r] = *reinterpret_cast<const uint32_t*>(&p0);
            o[2 * r + 1] = *reinterpret_cast<const uint32_t*>(&p1);
        }
        st_global_256(s2d + ((f * 21 + Y) * 21 + X) * (16 * C) + c * 16, o);
    }
}

// row maps + sequence lengths (block 0) + split copy of h0 (other blocks).  model.py:102-111 (shifted rows) and model.py:143.
__global__ void prep_rows_kernel(const uint8_t* __restrict__ burn, const uint8_t* __restrict__ learn,
                                 const uint8_t* __restrict__ fwd, const float* __restrict__ hidden, SplitW h0a, SplitW h0b, int B, int T, int F,
                                 int Rmax, int* __restrict__ row_src, int* __restrict__ len_full, int* __restrict__ len_learn,
                                 int* __restrict__ d_rows) {
    extern __shared__ int s_off[];
    if (blockIdx.x > 0) {                                            // blocks 1..: h0 (split) in front of both slots' state arrays
        for (int i = (blockIdx.x - 1) * blockDim.x + threadIdx.x; i < B * H; i += (gridDim.x - 1) * blockDim.x) {
            const float x = hidden[(size_t)(i / H) * 2 * H + (i % H)];
            put_split(h0a, i, x);
            put_split(h0b, i, x);
        }
        return;
    }
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
        const int tot = min(b + l + f, T);                           // a sequence longer than the workspace is truncated, never read past it
        len_full[n] = tot;
        len_learn[n] = min(b + l, T);
        for (int i = 0; i < l && s_off[n] + i < Rmax; ++i) {
            row_src[s_off[n] + i] = min(b + i, T - 1) * B + n;
            row_src[Rmax + s_off[n] + i] = max(0, min(b + F + i, tot - 1)) * B + n;
        }
    }
}

// recurrent state after step t in the [B][2][H] layout of Block.hidden / the `hidden` input (worker.py:198,340)
__global__ void state_after_kernel(SplitC Hs, const float* __restrict__ Cs, int B, int t, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, j = i - b * H;
    const size_t o = ((size_t)t * B + b) * H + j;
    out[(size_t)b * 2 * H + j] = split_load(Hs.hi, Hs.lo, o);
    out[(size_t)b * 2 * H + H + j] = Cs[o];
}

// U side columns: one-hot last action, last reward, zero pad  (model.py:92)
__global__ void side_columns_kernel(SplitW U, SplitW U2 /* second slot or {nullptr, nullptr} */, const uint8_t* __restrict__ last_action,
                                    const float* __restrict__ last_reward, int B, int T, int A, int KU) {
    const int row = blockIdx.x;                 // time-major row t*B + b
    const int t = row / B, b = row % B;
    const int f = b * T + t;
    for (int k = LATENT + threadIdx.x; k < KU; k += blockDim.x) {
        float v = 0.f;
        if (k < LATENT + A) v = last_action[(size_t)f * A + (k - LATENT)] ? 1.f : 0.f;
        else if (k == LATENT + A) v = last_reward[f];
        put_split(U, (size_t)row * KU + k, v);
        if (U2.hi) put_split(U2, (size_t)row * KU + k, v);
    }
}

// dueling output layer: one warp per row.  model.py:115-117
struct HeadJob { SplitC hid; size_t row_offset; const float *Wa2, *ba2, *Wv2, *bv2; float* q_out; };
struct HeadJobs { HeadJob job[3]; };
__global__ void head_out_kernel(const HeadJobs jobs, int rows_cap, int A) {
    const HeadJob& jb = jobs.job[blockIdx.y];
    const SplitC hid = jb.hid;
    const size_t row_offset = jb.row_offset;
    const float *Wa2 = jb.Wa2, *ba2 = jb.ba2, *Wv2 = jb.Wv2, *bv2 = jb.bv2;
    float* q_out = jb.q_out;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows_cap) return;
    const size_t base = (row_offset + warp) * 2 * H;
    // lane owns 16 consecutive hidden units of each branch: 16-byte loads of the split planes and of the weight rows
    constexpr int PER = H / 32;                                      // 16
    float ha[PER], hv[PER];
    split_load8(hid.hi, hid.lo, base + lane * PER, ha);
    split_load8(hid.hi, hid.lo, base + lane * PER + 8, ha + 8);
    split_load8(hid.hi, hid.lo, base + H + lane * PER, hv);
    split_load8(hid.hi, hid.lo, base + H + lane * PER + 8, hv + 8);
    float mine = 0.f, sum = 0.f;                                      // lane a keeps advantage a (A <= 32)
    for (int a = 0; a < A; ++a) {
        float s = 0.f;
        const float4* w = reinterpret_cast<const float4*>(Wa2 + a * H + lane * PER);
#pragma unroll
        for (int i = 0; i < PER / 4; ++i) {
            const float4 x = __ldg(w + i);
            s = fmaf(ha[4 * i], x.x, s); s = fmaf(ha[4 * i + 1], x.y, s); s = fmaf(ha[4 * i + 2], x.z, s); s = fmaf(ha[4 * i + 3], x.w, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s += ba2[a];
        if (lane == a) mine = s;
        sum += s;
    }
    float v = 0.f;
    {
        const float4* w = reinterpret_cast<const float4*>(Wv2 + lane * PER);
#pragma unroll
        for (int i = 0; i < PER / 4; ++i) {
            const float4 x = __ldg(w + i);
            v = fmaf(hv[4 * i], x.x, v); v = fmaf(hv[4 * i + 1], x.y, v); v = fmaf(hv[4 * i + 2], x.z, v); v = fmaf(hv[4 * i + 3], x.w, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    v += bv2[0];
    const float mean = sum / (float)A;
    if (lane < A) q_out[(size_t)warp * A + lane] = v + mine - mean;
}

// backward of the dueling output layer: dq -> (dadv | dval) and masked d(hidden layer)
__global__ void head_out_bwd_kernel(const float* __restrict__ dq, SplitC hid, const float* __restrict__ Wa2,
                                    const float* __restrict__ Wv2, const int* __restrict__ d_rows, int A,
                                    float* __restrict__ dout16, SplitW dhid) {
    const int r = blockIdx.x;
    __shared__ float s_d[kDW];
    const bool live = r < *d_rows;
    if (threadIdx.x < kDW) {
        float v = 0.f;
        if (live) {
            float tot = 0.f;
            for (int a = 0; a < A; ++a) tot += dq[(size_t)r * A + a];
            const int a = threadIdx.x;
            if (a < A) v = dq[(size_t)r * A + a] - tot / (float)A;     // d adv (through the mean)
            else if (a == A) v = tot;                                   // d val
        }
        s_d[threadIdx.x] = v;
        dout16[(size_t)r * kDW + threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * H; j += blockDim.x) {
        float g = 0.f;
        const size_t o = (size_t)r * 2 * H + j;
        if (live && (__bfloat16_as_ushort(hid.hi[o]) & 0x7FFFu)) {
            if (j < H) { for (int a = 0; a < A; ++a) g = fmaf(s_d[a], __ldg(Wa2 + a * H + j), g); }
            else g = s_d[A] * __ldg(Wv2 + j - H);
        }
        put_split(dhid, o, g);
    }
}

// layer-2 head weight gradients: [A+1 (padded to kDW)][1024] = dout^T . hid  (CUDA cores, row-chunk partials)
__global__ void head_w2_grad_kernel(const float* __restrict__ dout16, SplitC hid, int Rmax, int chunk, float* __restrict__ ws) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;          // column of hid, 0..1023
    const int p = blockIdx.y;
    const int r0 = p * chunk, r1 = min(Rmax, r0 + chunk);
    float acc[kDW];
#pragma unroll
    for (int a = 0; a < kDW; ++a) acc[a] = 0.f;
    for (int r = r0; r < r1; ++r) {
        const float h = split_load(hid.hi, hid.lo, (size_t)r * 2 * H + j);
        const float4* d = reinterpret_cast<const float4*>(dout16 + (size_t)r * kDW);
#pragma unroll
        for (int q = 0; q < kDW / 4; ++q) {
            const float4 dv = __ldg(d + q);
            acc[4 * q] = fmaf(dv.x, h, acc[4 * q]); acc[4 * q + 1] = fmaf(dv.y, h, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(dv.z, h, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(dv.w, h, acc[4 * q + 3]);
        }
    }
#pragma unroll
    for (int a = 0; a < kDW; ++a) ws[((size_t)p * kDW + a) * 2 * H + j] = acc[a];
}

// LSTM cell epilogue fused into the recurrent GEMM (gate-interleaved columns: n = 4*j + gate)
struct Epi2LstmCell {
    const float* xp_t; const float* c_prev; int ld_cprev; SplitC h_prev;
    SplitW h_out; float* c_out; float* gates_out; const int* len; int t, B;
    __device__ __forceinline__ void store16(int m, int n, const float (&a)[16], int) const {
        if (m >= B || n >= G4) return;
        const int j0 = n >> 2;
        const bool live = t < len[m];
        float hn4[4], cn4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 x = *reinterpret_cast<const float4*>(xp_t + (size_t)m * G4 + n + 4 * u);
            const float gi = 1.f / (1.f + expf(-(a[4 * u] + x.x)));
            const float gf = 1.f / (1.f + expf(-(a[4 * u + 1] + x.y)));
            const float gg = tanhf(a[4 * u + 2] + x.z);
            const float go = 1.f / (1.f + expf(-(a[4 * u + 3] + x.w)));
            const float cp = c_prev[(size_t)m * ld_cprev + j0 + u];
            const float cn = gf * cp + gi * gg;
            const float hn = go * tanhf(cn);
            if (gates_out) *reinterpret_cast<float4*>(gates_out + (size_t)m * G4 + n + 4 * u) = make_float4(gi, gf, gg, go);
            cn4[u] = live ? cn : cp;
            hn4[u] = live ? hn : split_load(h_prev.hi, h_prev.lo, (size_t)m * H + j0 + u);
        }
        *reinterpret_cast<float4*>(c_out + (size_t)m * H + j0) = make_float4(cn4[0], cn4[1], cn4[2], cn4[3]);
        uint32_t h[2], l[2];
        split2(hn4[0], hn4[1], h[0], l[0]);
        split2(hn4[2], hn4[3], h[1], l[1]);
        *reinterpret_cast<uint2*>(h_out.hi + (size_t)m * H + j0) = make_uint2(h[0], h[1]);
        *reinterpret_cast<uint2*>(h_out.lo + (size_t)m * H + j0) = make_uint2(l[0], l[1]);
    }
};

// BPTT pointwise step: d(pre-activation gates) at time t
__global__ void lstm_bwd_pointwise_kernel(const float* __restrict__ dH_t, const float* __restrict__ dhrec, int nparts,
                                          float* __restrict__ dcrec, const float* __restrict__ G_t,
                                          const float* __restrict__ C_t, const float* __restrict__ C_prev, int ld_cprev,
                                          const int* __restrict__ len_learn, int t, int B, SplitW DG, size_t dg_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, j = i % H;
    float out[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < len_learn[b]) {
        const float4 g = *reinterpret_cast<const float4*>(G_t + (size_t)b * G4 + 4 * j);
        const float gi = g.x, gf = g.y, gg = g.z, go = g.w;
        const float tc = tanhf(C_t[i]);
        float dh = dH_t[i];
        for (int p = 0; p < nparts; ++p) dh += dhrec[(size_t)p * B * H + i];      // split-K partials of dgates.W_hh
        const float dc = dcrec[i] + dh * go * (1.f - tc * tc);
        const float cp = C_prev[(size_t)b * ld_cprev + j];
        out[0] = dc * gg * gi * (1.f - gi);
        out[1] = dc * cp * gf * (1.f - gf);
        out[2] = dc * gi * (1.f - gg * gg);
        out[3] = dh * tc * go * (1.f - go);
        dcrec[i] = dc * gf;
    }
    uint32_t h[2], l[2];
    split2(out[0], out[1], h[0], l[0]);
    split2(out[2], out[3], h[1], l[1]);
    const size_t o = dg_off + (size_t)b * G4 + 4 * j;
    *reinterpret_cast<uint2*>(DG.hi + o) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(DG.lo + o) = make_uint2(l[0], l[1]);
}

// FC epilogue: latent -> U (time-major rows), ReLU
struct Epi2Latent {
    SplitW U; const float* bias; int B, T, KU;
    __device__ __forceinline__ void store16(int m, int n, const float (&v)[16], int) const {
        if (m >= B * T || n >= LATENT) return;
        const int b = m / T, t = m - b * T;
        const size_t o = ((size_t)t * B + b) * KU + n;
        float r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaxf(v[i] + __ldg(bias + n + i), 0.f);
        split_store16(U.hi, U.lo, o, r);                                        // KU % 16 == 0
    }
};
// d latent epilogue: rows are time-major (t,b); mask by latent>0; write frame-major for the encoder backward
struct Epi2DLatent {
    SplitW dlat; SplitC U; int B, T, KU;
    __device__ __forceinline__ void store16(int m, int n, const float (&v)[16], int) const {
        if (m >= B * T || n >= LATENT) return;
        const int t = m / B, b = m - t * B;
        float r[16];
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            const uint4 h = *reinterpret_cast<const uint4*>(U.hi + (size_t)m * KU + n + j);
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[j + 2 * i] = (hw[i] & 0x7FFFu) ? v[j + 2 * i] : 0.f;
                r[j + 2 * i + 1] = (hw[i] & 0x7FFF0000u) ? v[j + 2 * i + 1] : 0.f;
            }
        }
        split_store16(dlat.hi, dlat.lo, ((size_t)b * T + t) * LATENT + n, r);
    }
};
// FC data gradient -> dpre3 on conv3's INPUT grid (9x9, the 7x7 valid outputs at gy,gx < 7; the rest stays zero), masked
// by act3 > 0.  Column n = hw*64 + c of frame m.
struct Epi2MaskedToGrid3 {
    SplitW out; SplitC act; int M;
    __device__ __forceinline__ void store16(int m, int n, const float (&v)[16], int) const {
        if (m >= M || n >= FLAT3) return;
        const int hw = n >> 6, c = n & 63, oy = hw / 7, ox = hw - oy * 7;
        const size_t o = ((size_t)m * 81 + oy * 9 + ox) * 64 + c, a = (size_t)m * FLAT3 + n;
        float r[16];
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            const uint4 h = *reinterpret_cast<const uint4*>(act.hi + a + j);
            const uint32_t hw4[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[j + 2 * i] = (hw4[i] & 0x7FFFu) ? v[j + 2 * i] : 0.f;
                r[j + 2 * i + 1] = (hw4[i] & 0x7FFF0000u) ? v[j + 2 * i + 1] : 0.f;
            }
        }
        split_store16(out.hi, out.lo, o, r);
    }
};
// scatter rows of d(hidden rows) to dH[t][b] (fp32) through the row map
struct Epi2ScatterRows {
    float* dH; const int* src; const int* d_rows; int Rmax;
    __device__ __forceinline__ void store16(int m, int n, const float (&v)[16], int) const {
        if (m >= Rmax || n >= H || m >= *d_rows) return;
        const int s = src[m];
        if (s < 0) return;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dH + (size_t)s * H + n + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
};

// routing of a reduced weight-gradient entry (m, n) into the reference's parameter layout
enum RouteKind { R_C1, R_C2, R_C3, R_FC, R_WIH, R_WHH, R_H0, R_H2, R_C1W, R_C2W, R_C3W };   // R_C*W: window wgrad partials [tap*IC + c][out channel]
enum BiasKind { B_PLAIN, B_LSTM, B_H0, B_H2 };
__device__ __forceinline__ void route_weight(int kind, int m, int n, float s, float* __restrict__ g, const int64_t* __restrict__ off, int A, int C) {
    if (kind >= R_C1W) { const int t = m; m = n; n = t; kind = kind == R_C1W ? R_C1 : kind == R_C2W ? R_C2 : R_C3; }   // (k, out) -> (out, k)
    const int KIH = LATENT + A + 1;
    switch (kind) {
        case R_C1: {   // n = (dy*2+dx)*16C + c*16 + r*4 + q  ->  [m][c][4dy+r][4dx+q]
            const int tap = n / (16 * C), ch = n % (16 * C), dy = tap >> 1, dx = tap & 1, c = ch >> 4, r = (ch >> 2) & 3, q = ch & 3;
            g[off[P_C1W] + (int64_t)m * 64 * C + c * 64 + (4 * dy + r) * 8 + 4 * dx + q] = s;
        } break;
        case R_C2: {   // n = (dy*2+dx)*128 + (ry*2+rx)*32 + c  ->  [m][c][2dy+ry][2dx+rx]
            const int tap = n >> 7, sub = (n >> 5) & 3, c = n & 31, ky = 2 * (tap >> 1) + (sub >> 1), kx = 2 * (tap & 1) + (sub & 1);
            g[off[P_C2W] + m * 512 + c * 16 + ky * 4 + kx] = s;
        } break;
        case R_C3: { const int tap = n >> 6, c = n & 63; g[off[P_C3W] + m * 576 + c * 9 + tap] = s; } break;
        case R_FC: { const int hw = n >> 6, c = n & 63; g[off[P_FCW] + (int64_t)m * FLAT3 + c * 49 + hw] = s; } break;
        case R_WIH: { const int row = (m & 3) * H + (m >> 2); if (n < KIH) g[off[P_WIH] + (int64_t)row * KIH + n] = s; } break;
        case R_WHH: { const int row = (m & 3) * H + (m >> 2); g[off[P_WHH] + (int64_t)row * H + n] = s; } break;
        case R_H0: { if (m < H) g[off[P_A0W] + (int64_t)m * H + n] = s; else g[off[P_V0W] + (int64_t)(m - H) * H + n] = s; } break;
        case R_H2: { if (m < A && n < H) g[off[P_A2W] + m * H + n] = s; else if (m == A && n >= H) g[off[P_V2W] + n - H] = s; } break;
    }
}
__device__ __forceinline__ void route_bias(int kind, int n, float s, float* __restrict__ g, int64_t o0, int64_t o1, int A) {
    switch (kind) {
        case B_PLAIN: g[o0 + n] = s; break;
        case B_LSTM: { const int row = (n & 3) * H + (n >> 2); g[o0 + row] = s; g[o1 + row] = s; } break;   // b_ih and b_hh
        case B_H0: { if (n < H) g[o0 + n] = s; else g[o1 + n - H] = s; } break;
        case B_H2: { if (n < A) g[o0 + n] = s; else if (n == A) g[o1] = s; } break;
    }
}
// All pending split reductions of one group in one launch.  Block -> segment through the block0 table; sl = 8: a block
// reduces 32 outputs x 8 slices of the split index (many splits, few outputs: the window wgrads and every bias), sl = 1:
// one thread per output.  Fixed summation order either way (deterministic).
__global__ void __launch_bounds__(256) finalize_grads_kernel(const FinArgs fa, float* __restrict__ g, const int64_t* __restrict__ off, int A, int C) {
    int si = 0;
    for (int i = 1; i < fa.nseg; ++i)
        if ((int)blockIdx.x >= fa.seg[i].block0) si = i;
    const FinSeg& sg = fa.seg[si];
    const int blk = blockIdx.x - sg.block0;
    const int64_t MN = (int64_t)sg.M * sg.N;
    int64_t i;
    float s = 0.f;
    if (sg.sl == 1) {
        i = blk * 256ll + threadIdx.x;
        if (i >= MN) return;
#pragma unroll 4
        for (int z = 0; z < sg.splits; ++z) s += __ldcs(sg.part + (size_t)z * MN + i);
    } else {
        __shared__ float sm[8][33];
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        i = blk * 32ll + tx;
        const bool in = i < MN;
        float acc = 0.f;
        if (in) {                                   // eight independent loads in flight per thread (the chain is latency-bound otherwise)
            const float* pp = sg.part + i;
            float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int z = ty;
            for (; z + 56 < sg.splits; z += 64) {
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] += __ldcs(pp + (size_t)(z + 8 * u) * MN);
            }
            for (; z < sg.splits; z += 8) a[0] += __ldcs(pp + (size_t)z * MN);
            acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
        }
        sm[ty][tx] = acc;
        __syncthreads();
        if (ty != 0 || !in) return;
#pragma unroll
        for (int y = 0; y < 8; ++y) s += sm[y][tx];
    }
    s *= sg.scale;
    if (sg.kind >= kFinBias) route_bias(sg.kind - kFinBias, (int)i, s, g, sg.o0, sg.o1, A);
    else route_weight(sg.kind, (int)(i / sg.N), (int)(i % sg.N), s, g, off, A, C);
}

// deterministic column sums of a (split or fp32) [M][N] tensor: partial[p][n] over row chunk p, then final routing.
// Split tensors are read 8 columns (16 bytes per plane) at a time: N % 8 == 0.
template <bool kSplit>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, SplitC S, int M, int N, int chunk,
                                                             float* __restrict__ part) {
    const int p = blockIdx.y;
    const int r0 = p * chunk, r1 = min(M, r0 + chunk);
    if constexpr (kSplit) {
        __shared__ float s[256][9];
        const int groups = N >> 3;                               // 8-column groups per row
        const int gpb = min(groups, 32);                         // groups handled by one block (x dimension)
        const int g = blockIdx.x * gpb + (threadIdx.x % gpb);
        const int rl = threadIdx.x / gpb, rstep = 256 / gpb;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (g < groups)
            for (int r = r0 + rl; r < r1; r += rstep) {
                float v[8];
                split_load8(S.hi, S.lo, (size_t)r * N + g * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += v[i];
            }
#pragma unroll
        for (int i = 0; i < 8; ++i) s[threadIdx.x][i] = acc[i];
        __syncthreads();
        if (rl == 0 && g < groups) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float t = 0.f;
                for (int y = 0; y < rstep; ++y) t += s[y * gpb + (threadIdx.x % gpb)][i];
                part[(size_t)p * N + g * 8 + i] = t;
            }
        }
    } else {
        __shared__ float s[8][33];
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        const int col = blockIdx.x * 32 + tx;
        float acc = 0.f;
        if (col < N)
            for (int r = r0 + ty; r < r1; r += 8) acc += X[(size_t)r * N + col];
        s[ty][tx] = acc;
        __syncthreads();
        if (ty == 0 && col < N) {
            float t = 0.f;
            for (int y = 0; y < 8; ++y) t += s[y][tx];
            part[(size_t)p * N + col] = t;
        }
    }
}
constexpr int kColP = 128;

}  // namespace r2d2

using namespace r2d2;

// ---- deferred reductions: bump allocation of the partial workspaces + the pending-segment list of a handle
static float* ws_take(r2d2_net* n, size_t floats) {
    const size_t a = (n->ws_used + 63) & ~(size_t)63;
    if (a + floats > n->ws_floats) return nullptr;
    n->ws_used = a + floats;
    return n->ws + a;
}
static float* colws_take(r2d2_net* n, size_t floats) {
    const size_t a = (n->colws_used + 63) & ~(size_t)63;
    if (a + floats > n->colws_floats) return nullptr;
    n->colws_used = a + floats;
    return n->colws + a;
}
static cudaError_t fin_add(r2d2_net* n, const float* part, int splits, int M, int N, int kind, float scale, int64_t o0 = 0, int64_t o1 = 0) {
    FinArgs& fa = n->pend;
    if (fa.nseg >= kFinMax) return cudaErrorInvalidValue;
    const int sl = (splits >= 16 || kind >= kFinBias) ? 8 : 1;
    FinSeg& sg = fa.seg[fa.nseg++];
    sg = FinSeg{part, splits, M, N, kind, sl, fa.nblocks, scale, (long long)o0, (long long)o1};
    fa.nblocks += cdiv((int64_t)M * N, sl == 8 ? 32 : 256);
    return cudaSuccess;
}
static cudaError_t fin_flush(r2d2_net* n, float* grads, const int64_t* d_off, cudaStream_t s) {
    FinArgs& fa = n->pend;
    if (fa.nseg == 0) return cudaSuccess;
    finalize_grads_kernel<<<fa.nblocks, 256, 0, s>>>(fa, grads, d_off, n->A, n->C);
    fa.nseg = 0;
    fa.nblocks = 0;
    return cudaGetLastError();
}
static void fin_begin(r2d2_net* n) { n->ws_used = 0; n->colws_used = 0; n->pend.nseg = 0; n->pend.nblocks = 0; }

// column sums (bias gradients) of a split / fp32 [M][N] tensor: partial rows now, final reduction deferred
static cudaError_t colsum_split(SplitC S, int M, int N, int kind, int64_t o0, int64_t o1, r2d2_net* n, cudaStream_t s) {
    float* part = colws_take(n, (size_t)kColP * N);
    if (!part) return cudaErrorInvalidValue;
    const int chunk = (M + kColP - 1) / kColP;
    const int groups = N / 8, gpb = groups < 32 ? groups : 32;
    dim3 grid((groups + gpb - 1) / gpb, kColP);
    colsum_partial_kernel<true><<<grid, 256, 0, s>>>(nullptr, S, M, N, chunk, part);
    cudaError_t e = cudaGetLastError();
    return e != cudaSuccess ? e : fin_add(n, part, kColP, 1, N, kFinBias + kind, 1.f, o0, o1);
}
static cudaError_t colsum_f32(const float* X, int M, int N, int kind, int64_t o0, int64_t o1, r2d2_net* n, cudaStream_t s) {
    float* part = colws_take(n, (size_t)kColP * N);
    if (!part) return cudaErrorInvalidValue;
    const int chunk = (M + kColP - 1) / kColP;
    dim3 grid((N + 31) / 32, kColP);
    colsum_partial_kernel<false><<<grid, 256, 0, s>>>(X, SplitC{nullptr, nullptr}, M, N, chunk, part);
    cudaError_t e = cudaGetLastError();
    return e != cudaSuccess ? e : fin_add(n, part, kColP, 1, N, kFinBias + kind, 1.f, o0, o1);
}

template <int UBN, int POL = LO_STRICT, class AS, class BS>
static cudaError_t wgrad2(const AS& a, const BS& b, int M, int N, int K, int splits, int kind, r2d2_net* net, float scale, cudaStream_t s) {
    float* ws = ws_take(net, (size_t)splits * M * N);
    if (!ws) return cudaErrorInvalidValue;
    Epi2Partial ep{ws, M, N};
    cudaError_t e = launch_umma2<UBN, POL>(a, b, ep, M, N, K, splits, s);
    return e != cudaSuccess ? e : fin_add(net, ws, splits, M, N, kind, scale);
}

// plain-matrix weight gradient on CTA pairs (umma3.cuh): both operands MN-major ([K][rows] storage)
template <int POL = LO_STRICT>
static cudaError_t wgrad3(const Mat3& a, const Mat3& b, int M, int N, int K, int splits, int kind, r2d2_net* net, float scale, cudaStream_t s) {
    splits = umma3_effective_splits(K, splits);
    float* ws = ws_take(net, (size_t)splits * M * N);
    if (!ws) return cudaErrorInvalidValue;
    Epi2Partial ep{ws, M, N};
    cudaError_t e = launch_umma3<true, true, POL>(a, b, ep, M, N, K, splits, s);
    return e != cudaSuccess ? e : fin_add(net, ws, splits, M, N, kind, scale);
}

// 1 (default): the plain-matrix GEMMs (FC, input projection, their data / weight gradients) run on CTA pairs with TMA
// (umma3.cuh); 0: single-CTA cp.async kernels (umma2.cuh).  R2D2_PAIR_GEMM=0 in the environment or r2d2_set_pair_gemm.
int g_pair_gemm = [] { const char* e = getenv("R2D2_PAIR_GEMM"); return (e && e[0] == '0') ? 0 : 1; }();

// 1 (default): forward recurrence inside 16-CTA clusters (recurrence2.cuh: W_hh resident in TMEM + shared memory, h exchanged
// over distributed shared memory); 0: the L2-flag persistent kernel of recurrence.cuh.  R2D2_CLUSTER_REC=0 / r2d2_set_cluster_recurrence.
int g_cluster_recurrence = [] { const char* e = getenv("R2D2_CLUSTER_REC"); return (e && e[0] == '0') ? 0 : 1; }();
namespace r2d2 { int g_config_epoch = 0; }
namespace r2d2 { int g_rec2_ns = [] { const char* e = getenv("R2D2_REC2_NS"); return e ? atoi(e) : 0; }(); }   // sequences per recurrence cluster: 0 auto, 16, 32
int g_persistent_recurrence = 1;
unsigned long long* g_rec_trace_bwd = nullptr; // same for the cluster BPTT kernel (r2d2_debug_rec_trace_bwd)
unsigned long long* g_rec_trace = nullptr;   // debug: device buffer [T][8] of step timestamps (r2d2_debug_rec_trace)     // 0: per-step launches (also the path for B > 64)

// device copy of the parameter offsets, kept in a side table keyed by handle
static std::map<r2d2_net*, int64_t*> g_doff;

static int alloc_f(float** p, size_t n) {
    R2D2_CUDA_CHECK(cudaMalloc(p, n * sizeof(float)));
    R2D2_CUDA_CHECK(cudaMemset(*p, 0, n * sizeof(float)));
    return R2D2_OK;
}
static int alloc_s(SplitW* w, size_t n) {
    R2D2_CUDA_CHECK(cudaMalloc(&w->hi, n * sizeof(bf16)));
    R2D2_CUDA_CHECK(cudaMalloc(&w->lo, n * sizeof(bf16)));
    R2D2_CUDA_CHECK(cudaMemset(w->hi, 0, n * sizeof(bf16)));
    R2D2_CUDA_CHECK(cudaMemset(w->lo, 0, n * sizeof(bf16)));
    return R2D2_OK;
}
static void free_s(SplitW& w) { cudaFree(w.hi); cudaFree(w.lo); }

extern "C" {

int r2d2_net_param_layout(int A, int C, int64_t* offsets_out /* [21] */) {
    R2D2_REQUIRE(offsets_out && A >= 1 && A <= 31 && C >= 1 && C <= 16, "bad arguments");
    int64_t n[NPARAM];
    param_sizes(A, C, n);
    int64_t o = 0;
    for (int i = 0; i < NPARAM; ++i) { offsets_out[i] = o; o += (n[i] + 3) / 4 * 4; }   // 16-byte aligned tensors
    offsets_out[NPARAM] = o;
    return R2D2_OK;
}

int r2d2_net_create(int B, int T, int C, int A, int Lmax, int max_forward, r2d2_net** out) {
    R2D2_REQUIRE(out && B >= 1 && B <= 4096 && T >= 1 && T <= 255 && (C == 1 || C == 4) && A >= 1 && A <= 31 && Lmax >= 1 &&
                     Lmax <= T && max_forward >= 0,
                 "bad shape (frame channels must be 1 or 4, action_dim <= 31: the full ALE action set has 18)");
    r2d2_net* n = new r2d2_net();
    memset(n, 0, sizeof(*n));
    n->B = B; n->T = T; n->C = C; n->A = A; n->Lmax = Lmax; n->F = max_forward;
    n->KIH = LATENT + A + 1;
    n->KU = (n->KIH + 15) / 16 * 16;
    n->NF = B * T;
    n->Rmax = (B * Lmax + 7) / 8 * 8;
    r2d2_net_param_layout(A, C, n->off);
    int64_t* d_off = nullptr;
    R2D2_CUDA_CHECK(cudaMalloc(&d_off, sizeof(n->off)));
    R2D2_CUDA_CHECK(cudaMemcpy(d_off, n->off, sizeof(n->off), cudaMemcpyHostToDevice));
    g_doff[n] = d_off;
    const size_t NF = n->NF, TB = (size_t)T * B;
    int rc = 0;
    for (int k = 0; k < 2 && !rc; ++k) {
        Packed& p = n->pk[k];
        rc |= alloc_s(&p.W1s, 32ull * 64 * C); rc |= alloc_s(&p.W2p, 64 * 512); rc |= alloc_s(&p.W3p, 64 * 576);
        rc |= alloc_s(&p.Wfcp, 512ull * FLAT3); rc |= alloc_s(&p.Wih_p, (size_t)G4 * n->KU); rc |= alloc_s(&p.Whh_p, (size_t)G4 * H); rc |= alloc_s(&p.WhhT_p, (size_t)G4 * H);
        rc |= alloc_s(&p.Wh0, 2 * H * H); rc |= alloc_s(&p.W3d, 64 * 576); rc |= alloc_s(&p.W2q, 128 * 256);
        rc |= alloc_f(&p.bias_p, G4); rc |= alloc_f(&p.bh0, 2 * H);
        Acts& a = n->ac[k];
        rc |= alloc_s(&a.act1, NF * 12800); rc |= alloc_s(&a.act2, NF * 5184); rc |= alloc_s(&a.act3, NF * FLAT3);
        rc |= alloc_s(&a.U, TB * n->KU); rc |= alloc_s(&a.HsX, (TB + B) * H); rc |= alloc_s(&a.hid, (size_t)2 * n->Rmax * 2 * H);
        a.Hs = SplitW{a.HsX.hi + (size_t)B * H, a.HsX.lo + (size_t)B * H};
        rc |= alloc_f(&a.XP, TB * G4); rc |= alloc_f(&a.Cs, TB * H); rc |= alloc_f(&a.Gs, k == 0 ? TB * G4 : 4);
    }
    if (rc) return rc;
    R2D2_CUDA_CHECK(cudaMalloc(&n->s2d, (NF * 441 + 32) * 16 * C * sizeof(bf16)));     // + slack rows: the C = 1 wgrad reads taps of junk pixels
    R2D2_CUDA_CHECK(cudaMemset(n->s2d, 0, (NF * 441 + 32) * 16 * C * sizeof(bf16)));
    n->s2d_buf[0] = n->s2d; n->s2d_buf[1] = nullptr; n->s2d_idx = 0;
    rc |= alloc_s(&n->W1both, 64ull * 64 * C);
    R2D2_CUDA_CHECK(cudaMalloc(&n->row_src, 2 * n->Rmax * sizeof(int)));
    R2D2_CUDA_CHECK(cudaMalloc(&n->len_full, B * sizeof(int)));
    R2D2_CUDA_CHECK(cudaMalloc(&n->len_learn, B * sizeof(int)));
    R2D2_CUDA_CHECK(cudaMalloc(&n->d_rows, sizeof(int)));
    R2D2_CUDA_CHECK(cudaMalloc(&n->rec_bar, 64 * sizeof(unsigned int)));
    rc |= alloc_s(&n->dhid, (size_t)n->Rmax * 2 * H); rc |= alloc_s(&n->DG, TB * G4); rc |= alloc_s(&n->dlat, NF * LATENT);
    rc |= alloc_s(&n->dpre3, NF * 5184); rc |= alloc_s(&n->dpre2, NF * 6400); rc |= alloc_s(&n->dpre1g, NF * 14112);   // gradient grids (see struct)
    rc |= alloc_f(&n->dH, TB * H); rc |= alloc_f(&n->dhrec, (size_t)kRecSplits * B * H); rc |= alloc_f(&n->dcrec, (size_t)B * H);
    rc |= alloc_f(&n->dout16, (size_t)n->Rmax * kDW); rc |= alloc_f(&n->rec_partial, 2ull * 8 * 64 * H);
    {   // partial workspaces of one backward pass: every pending reduction keeps its own region until the flush (sizes as in
        // r2d2_net_backward: split counts of the dense weight gradients, one partial per <= 4096-pixel chunk of the conv ones)
        auto chunks = [](size_t rows, size_t chunk) { return (rows + chunk - 1) / chunk; };
        auto up = [](size_t x) { return (x + 63) & ~(size_t)63; };
        const size_t spl = std::max<size_t>(4, chunks(TB, 4096));
        size_t need = up(128ull * kDW * 2 * H) + up(4ull * 2 * H * H) + up(spl * G4 * H) + up(spl * G4 * n->KU) + up(spl * LATENT * FLAT3);
        need += up(chunks(NF * 81, 3072) * 576 * 64) + up(chunks(NF * 100, 3712) * 512 * 64);
        need += up(std::max(chunks(NF * 441, 4096) * 256 * 32, chunks(NF * 441, 4096) * 32 * 64 * (size_t)C));
        n->ws_floats = need + 1024;
        n->colws_floats = up((size_t)kColP * kDW) + up((size_t)kColP * 2 * H) + up((size_t)kColP * G4) + up((size_t)kColP * LATENT) + up((size_t)kColP * 32) +
                          up(chunks(NF * 81, 3072) * 64) + up(chunks(NF * 100, 3712) * 64) + up(chunks(NF * 441, 4096) * 32) + 1024;
    }
    rc |= alloc_f(&n->ws, n->ws_floats); rc |= alloc_f(&n->colws, n->colws_floats);
    if (rc) return rc;
    R2D2_CUDA_CHECK(cudaDeviceSynchronize());
    *out = n;
    return R2D2_OK;
}

int r2d2_net_destroy(r2d2_net* n) {
    if (!n) return R2D2_OK;
    for (int k = 0; k < 2; ++k) {
        Packed& p = n->pk[k];
        SplitW* ps[] = {&p.W1s, &p.W2p, &p.W3p, &p.Wfcp, &p.Wih_p, &p.Whh_p, &p.WhhT_p, &p.Wh0, &p.W3d, &p.W2q};
        for (SplitW* x : ps) free_s(*x);
        cudaFree(p.bias_p); cudaFree(p.bh0);
        Acts& a = n->ac[k];
        SplitW* as[] = {&a.act1, &a.act2, &a.act3, &a.U, &a.HsX, &a.hid};
        for (SplitW* x : as) free_s(*x);
        cudaFree(a.XP); cudaFree(a.Cs); cudaFree(a.Gs);
    }
    SplitW* ss[] = {&n->W1both, &n->dhid, &n->DG, &n->dlat, &n->dpre3, &n->dpre2, &n->dpre1g};
    for (SplitW* x : ss) free_s(*x);
    float* fs[] = {n->dH, n->dhrec, n->dcrec, n->dout16, n->ws, n->colws, n->rec_partial};
    for (float* x : fs) cudaFree(x);
    cudaFree(n->rec_bar); cudaFree(n->s2d_buf[0]); cudaFree(n->s2d_buf[1]); cudaFree(n->row_src); cudaFree(n->len_full); cudaFree(n->len_learn); cudaFree(n->d_rows);
    cudaFree(g_doff[n]);
    g_doff.erase(n);
    delete n;
    return R2D2_OK;
}

int r2d2_net_rows_capacity(const r2d2_net* n) { return n ? n->Rmax : -1; }
/* device pointer of the space-to-depth frame staging buffer, bf16 [B*T][21][21][16*C] */
void* r2d2_net_s2d_buffer(r2d2_net* n) { return n ? (void*)n->s2d : nullptr; }
/* Two staging buffers: r2d2_net_s2d_buffer_at(n, idx) is the address to gather a batch into (the second buffer is allocated
 * on first use -- not during a stream capture), r2d2_net_select_s2d(n, idx) makes it the one the following forward / backward
 * calls read.  A learner gathers batch i+1 into the idle buffer while update i runs (worker.py:309-316 keeps its batches
 * prefetched in a queue the same way). */
static int ensure_s2d(r2d2_net* n, int idx) {
    R2D2_REQUIRE(n && (idx == 0 || idx == 1), "bad staging buffer index");
    if (!n->s2d_buf[idx]) {
        const size_t bytes = ((size_t)n->NF * 441 + 32) * 16 * n->C * sizeof(r2d2::bf16);
        R2D2_CUDA_CHECK(cudaMalloc(&n->s2d_buf[idx], bytes));
        R2D2_CUDA_CHECK(cudaMemset(n->s2d_buf[idx], 0, bytes));
        R2D2_CUDA_CHECK(cudaDeviceSynchronize());      // the memset runs on the legacy stream: it must not land after a gather on another stream
    }
    return R2D2_OK;
}
void* r2d2_net_s2d_buffer_at(r2d2_net* n, int idx) { return (n && ensure_s2d(n, idx) == R2D2_OK) ? (void*)n->s2d_buf[idx] : nullptr; }
int r2d2_net_select_s2d(r2d2_net* n, int idx) {
    int rc = ensure_s2d(n, idx);
    if (rc) return rc;
    n->s2d = n->s2d_buf[idx];
    n->s2d_idx = idx;
    return R2D2_OK;
}
int r2d2_net_ku(const r2d2_net* n) { return n ? n->KU : -1; }

/* re-lay out the caller's flat parameter buffer (reference state_dict layout) for slot `which` */
int r2d2_net_pack(r2d2_net* n, int which, const float* params, void* stream) {
    R2D2_REQUIRE(n && (which == 0 || which == 1) && params, "bad arguments");
    const int64_t work = 512ll * FLAT3 > (int64_t)G4 * n->KU ? 512ll * FLAT3 : (int64_t)G4 * n->KU;
    pack_kernel<<<cdiv(work, 256), 256, 0, as_stream(stream)>>>(params, g_doff[n], n->pk[which], n->W1both, which, n->A, n->C, n->KU);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

}  // extern "C"

// Window-convolution epilogues (winconv.cuh): p is a pixel of the GW x GH INPUT grid; only the OW x OH pixels whose
// window stays inside the frame are stored (row = f*OH*OW + gy*OW + gx of the usual NHWC activation matrix).
static const bool g_window_conv = [] { const char* e = getenv("R2D2_WINDOW_CONV"); return !(e && e[0] == '0'); }();
template <int GW, int GH, int OW, int OH>
struct EpiWinBiasSplit {        // out(split)[row*64 + n] = relu(v + bias[n])
    SplitW out; const float* bias;
    struct Pre {};
    __device__ __forceinline__ void prefetch(long long, int, Pre&) const {}
    __device__ __forceinline__ void store16(long long p, int n, const float (&v)[16], int, const Pre&) const {
        const int r = (int)(p % (GW * GH)), gy = r / GW, gx = r - gy * GW;
        if (gy >= OH || gx >= OW) return;
        const size_t row = (size_t)(p / (GW * GH)) * (OH * OW) + gy * OW + gx;
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = fmaxf(v[i] + __ldg(bias + n + i), 0.f);
        split_store16(out.hi, out.lo, row * 64 + n, o);
    }
};
struct EpiWinConv1Pair {        // 21x21 s2d grid -> 20x20; columns 0-31 online act1, 32-63 target act1
    SplitW out0, out1; const float* bias0; const float* bias1; float scale;
    struct Pre {};
    __device__ __forceinline__ void prefetch(long long, int, Pre&) const {}
    __device__ __forceinline__ void store16(long long p, int n, const float (&v)[16], int, const Pre&) const {
        const int r = (int)(p % 441), gy = r / 21, gx = r - gy * 21;
        if (gy >= 20 || gx >= 20) return;
        const size_t o1 = ((size_t)(p / 441) * 100 + (gy >> 1) * 10 + (gx >> 1)) * 128 + ((gy & 1) * 2 + (gx & 1)) * 32;   // act1: s2d-by-2
        const SplitW& o = n < 32 ? out0 : out1;
        const float* b = n < 32 ? bias0 : bias1;
        const int c = n & 31;
        float q[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) q[i] = fmaxf(v[i] * scale + __ldg(b + c + i), 0.f);
        split_store16(o.hi, o.lo, o1 + c, q);
    }
};

// conv3 data gradient: all 81 pixels of the 9x9 grid are real act2 pixels; ReLU mask from act2 (same rows), result on
// conv2's 10x10 gradient grid (row f*100 + y*10 + x; gy == 9 / gx == 9 stay zero)
struct EpiWinDgrad3 {
    SplitW out; SplitC act;
    struct Pre { uint4 m[8]; };                                             // masks of the row's 64 columns (EW = 4: col0 = 0)
    __device__ __forceinline__ void prefetch(long long p, int col0, Pre& pre) const {
#pragma unroll
        for (int i = 0; i < 8; ++i) pre.m[i] = __ldg(reinterpret_cast<const uint4*>(act.hi + (size_t)p * 64 + col0) + i);
    }
    __device__ __forceinline__ void store16(long long p, int n, const float (&v)[16], int col0, const Pre& pre) const {
        const int r = (int)(p % 81), y = r / 9, x = r - y * 9;
        const size_t o = ((size_t)(p / 81) * 100 + y * 10 + x) * 64 + n;
        float q[16];
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            const uint4 h = pre.m[(n - col0 + j) >> 3];
            const uint32_t hw4[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                q[j + 2 * i] = (hw4[i] & 0x7FFFu) ? v[j + 2 * i] : 0.f;
                q[j + 2 * i + 1] = (hw4[i] & 0x7FFF0000u) ? v[j + 2 * i + 1] : 0.f;
            }
        }
        split_store16(out.hi, out.lo, o, q);
    }
};
// conv2 data gradient: row p = (f, Y, X) of the 10x10 s2d-by-2 grid, column n = (ry*2+rx)*32 + c = act1 pixel
// (2Y+ry, 2X+rx); ReLU mask from act1 (same layout); result on conv1's 21x21 gradient grid, 32 channels per pixel
struct EpiWinDgrad2 {
    SplitW out; SplitC act;
    struct Pre { uint4 m[16]; };                                            // masks of the row's 128 columns (EW = 4: col0 = 0)
    __device__ __forceinline__ void prefetch(long long p, int col0, Pre& pre) const {
#pragma unroll
        for (int i = 0; i < 16; ++i) pre.m[i] = __ldg(reinterpret_cast<const uint4*>(act.hi + (size_t)p * 128 + col0) + i);
    }
    __device__ __forceinline__ void store16(long long p, int n, const float (&v)[16], int col0, const Pre& pre) const {
        const int r = (int)(p % 100), Y = r / 10, X = r - Y * 10, sub = n >> 5, c = n & 31;
        const size_t o = ((size_t)(p / 100) * 441 + (2 * Y + (sub >> 1)) * 21 + 2 * X + (sub & 1)) * 32 + c;
        float q[16];
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            const uint4 h = pre.m[(n - col0 + j) >> 3];
            const uint32_t hw4[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                q[j + 2 * i] = (hw4[i] & 0x7FFFu) ? v[j + 2 * i] : 0.f;
                q[j + 2 * i + 1] = (hw4[i] & 0x7FFF0000u) ? v[j + 2 * i + 1] : 0.f;
            }
        }
        split_store16(out.hi, out.lo, o, q);
    }
};
// window weight gradient + split reduction into the reference layout
// (+ the layer's bias gradient = column sums of G, from the same kernel)
template <int GW, int IC, int KH, int KW, int TG, int NO, bool X_HAS_LO, int KP>
static cudaError_t winwgrad(SplitC X, SplitC G, long long R, int chunk, int kind, int64_t bias_off, r2d2_net* net, float scale, cudaStream_t s) {
    constexpr int M = KH * KW * IC;
    const int splits = (int)((R + chunk - 1) / chunk);
    float* ws = ws_take(net, (size_t)splits * M * NO);
    float* bws = colws_take(net, (size_t)splits * NO);
    if (!ws || !bws || chunk % KP) return cudaErrorInvalidValue;
    cudaError_t e = launch_winwgrad<GW, IC, KH, KW, TG, NO, X_HAS_LO, KP, true>(X, G, R, chunk, ws, bws, s);
    if (e != cudaSuccess) return e;
    e = fin_add(net, ws, splits, M, NO, kind, scale);
    return e != cudaSuccess ? e : fin_add(net, bws, splits, 1, NO, kFinBias + B_PLAIN, 1.f, bias_off, 0);
}

struct FwdArgs {
    const float* params; const uint8_t* obs; const uint8_t* last_action; const float* last_reward; const float* hidden;
};

// conv1 of BOTH slots in one launch: columns 0-31 -> online act1, 32-63 -> target act1 (same frames, stacked weights)
struct Epi2Conv1Pair {
    SplitW out0, out1; const float* bias0; const float* bias1; int M; float scale;
    __device__ __forceinline__ void store16(int m, int n, const float (&v)[16], int) const {
        if (m >= M || n >= 64) return;
        const SplitW& o = n < 32 ? out0 : out1;
        const float* b = n < 32 ? bias0 : bias1;
        const int c = n & 31;
        const int f = m / 400, pp = m - f * 400, gy = pp / 20, gx = pp - gy * 20;
        const size_t o1 = ((size_t)f * 100 + (gy >> 1) * 10 + (gx >> 1)) * 128 + ((gy & 1) * 2 + (gx & 1)) * 32;      // act1: s2d-by-2
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = fmaxf(v[j + i] * scale + __ldg(b + c + j + i), 0.f);
            split_store8(o.hi, o.lo, o1 + c + j, r);
        }
    }
};
template <int CH>
static cudaError_t conv1_forward(r2d2_net* n, int which, const float* params, cudaStream_t s) {
    SrcConvK<21, 21, 16 * CH, 20, 20, 2, 2, 1, false> a{n->s2d, nullptr, n->NF};
    const Packed& pk = n->pk[which];
    SrcMatK b{pk.W1s.hi, pk.W1s.lo, 32, 64 * CH, 64 * CH};
    Epi2Conv1Pair e{n->ac[which].act1, n->ac[which].act1, params + n->off[P_C1B], params + n->off[P_C1B], n->NF * 400, 1.f / 255.f};
    return launch_umma2<32, LO_WEIGHT_B>(a, b, e, n->NF * 400, 32, 64 * CH, 1, s);
}
template <int CH>
static cudaError_t conv1_forward_pair(r2d2_net* n, const float* p0, const float* p1, cudaStream_t s) {
    if constexpr (CH == 4) {
        if (g_window_conv) {
            EpiWinConv1Pair we{n->ac[0].act1, n->ac[1].act1, p0 + n->off[P_C1B], p1 + n->off[P_C1B], 1.f / 255.f};
            return launch_winconv<21, 64, 2, 2, 64, false, false, 8>(SplitC{n->s2d, nullptr}, (long long)n->NF * 441, SplitC{n->W1both.hi, n->W1both.lo}, we, s);
        }
    }
    SrcConvK<21, 21, 16 * CH, 20, 20, 2, 2, 1, false> a{n->s2d, nullptr, n->NF};
    SrcMatK b{n->W1both.hi, n->W1both.lo, 64, 64 * CH, 64 * CH};
    Epi2Conv1Pair e{n->ac[0].act1, n->ac[1].act1, p0 + n->off[P_C1B], p1 + n->off[P_C1B], n->NF * 400, 1.f / 255.f};
    return launch_umma2<64, LO_WEIGHT_B>(a, b, e, n->NF * 400, 64, 64 * CH, 1, s);
}
template <int CH>
static cudaError_t conv1_wgrad(r2d2_net* n, cudaStream_t s) {
    const long long NP = (long long)n->NF * 441;               // dpre1g lives on the 21x21 s2d grid (junk row/column are zero)
    if constexpr (CH == 4) {
        return winwgrad<21, 64, 2, 2, 4, 32, false, 128>(SplitC{n->s2d, nullptr}, ro(n->dpre1g), NP, 4096, R_C1W, n->off[P_C1B], n, 1.f / 255.f, s);
    } else {
        SrcMatMN a{n->dpre1g.hi, n->dpre1g.lo, 32, (int)NP, 32};
        SrcConvMN<21, 21, 16 * CH, 21, 21, 2, 2, 1, false> b{n->s2d, nullptr, n->NF};       // junk pixels read the slack rows: finite x 0
        const int splits = (int)((NP + 4095) / 4096);
        cudaError_t e = wgrad2<64, LO_NO_WEIGHT>(a, b, 32, 64 * CH, (int)NP, splits, R_C1, n, 1.f / 255.f, s);
        if (e != cudaSuccess) return e;
        return colsum_split(ro(n->dpre1g), (int)NP, 32, B_PLAIN, n->off[P_C1B], 0, n, s);
    }
}

// frames -> space-to-depth bf16 (once per batch, shared by both slots) + row maps + h0 split
static int net_prep(r2d2_net* n, const uint8_t* obs, const float* hidden, const uint8_t* burn, const uint8_t* learn,
                    const uint8_t* fwd, cudaStream_t s) {
    prep_rows_kernel<<<1 + 32, 256, (n->B + 1) * sizeof(int), s>>>(burn, learn, fwd, hidden, n->ac[0].HsX, n->ac[1].HsX, n->B, n->T, n->F, n->Rmax, n->row_src,
                                                             n->len_full, n->len_learn, n->d_rows);
    if (obs) {                                             // obs == NULL: the frames were staged by r2d2_replay_gather_s2d
        const int64_t total = (int64_t)n->NF * n->C * 441;
        s2d_kernel<<<cdiv(total, 512), 256, 0, s>>>(obs, n->s2d, n->C, total);
    }
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

// encoder + input projection of one slot (model.py:39-49,92); 1/255 of worker.py:342 folded into the conv1 epilogue
static int net_encode(r2d2_net* n, int which, const FwdArgs& fa, cudaStream_t s, bool conv1_done = false) {
    const int B = n->B, T = n->T, A = n->A, KU = n->KU, NF = n->NF;
    Packed& pk = n->pk[which];
    Acts& ac = n->ac[which];
    const int64_t* off = n->off;
    const float* params = fa.params;
    if (!conv1_done || which == 0) {     // pair forward: slot 0's call fills the side columns of both slots (same batch)
        side_columns_kernel<<<T * B, 32, 0, s>>>(ac.U, conv1_done ? n->ac[1].U : SplitW{nullptr, nullptr}, fa.last_action, fa.last_reward, B, T, A, KU);
        R2D2_LAUNCH_CHECK();
    }
    if (!conv1_done) R2D2_CUDA_CHECK(n->C == 1 ? conv1_forward<1>(n, which, params, s) : conv1_forward<4>(n, which, params, s));
    if (g_window_conv) {   // conv2 = 2x2 stride-1 window conv over the 10x10 s2d-by-2 grid of act1 (128 channels)
        EpiWinBiasSplit<10, 10, 9, 9> e{ac.act2, params + off[P_C2B]};
        R2D2_CUDA_CHECK((launch_winconv<10, 128, 2, 2, 64, true>(SplitC{ac.act1.hi, ac.act1.lo}, (long long)NF * 100, SplitC{pk.W2p.hi, pk.W2p.lo}, e, s)));
    } else {
        SrcConvK<10, 10, 128, 9, 9, 2, 2, 1> a{ac.act1.hi, ac.act1.lo, NF};
        SrcMatK b{pk.W2p.hi, pk.W2p.lo, 64, 512, 512};
        Epi2BiasSplit<true> e{ac.act2, params + off[P_C2B], NF * 81, 64, 64, 1.f};
        R2D2_CUDA_CHECK((launch_umma2<64, LO_WEIGHT_B>(a, b, e, NF * 81, 64, 512, 1, s)));
    }
    if (g_window_conv) {
        EpiWinBiasSplit<9, 9, 7, 7> e{ac.act3, params + off[P_C3B]};
        R2D2_CUDA_CHECK((launch_winconv<9, 64, 3, 3, 64, true>(SplitC{ac.act2.hi, ac.act2.lo}, (long long)NF * 81, SplitC{pk.W3p.hi, pk.W3p.lo}, e, s)));
    } else {
        SrcConvK<9, 9, 64, 7, 7, 3, 3, 1> a{ac.act2.hi, ac.act2.lo, NF};
        SrcMatK b{pk.W3p.hi, pk.W3p.lo, 64, 576, 576};
        Epi2BiasSplit<true> e{ac.act3, params + off[P_C3B], NF * 49, 64, 64, 1.f};
        R2D2_CUDA_CHECK((launch_umma2<64, LO_WEIGHT_B>(a, b, e, NF * 49, 64, 576, 1, s)));
    }
    {
        SrcMatK a{ac.act3.hi, ac.act3.lo, NF, FLAT3, FLAT3};
        SrcMatK b{pk.Wfcp.hi, pk.Wfcp.lo, LATENT, FLAT3, FLAT3};
        Epi2Latent e{ac.U, params + off[P_FCB], B, T, KU};
        if (g_pair_gemm)
            R2D2_CUDA_CHECK((launch_umma3<false, false, LO_WEIGHT_B>(Mat3{ac.act3.hi, ac.act3.lo, NF, FLAT3, FLAT3}, Mat3{pk.Wfcp.hi, pk.Wfcp.lo, LATENT, FLAT3, FLAT3},
                                                                    e, NF, LATENT, FLAT3, 1, s)));
        else
        R2D2_CUDA_CHECK((launch_umma2<128, LO_WEIGHT_B>(a, b, e, NF, LATENT, FLAT3, 1, s)));
    }
    {   // LSTM input projection for all steps at once (hoisted out of the recurrence)
        SrcMatK a{ac.U.hi, ac.U.lo, T * B, KU, KU};
        SrcMatK b{pk.Wih_p.hi, pk.Wih_p.lo, G4, KU, KU};
        Epi2F32 e{ac.XP, pk.bias_p, T * B, G4, G4, 1.f};
        if (g_pair_gemm)
            R2D2_CUDA_CHECK((launch_umma3<false, false>(Mat3{ac.U.hi, ac.U.lo, T * B, KU, KU}, Mat3{pk.Wih_p.hi, pk.Wih_p.lo, G4, KU, KU}, e, T * B, G4, KU, 1, s)));
        else
        R2D2_CUDA_CHECK((launch_umma2<128>(a, b, e, T * B, G4, KU, 1, s)));
    }
    return R2D2_OK;
}

struct StepOps { SrcMatK a; SrcMatK b; Epi2LstmCell e; };
static StepOps lstm_step_ops(r2d2_net* n, int which, const float* hidden, int t) {
    const int B = n->B;
    Packed& pk = n->pk[which];
    Acts& ac = n->ac[which];
    const size_t prev = (size_t)(t - 1) * B * H, cur = (size_t)t * B * H;
    const SplitC hp{ac.HsX.hi + cur, ac.HsX.lo + cur};           // state before step t (block 0 = h0)
    const float* cp = t ? ac.Cs + prev : hidden + H;
    return StepOps{SrcMatK{hp.hi, hp.lo, B, H, H}, SrcMatK{pk.Whh_p.hi, pk.Whh_p.lo, G4, H, H},
                   Epi2LstmCell{ac.XP + (size_t)t * B * G4, cp, t ? H : 2 * H, hp, SplitW{ac.Hs.hi + cur, ac.Hs.lo + cur}, ac.Cs + cur,
                                which == 0 ? ac.Gs + (size_t)t * B * G4 : nullptr, n->len_full, t, B}};
}

// recurrence (model.py:95-100 / 134-141): sequences advance while t < b+l+f.  which = 0/1: one slot; 2: both slots
// in the same launches (two independent recurrences hide each other's per-step latency).
static int net_recurrence(r2d2_net* n, int which, const float* hidden, cudaStream_t s) {
    const int B = n->B, T = n->T;
    if (g_persistent_recurrence && (B <= 64 || g_cluster_recurrence)) {   // one launch for all T steps
        RecFwdParams P;
        for (int k = 0; k < 2; ++k) {
            P.Whi[k] = n->pk[k].Whh_p.hi; P.Wlo[k] = n->pk[k].Whh_p.lo; P.XP[k] = n->ac[k].XP;
            P.Hhi[k] = n->ac[k].HsX.hi; P.Hlo[k] = n->ac[k].HsX.lo; P.Cs[k] = n->ac[k].Cs;
            P.Gs[k] = k == 0 ? n->ac[k].Gs : nullptr;
        }
        P.c0 = hidden + H; P.ld_c0 = 2 * H; P.len = n->len_full; P.bar = n->rec_bar; P.B = B; P.T = T;
        P.net_base = which == 2 ? 0 : which;
        P.fast = g_fast_math == 1;
        P.trace = g_rec_trace;
        cudaError_t e = g_cluster_recurrence ? launch_rec2_fwd(P, which == 2 ? 2 : 1, s) : cudaErrorNotSupported;
        if (e == cudaErrorNotSupported && B <= 64) e = launch_rec_fwd(P, which == 2 ? 2 : 1, s);   // device cannot host 16-CTA clusters
        if (e != cudaErrorNotSupported) {
            R2D2_CUDA_CHECK(e);
            return R2D2_OK;
        }
    }
    for (int t = 0; t < T; ++t) {
        if (which == 2) {
            StepOps o0 = lstm_step_ops(n, 0, hidden, t), o1 = lstm_step_ops(n, 1, hidden, t);
            Pair<SrcMatK> pa{o0.a, o1.a}, pb{o0.b, o1.b};
            Pair<Epi2LstmCell> pe{o0.e, o1.e};
            R2D2_CUDA_CHECK((launch_umma2<64>(pa, pb, pe, B, G4, H, 2, s)));
        } else {
            StepOps o = lstm_step_ops(n, which, hidden, t);
            R2D2_CUDA_CHECK((launch_umma2<64>(o.a, o.b, o.e, B, G4, H, 1, s)));
        }
    }
    return R2D2_OK;
}

// dueling head on the gathered rows (model.py:102-117, 143-148)
static int net_heads(r2d2_net* n, int which, const float* params, float* q_learn_out, float* q_shift_out, cudaStream_t s, bool launch_out = true) {
    const int A = n->A, Rmax = n->Rmax;
    Packed& pk = n->pk[which];
    Acts& ac = n->ac[which];
    const int64_t* off = n->off;
    {   // rows [0, Rmax) = learning positions (q), [Rmax, 2 Rmax) = n-step-shifted positions; the target network needs only the latter
        const int r0 = q_learn_out ? 0 : Rmax, nr = 2 * Rmax - r0;
        SrcRowGatherK a{ac.Hs.hi, ac.Hs.lo, n->row_src + r0, nr, H, H};
        SrcMatK b{pk.Wh0.hi, pk.Wh0.lo, 2 * H, H, H};
        Epi2BiasSplit<true> e{SplitW{ac.hid.hi + (size_t)r0 * 2 * H, ac.hid.lo + (size_t)r0 * 2 * H}, pk.bh0, nr, 2 * H, 2 * H, 1.f};
        R2D2_CUDA_CHECK((launch_umma2<128>(a, b, e, nr, 2 * H, H, 1, s)));
    }
    if (!launch_out) return R2D2_OK;                // the caller issues one head_out launch for several networks (net_head_out)
    HeadJobs jobs;
    int nj = 0;
    const HeadJob base{ro(ac.hid), 0, params + off[P_A2W], params + off[P_A2B], params + off[P_V2W], params + off[P_V2B], nullptr};
    if (q_learn_out) { jobs.job[nj] = base; jobs.job[nj].q_out = q_learn_out; ++nj; }
    if (q_shift_out) { jobs.job[nj] = base; jobs.job[nj].row_offset = (size_t)Rmax; jobs.job[nj].q_out = q_shift_out; ++nj; }
    if (nj) head_out_kernel<<<dim3(cdiv((int64_t)Rmax * 32, 256), nj), 256, 0, s>>>(jobs, Rmax, A);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

extern "C" {

/* Forward unroll of slot `which` (0 online, 1 target).  model.py:81-150.  See include/r2d2_b200.h. */
int r2d2_net_forward(r2d2_net* n, int which, const float* params, const uint8_t* obs, const uint8_t* last_action,
                     const float* last_reward, const float* hidden, const uint8_t* burn, const uint8_t* learn,
                     const uint8_t* fwd, float* q_learn_out, float* q_shift_out, void* stream) {
    R2D2_REQUIRE(n && (which == 0 || which == 1) && params && obs && last_action && last_reward && hidden && burn && learn && fwd,
                 "bad arguments");
    cudaStream_t s = as_stream(stream);
    if (which == 0) n->hidden = hidden;
    FwdArgs fa{params, obs, last_action, last_reward, hidden};
    int rc = net_prep(n, obs, hidden, burn, learn, fwd, s);
    if (!rc) rc = net_encode(n, which, fa, s);
    if (!rc) rc = net_recurrence(n, which, hidden, s);
    if (!rc) rc = net_heads(n, which, params, q_learn_out, q_shift_out, s);
    return rc;
}

/* cudaEvent_t (or NULL) that r2d2_net_backward records on its stream as soon as the gradients of feature.7.weight and of
 * every later tensor of the flat layout are final (before the conv layers' backward).  Multi-GPU learners wait on it from
 * a side stream to overlap the all-reduce of that range with the rest of the backward pass. */
int r2d2_net_set_dense_grads_event(r2d2_net* n, void* cuda_event) {
    R2D2_REQUIRE(n, "null handle");
    n->dense_grads_event = cuda_event;
    return R2D2_OK;
}
/* r2d2_net_shadow_gate: a one-warp kernel on `stream` that returns once the BPTT recurrence of the update it belongs to is
 * executing (or after ~1 s: placement only, correctness never depends on it).  That kernel holds 4 clusters of 16 SMs for ~350 us and leaves the other 84 SMs idle; work
 * enqueued behind the gate (the gather of the next batch, launched with a shared-memory footprint that does not fit next to a
 * recurrence CTA) runs on exactly those SMs.  Launching it BEFORE the recurrence instead would let its CTAs take SMs in every
 * GPC, and a 16-CTA cluster needs a whole GPC's worth of free SMs to start.
 * Pairing: every update launches one BPTT kernel (it counts itself in ctl[0] when it begins) and the caller one gate;
 * r2d2_net_shadow_gate_reset (stream-ordered, before the first update of the pipeline) notes how many BPTT kernels have run
 * so far, gate number g then waits for BPTT number base + g.  A gate that runs late (its recurrence already over) passes. */
__global__ void r2d2_shadow_gate_kernel(volatile unsigned int* ctl /* [0] BPTT kernels started, [1] base, [2] gates so far */) {
    if (threadIdx.x != 0) return;
    const unsigned int g = ctl[2] + 1u;
    const unsigned int target = ctl[1] + g;
    const long long t0 = clock64();
    while ((int)(ctl[0] - target) < 0 && clock64() - t0 < 2000000000ll) __nanosleep(500);
    ctl[2] = g;
    __nanosleep(3000);                                       // the other clusters of the launch are placed within this
}
__global__ void r2d2_shadow_gate_reset_kernel(volatile unsigned int* ctl) {
    if (threadIdx.x == 0) { ctl[1] = ctl[0]; ctl[2] = 0u; }
}
static int shadow_gate_configure() {
    static unsigned long long configured = 0;
    int dev = 0;
    R2D2_CUDA_CHECK(cudaGetDevice(&dev));
    if (!(configured & (1ull << (dev & 63)))) {              // a shared-memory-free CTA must not flip its SM to the large-L1 carveout (dp.cu)
        R2D2_CUDA_CHECK(cudaFuncSetAttribute(r2d2_shadow_gate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        R2D2_CUDA_CHECK(cudaFuncSetAttribute(r2d2_shadow_gate_reset_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured |= 1ull << (dev & 63);
    }
    return R2D2_OK;
}
int r2d2_net_shadow_gate(r2d2_net* n, void* stream) {
    R2D2_REQUIRE(n, "null handle");
    int rc = shadow_gate_configure();
    if (rc) return rc;
    r2d2_shadow_gate_kernel<<<1, 32, 0, as_stream(stream)>>>(n->rec_bar + 60);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}
int r2d2_net_shadow_gate_reset(r2d2_net* n, void* stream) {
    R2D2_REQUIRE(n, "null handle");
    int rc = shadow_gate_configure();
    if (rc) return rc;
    r2d2_shadow_gate_reset_kernel<<<1, 32, 0, as_stream(stream)>>>(n->rec_bar + 60);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}
// inside a stream capture a plain record would only mark a capture-internal point; callers wait on these events from a
// stream OUTSIDE the graph, so the record must become an external event-record node
static cudaError_t record_user_event(void* ev, cudaStream_t s) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaError_t e = cudaStreamIsCapturing(s, &cap);
    if (e != cudaSuccess) return e;
    return cudaEventRecordWithFlags((cudaEvent_t)ev, s, cap == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
}
/* cudaEventRecord for events that streams OUTSIDE a CUDA graph wait on: an external event-record node while `stream` is being
 * captured, a plain record otherwise.  (The learner marks "K2 done: priorities final" this way for its sampling stream.) */
int r2d2_event_record(void* cuda_event, void* stream) {
    R2D2_REQUIRE(cuda_event, "null event");
    R2D2_CUDA_CHECK(record_user_event(cuda_event, as_stream(stream)));
    return R2D2_OK;
}

/* (h, c) of slot `which` after time step t of the last forward, as [B][2][512] -- what an actor carries to its next
 * step (model.py:65-79 returns it; worker.py:533-541).  With T = 1 nets this turns r2d2_net_forward into a batched
 * single-step actor inference: hidden_out of one call is the `hidden` argument of the next. */
int r2d2_net_state_after(r2d2_net* n, int which, int t, float* hidden_out, void* stream) {
    R2D2_REQUIRE(n && (which == 0 || which == 1) && t >= 0 && t < n->T && hidden_out, "bad arguments");
    const Acts& ac = n->ac[which];
    state_after_kernel<<<cdiv((int64_t)n->B * H, 256), 256, 0, as_stream(stream)>>>(ro(ac.Hs), ac.Cs, n->B, t, hidden_out);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

/* The learner's three Q tensors in one call (worker.py:346,347,352): online and target unrolls on the same
 * batch with the two recurrences advanced together. */
int r2d2_net_forward_pair(r2d2_net* n, const float* params_online, const float* params_target, const uint8_t* obs,
                          const uint8_t* last_action, const float* last_reward, const float* hidden, const uint8_t* burn,
                          const uint8_t* learn, const uint8_t* fwd, float* q_learn_out, float* qn_online_out,
                          float* qn_target_out, void* stream) {
    R2D2_REQUIRE(n && params_online && params_target && last_action && last_reward && hidden && burn && learn && fwd &&
                     q_learn_out && qn_online_out && qn_target_out,
                 "bad arguments");
    cudaStream_t s = as_stream(stream);
    n->hidden = hidden;
    FwdArgs f0{params_online, obs, last_action, last_reward, hidden}, f1{params_target, obs, last_action, last_reward, hidden};
    int rc = net_prep(n, obs, hidden, burn, learn, fwd, s);
    if (!rc) {
        cudaError_t e = n->C == 1 ? conv1_forward_pair<1>(n, params_online, params_target, s)
                                  : conv1_forward_pair<4>(n, params_online, params_target, s);
        R2D2_CUDA_CHECK(e);
    }
    if (!rc) rc = net_encode(n, 0, f0, s, true);
    if (!rc) rc = net_encode(n, 1, f1, s, true);
    if (!rc) rc = net_recurrence(n, 2, hidden, s);
    if (!rc) rc = net_heads(n, 0, params_online, q_learn_out, qn_online_out, s, false);
    if (!rc) rc = net_heads(n, 1, params_target, nullptr, qn_target_out, s, false);
    if (!rc) {      // the three Q tensors (worker.py:346,347,352) from one launch of the dueling output layer
        const int64_t* off = n->off;
        const int Rmax = n->Rmax;
        auto mk = [&](int k, const float* p, size_t row0, float* out) {
            return HeadJob{ro(n->ac[k].hid), row0, p + off[P_A2W], p + off[P_A2B], p + off[P_V2W], p + off[P_V2B], out};
        };
        HeadJobs jobs{{mk(0, params_online, 0, q_learn_out), mk(0, params_online, (size_t)Rmax, qn_online_out), mk(1, params_target, (size_t)Rmax, qn_target_out)}};
        head_out_kernel<<<dim3(cdiv((int64_t)Rmax * 32, 256), 3), 256, 0, s>>>(jobs, Rmax, n->A);
        R2D2_LAUNCH_CHECK();
    }
    return rc;
}

/* BPTT + encoder backward of the ONLINE slot (loss.backward(), worker.py:363).
 *   dq [Rmax][A] = d loss / d q_learn (rows >= sum(learn) ignored); grads: flat buffer in the
 *   parameter layout, fully overwritten. */
int r2d2_net_backward(r2d2_net* n, const float* params, const float* dq, float* grads, void* stream) {
    R2D2_REQUIRE(n && params && dq && grads && n->hidden, "bad arguments / forward(online) not run");
    cudaStream_t s = as_stream(stream);
    const int B = n->B, T = n->T, A = n->A, KU = n->KU, NF = n->NF, Rmax = n->Rmax;
    Packed& pk = n->pk[0];
    Acts& ac = n->ac[0];
    const int64_t* off = n->off;
    const int64_t* d_off = g_doff[n];
    const size_t TB = (size_t)T * B;

    fin_begin(n);
    // ---- head
    head_out_bwd_kernel<<<Rmax, 256, 0, s>>>(dq, ro(ac.hid), params + off[P_A2W], params + off[P_V2W], n->d_rows, A, n->dout16,
                                            n->dhid);
    R2D2_LAUNCH_CHECK();
    {   // layer-2 weights: [A+1 (pad 16)] x [1024] = dout16^T . hid   (tiny: CUDA cores)
        const int P = 128, chunk = (Rmax + P - 1) / P;
        float* ws = ws_take(n, (size_t)P * kDW * 2 * H);
        R2D2_REQUIRE(ws, "gradient workspace exhausted");
        head_w2_grad_kernel<<<dim3(2 * H / 128, P), 128, 0, s>>>(n->dout16, ro(ac.hid), Rmax, chunk, ws);
        R2D2_LAUNCH_CHECK();
        R2D2_CUDA_CHECK(fin_add(n, ws, P, kDW, 2 * H, R_H2, 1.f));
        R2D2_CUDA_CHECK(colsum_f32(n->dout16, Rmax, kDW, B_H2, off[P_A2B], off[P_V2B], n, s));
    }
    {   // layer-0 weights: [1024] x [512] = dhid^T . Hsel
        SrcMatMN a{n->dhid.hi, n->dhid.lo, 2 * H, Rmax, 2 * H};
        SrcRowGatherMN b{ac.Hs.hi, ac.Hs.lo, n->row_src, H, Rmax, H};
        R2D2_CUDA_CHECK((wgrad2<128>(a, b, 2 * H, H, Rmax, 4, R_H0, n, 1.f, s)));
        R2D2_CUDA_CHECK(colsum_split(ro(n->dhid), Rmax, 2 * H, B_H0, off[P_A0B], off[P_V0B], n, s));
    }
    R2D2_CUDA_CHECK(cudaMemsetAsync(n->dH, 0, TB * H * sizeof(float), s));
    {   // d hidden rows -> dH[t][b]
        SrcMatK a{n->dhid.hi, n->dhid.lo, Rmax, 2 * H, 2 * H};
        SrcMatMN b{pk.Wh0.hi, pk.Wh0.lo, H, 2 * H, H};
        Epi2ScatterRows e{n->dH, n->row_src, n->d_rows, Rmax};
        R2D2_CUDA_CHECK((launch_umma2<128>(a, b, e, Rmax, H, 2 * H, 1, s)));
    }
    // ---- BPTT through all b+l steps, burn-in included (no detach anywhere in model.py:122-150)
    bool bptt_done = false;
    if (g_persistent_recurrence && (B <= 64 || g_cluster_recurrence)) {   // one launch for all T steps
        RecBwdParams P;
        P.WThi = pk.WhhT_p.hi; P.WTlo = pk.WhhT_p.lo; P.dH = n->dH; P.Gs = ac.Gs; P.Cs = ac.Cs;
        P.c0 = n->hidden + H; P.ld_c0 = 2 * H; P.len = n->len_learn; P.DGhi = n->DG.hi; P.DGlo = n->DG.lo;
        P.partial = n->rec_partial; P.flags = n->rec_bar + 32; P.B = B; P.T = T; P.fast = g_fast_math == 1;
        P.trace = g_rec_trace_bwd;
        P.started = n->rec_bar + 60;
        cudaError_t e = g_cluster_recurrence ? launch_rec2_bwd(P, s) : cudaErrorNotSupported;     // 16-CTA clusters, DSMEM reduce-scatter
        if (e == cudaErrorNotSupported && B <= 64) e = launch_rec_bwd(P, s);                       // L2-flag cooperative kernel
        if (e != cudaErrorNotSupported) {
            R2D2_CUDA_CHECK(e);
            bptt_done = true;
        }
    }
    if (!bptt_done) {
    R2D2_CUDA_CHECK(cudaMemsetAsync(n->dcrec, 0, (size_t)B * H * sizeof(float), s));
    for (int t = T - 1; t >= 0; --t) {
        const float* cprev = t ? ac.Cs + (size_t)(t - 1) * B * H : n->hidden + H;
        lstm_bwd_pointwise_kernel<<<cdiv(B * H, 256), 256, 0, s>>>(n->dH + (size_t)t * B * H, n->dhrec, t == T - 1 ? 0 : kRecSplits,
                                                                  n->dcrec, ac.Gs + (size_t)t * B * G4, ac.Cs + (size_t)t * B * H,
                                                                  cprev, t ? H : 2 * H, n->len_learn, t, B, n->DG, (size_t)t * B * G4);
        if (t > 0) {
            const size_t o = (size_t)t * B * G4;
            SrcMatK a{n->DG.hi + o, n->DG.lo + o, B, G4, G4};
            SrcMatMN b{pk.Whh_p.hi, pk.Whh_p.lo, H, G4, H};
            Epi2Partial e{n->dhrec, B, H};
            R2D2_CUDA_CHECK((launch_umma2<64>(a, b, e, B, H, G4, kRecSplits, s)));
        }
    }
    }
    R2D2_LAUNCH_CHECK();
    {   // recurrent weight gradients over all (t,b) rows
        SrcMatMN a{n->DG.hi, n->DG.lo, G4, (int)TB, G4};
        SrcMatMN bh{ac.HsX.hi, ac.HsX.lo, H, (int)TB, H};        // row (t,b) of HsX is the state BEFORE step t
        SrcMatMN bu{ac.U.hi, ac.U.lo, KU, (int)TB, KU};
        if (g_pair_gemm) {       // 8 x 2 (x 3) pair tiles: split K so that one wave of 74 pairs is (nearly) full, K <= 4096 per partial
            const Mat3 dg{n->DG.hi, n->DG.lo, G4, (int)TB, G4};
            const int sp = std::max(4, cdiv((long long)TB, 4096));
            R2D2_CUDA_CHECK((wgrad3<>(dg, Mat3{ac.HsX.hi, ac.HsX.lo, H, (int)TB, H}, G4, H, (int)TB, sp, R_WHH, n, 1.f, s)));
            R2D2_CUDA_CHECK((wgrad3<>(dg, Mat3{ac.U.hi, ac.U.lo, KU, (int)TB, KU}, G4, KU, (int)TB, std::max(3, cdiv((long long)TB, 4096)), R_WIH, n, 1.f, s)));
        } else {
        R2D2_CUDA_CHECK((wgrad2<128>(a, bh, G4, H, (int)TB, 2, R_WHH, n, 1.f, s)));
        R2D2_CUDA_CHECK((wgrad2<128>(a, bu, G4, KU, (int)TB, 2, R_WIH, n, 1.f, s)));
        }
        R2D2_CUDA_CHECK(colsum_split(ro(n->DG), (int)TB, G4, B_LSTM, off[P_BIH], off[P_BHH], n, s));
    }
    {   // d latent (ReLU-masked), frame-major
        SrcMatK a{n->DG.hi, n->DG.lo, (int)TB, G4, G4};
        SrcMatMN b{pk.Wih_p.hi, pk.Wih_p.lo, LATENT, G4, KU};
        Epi2DLatent e{n->dlat, ro(ac.U), B, T, KU};
        if (g_pair_gemm)
            R2D2_CUDA_CHECK((launch_umma3<false, true>(Mat3{n->DG.hi, n->DG.lo, (int)TB, G4, G4}, Mat3{pk.Wih_p.hi, pk.Wih_p.lo, LATENT, G4, KU}, e, (int)TB, LATENT, G4, 1, s)));
        else
        R2D2_CUDA_CHECK((launch_umma2<128>(a, b, e, (int)TB, LATENT, G4, 1, s)));
    }
    // ---- encoder backward
    {
        SrcMatMN a{n->dlat.hi, n->dlat.lo, LATENT, NF, LATENT};
        SrcMatMN b{ac.act3.hi, ac.act3.lo, FLAT3, NF, FLAT3};
        if (g_pair_gemm)     // 2 x 13 pair tiles x 2 splits: one wave
            R2D2_CUDA_CHECK((wgrad3<LO_NO_WEIGHT>(Mat3{n->dlat.hi, n->dlat.lo, LATENT, NF, LATENT}, Mat3{ac.act3.hi, ac.act3.lo, FLAT3, NF, FLAT3}, LATENT, FLAT3, NF,
                                                  std::max(2, cdiv(NF, 4096)), R_FC, n, 1.f, s)));
        else
        R2D2_CUDA_CHECK((wgrad2<128, LO_NO_WEIGHT>(a, b, LATENT, FLAT3, NF, 2, R_FC, n, 1.f, s)));
        R2D2_CUDA_CHECK(colsum_split(ro(n->dlat), NF, LATENT, B_PLAIN, off[P_FCB], 0, n, s));
        R2D2_CUDA_CHECK(fin_flush(n, grads, d_off, s));            // every dense-layer gradient (heads, LSTM, FC) in one reduction launch
        // every gradient from feature.7.weight to the end of the flat layout (FC, LSTM, heads: 98 % of the bytes) is final:
        // a data-parallel caller can start reducing that range while the conv layers' backward still runs
        if (n->dense_grads_event) R2D2_CUDA_CHECK(record_user_event(n->dense_grads_event, s));
        SrcMatK a2{n->dlat.hi, n->dlat.lo, NF, LATENT, LATENT};
        SrcMatMN b2{pk.Wfcp.hi, pk.Wfcp.lo, FLAT3, LATENT, FLAT3};
        Epi2MaskedToGrid3 e{n->dpre3, ro(ac.act3), NF};
        if (g_pair_gemm)
            R2D2_CUDA_CHECK((launch_umma3<false, true, LO_WEIGHT_B>(Mat3{n->dlat.hi, n->dlat.lo, NF, LATENT, LATENT}, Mat3{pk.Wfcp.hi, pk.Wfcp.lo, FLAT3, LATENT, FLAT3}, e,
                                                                   NF, FLAT3, LATENT, 1, s)));
        else
        R2D2_CUDA_CHECK((launch_umma2<128, LO_WEIGHT_B>(a2, b2, e, NF, FLAT3, LATENT, 1, s)));
    }
    // The conv layers run as window convolutions (winconv.cuh): gradients live on each layer's input grid.
    {   // conv3: weights from act2 (9x9 grid) x dpre3 (same grid); data gradient = 3x3 window conv of dpre3 with flipped taps
        const long long R3 = (long long)NF * 81;
        R2D2_CUDA_CHECK((winwgrad<9, 64, 3, 3, 9, 64, true, 64>(ro(ac.act2), ro(n->dpre3), R3, 3072, R_C3W, off[P_C3B], n, 1.f, s)));
        EpiWinDgrad3 e{n->dpre2, ro(ac.act2)};
        R2D2_CUDA_CHECK((launch_winconv<9, 64, 3, 3, 64, true, true>(ro(n->dpre3), R3, SplitC{pk.W3d.hi, pk.W3d.lo}, e, s)));
    }
    {   // conv2 on the 10x10 s2d-by-2 grid of act1
        const long long R2 = (long long)NF * 100;
        R2D2_CUDA_CHECK((winwgrad<10, 128, 2, 2, 4, 64, true, 64>(ro(ac.act1), ro(n->dpre2), R2, 3712, R_C2W, off[P_C2B], n, 1.f, s)));
        EpiWinDgrad2 e{n->dpre1g, ro(ac.act1)};
        R2D2_CUDA_CHECK((launch_winconv<10, 64, 2, 2, 128, true, true>(ro(n->dpre2), R2, SplitC{pk.W2q.hi, pk.W2q.lo}, e, s)));
    }
    {   // conv1 (weights only; frames need no gradient)
        R2D2_CUDA_CHECK(n->C == 1 ? conv1_wgrad<1>(n, s) : conv1_wgrad<4>(n, s));
    }
    R2D2_CUDA_CHECK(fin_flush(n, grads, d_off, s));                // conv weights and biases
    return R2D2_OK;
}

/* Incremented by every r2d2_set_* knob: lets a caller that caches launch sequences (CUDA graphs) notice a mode change. */
int r2d2_config_epoch(void) { return g_config_epoch; }

/* diagnostics: number of 16-CTA recurrence clusters the current device keeps resident at once (< 0: query failed) */
int r2d2_debug_cluster_capacity(void) { return rec2_max_active_clusters<1>() * 100 + rec2_max_active_clusters<2>(); }

/* 1 (default): forward LSTM recurrence inside 16-CTA clusters (distributed-shared-memory exchange of h); 0: L2-flag kernel. */
int r2d2_set_cluster_recurrence(int on) {
    ++g_config_epoch;
    int prev = g_cluster_recurrence;
    g_cluster_recurrence = on ? 1 : 0;
    return prev;
}

/* 1 (default): plain-matrix GEMMs of K1/K1b on CTA pairs (cta_group::2, TMA); 0: single-CTA cp.async kernels.  Returns the previous value. */
int r2d2_set_pair_gemm(int on) {
    ++g_config_epoch;
    int prev = g_pair_gemm;
    g_pair_gemm = on ? 1 : 0;
    return prev;
}

/* 1 (default): the T-step recurrence runs as one persistent cooperative kernel when B <= 64; 0: per-step launches. */
int r2d2_set_persistent_recurrence(int on) {
    ++g_config_epoch;
    int prev = g_persistent_recurrence;
    g_persistent_recurrence = on ? 1 : 0;
    return prev;
}

/* debug: attach a device buffer of T*8 uint64 receiving globaltimer stamps of CTA 0 of the persistent recurrence */
int r2d2_debug_rec_trace(void* device_buffer) {
    ++g_config_epoch;
    g_rec_trace = (unsigned long long*)device_buffer;
    return R2D2_OK;
}

/* debug: per-step globaltimer stamps of CTA 0 of the cluster BPTT kernel (T*8 uint64) */
int r2d2_debug_rec_trace_bwd(void* device_buffer) {
    ++g_config_epoch;
    g_rec_trace_bwd = (unsigned long long*)device_buffer;
    return R2D2_OK;
}

/* test/debug access to intermediates (bf16 planes are named "<tensor>.hi" / "<tensor>.lo") */
void* r2d2_net_debug_ptr(r2d2_net* n, int which, const char* name) {
    if (!n || !name) return nullptr;
    Acts& a = n->ac[which & 1];
    struct { const char* k; void* v; } tab[] = {
        {"U.hi", a.U.hi}, {"U.lo", a.U.lo}, {"Hs.hi", a.Hs.hi}, {"Hs.lo", a.Hs.lo}, {"act1.hi", a.act1.hi}, {"act1.lo", a.act1.lo},
        {"act3.hi", a.act3.hi}, {"act3.lo", a.act3.lo}, {"XP", a.XP}, {"Cs", a.Cs}, {"Gs", a.Gs}, {"dH", n->dH},
        {"DG.hi", n->DG.hi}, {"DG.lo", n->DG.lo}, {"row_src", n->row_src}, {"rows", n->d_rows}, {"s2d", n->s2d}};
    for (auto& e : tab)
        if (!strcmp(e.k, name)) return e.v;
    return nullptr;
}

}  // extern "C"
