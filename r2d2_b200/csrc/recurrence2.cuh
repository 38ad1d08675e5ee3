// Cluster-resident LSTM recurrence: the hidden state never leaves the SMs between time steps.
//
// recurrence.cuh exchanges h_t between its 64 CTAs per network through L2: publish -> fence -> flag -> poll -> reload,
// four L2 traversals on the critical path of every step (5.9 us per step measured, tensor pipe 7-14 % busy).  Here one
// thread-block CLUSTER of 16 CTAs owns a whole recurrence for 16 sequences:
//
//   grid = (networks x batch quarters) clusters of 16 CTAs.  CTA c of a cluster owns hidden units [32c, 32c+32), i.e. the
//   128 gate-interleaved rows [128c, 128c+128) of W_hh, for the 16 sequences of its quarter.
//   * W_hh slice resident for the whole kernel: the bf16 hi plane in TENSOR MEMORY (A operand of tcgen05.mma read from
//     TMEM: 128 lanes x 256 columns), the lo plane in shared memory (128 KB, K-major SWIZZLE_128B).
//   * per step  gates[128 x 16] = W_slice[128 x 512] . h_{t-1}[16 x 512]^T  as M=128, N=16 MMAs (bf16x3: W_hi.h_hi +
//     W_hi.h_lo + W_lo.h_hi), consumed source-CTA by source-CTA as the pieces of h_{t-1} arrive;
//   * h_{t-1} lives in shared memory as 16 tiles (one per source CTA) of [16 sequences][32 units] bf16 hi|lo, K-major
//     SWIZZLE_64B.  After the cell update a CTA writes its own tile once into a staging buffer and pushes it to all 16
//     CTAs of the cluster with 16 bulk copies over distributed shared memory (cp.async.bulk shared::cta ->
//     shared::cluster); each copy completes on an mbarrier of the DESTINATION CTA, which is what its MMA warp waits on.
//     No global-memory round trip, no flags, no fences: one DSMEM hop (~0.1 us) per step.
//   * epilogue: thread = TMEM lane = gate row (4 unit + gate); a 4-lane shuffle transpose gives every thread the four
//     gates of (unit, 4 sequences); c and h stay in registers across steps; h, c and the gates are also streamed to
//     global memory for the heads and the backward pass (off the critical path).
// Buffers alternate with the step parity; a CTA can run at most one step ahead of the slowest CTA of its cluster (it
// needs everybody's h_t to produce h_{t+1}), which is what makes two buffers and two staging tiles sufficient.
#pragma once
#include <algorithm>

#include "recurrence.cuh"
#include "umma3.cuh"

namespace r2d2 {

constexpr int R2_CL = 16;                          // CTAs per cluster (non-portable size)
constexpr int R2_EW = 16;                          // epilogue warps: 4 TMEM lane quadrants x 4 sequence groups
constexpr int R2_THREADS = 32 * R2_EW + 32;        // + 1 MMA-issue warp
constexpr int R2_GRP = 4;                          // source CTAs per arrival barrier (one mbarrier wait costs ~90 clk even when complete)

// A cluster runs NSTR independent recurrences ("streams") of 16 sequences each, sharing the resident W_hh slice.
// NSTR = 1: 16 sequences per cluster.  A B200 keeps only 7 such clusters resident, and 64 sequences x 2 networks are 8,
// so the default for full batches is NSTR = 2: 32 sequences per cluster as two streams whose steps interleave -- while
// the cell warps work on stream 0, the MMA warp already multiplies stream 1 -- which hides most of the per-step latency
// chain (MMA -> cell -> DSMEM exchange) of one stream behind the other.  Its h buffers are twice as large, so half of the
// W lo plane (k < 256) moves to tensor memory next to the hi plane.
template <int NSTR> struct Rec2Cfg {
    static constexpr int NS = 16;                              // sequences per stream
    static constexpr int kPlane = NS * 64;                     // one bf16 plane of one source tile: [16][32 units]
    static constexpr int kTile = 2 * kPlane;                   // hi | lo
    static constexpr int kKT = NSTR == 2 ? 256 : 0;            // reduction indices of W lo that live in TMEM
    static constexpr int kWloSmem = (512 - kKT) / 64 * 16384;  // W lo plane in shared memory: k-blocks of [128 rows][64 k]
    static constexpr int kSH = NSTR * 2 * R2_CL * kTile;       // [stream][parity][source]
    static constexpr int kStage = NSTR * 2 * kTile;            // [stream][parity] own outgoing tile
    static constexpr int kSmem = kWloSmem + kSH + kStage + 1024 + 512;
    static constexpr int kWloCol = 256;                        // TMEM: W_hi columns [0,256), W_lo (k < kKT) [256, 256 + kKT/2)
    static constexpr int kAccCol = 384;                        // accumulators: 32 columns (W_hi.h_hi | W_hi.h_lo) per (stream, parity)
    static_assert(NSTR == 1 || NSTR == 2, "streams per cluster");
    static_assert(kAccCol + NSTR * 2 * 2 * NS <= 512 && kWloCol + kKT / 2 <= kAccCol, "TMEM budget");
};

// byte offset of (row, unit) inside a [rows][32] bf16 SWIZZLE_64B tile
__device__ __forceinline__ uint32_t sw64_off(int row, int unit) {
    return (uint32_t)((row >> 3) * 512 + (row & 7) * 64 + ((((unit >> 3) ^ (row >> 1)) & 3) << 4) + (unit & 7) * 2);
}
template <int NC, int N>
__device__ __forceinline__ float pickq(const float (&v)[N], int q, int i) {       // v[NC q + i], q in 0..3, without dynamic register indexing
    const float a = (q & 1) ? v[NC + i] : v[i], b = (q & 1) ? v[3 * NC + i] : v[2 * NC + i];
    return (q & 2) ? b : a;
}
__device__ __forceinline__ float4 ld_nc_f4(const float* p) {                      // issued where it is written (never sunk to the use)
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
template <int N> __device__ __forceinline__ void tmem_ld_n_issue(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_ld_n_issue<4>(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
template <> __device__ __forceinline__ void tmem_ld_n_issue<8>(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
template <int N> __device__ __forceinline__ void tmem_ld_n_wait(uint32_t (&a)[N], uint32_t (&b)[N]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(a[i]), "+r"(b[i]));          // no use of the registers may move above the wait
}

template <int NSTR>
__global__ void __launch_bounds__(R2_THREADS, 1) rec2_fwd_kernel(const RecFwdParams P, int nq) {
    using Cfg = Rec2Cfg<NSTR>;
    constexpr int NS = 16, TILE = Cfg::kTile, PLANE = Cfg::kPlane, NGRP = R2_CL / R2_GRP;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t sW = raw + pad;                     // W lo plane (k >= kKT)
    const uint32_t sH = sW + Cfg::kWloSmem;            // h tiles [stream][parity][source]
    const uint32_t sS = sH + Cfg::kSH;                 // staging [stream][parity]
    uint8_t* stage_ptr = smem + Cfg::kWloSmem + Cfg::kSH;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kWloSmem + Cfg::kSH + Cfg::kStage);       // full[stream][2][NGRP] | accf[stream][2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NSTR * (2 * NGRP + 2));
    auto bar_full = [&](int st, int par, int gr) { return smem_u32(&bars[(st * 2 + par) * NGRP + gr]); };
    auto bar_acc = [&](int st, int par) { return smem_u32(&bars[NSTR * 2 * NGRP + st * 2 + par]); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t c = cluster_ctarank();
    const int cid = blockIdx.x / R2_CL;
    const int net = P.net_base + cid / nq, quarter = cid % nq;      // quarter: which group of NSTR * 16 sequences
    const int B = P.B, T = P.T;
    const bool want_lo = !P.fast;

    if (tid == 0) {
        for (int i = 0; i < NSTR * (2 * NGRP + 2); ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == R2_EW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // W lo plane, k >= kKT -> shared memory (all threads), K-major SWIZZLE_128B, 16 KB per k-block of 64
    if (want_lo) {
        constexpr int CH0 = Cfg::kKT / 8, NCH = 64 - CH0;      // 16-byte chunks per row that go to shared memory
        for (int u = tid; u < 128 * NCH; u += R2_THREADS) {
            const int row = u / NCH, ch = CH0 + u % NCH, kb = (ch - CH0) >> 3, jj = ch & 7;
            const uint32_t dst = (uint32_t)(kb * 16384 + (row >> 3) * 1024 + (row & 7) * 128 + ((jj ^ (row & 7)) << 4));
            cp_async16(sW + dst, P.Wlo[net] + (size_t)(128 * c + row) * REC_H + ch * 8, true);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int q = warp & 3, cg = warp >> 2;            // epilogue warps: TMEM lane quadrant, sequence group (4 sequences)
    if (warp < R2_EW) {                                // W hi plane (and W lo, k < kKT) -> tensor memory: lane = gate row, column = k / 2
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int row = 128 * (int)c + 32 * q + lane;
        const uint4* wrow = reinterpret_cast<const uint4*>(P.Whi[net] + (size_t)row * REC_H);
#pragma unroll 1
        for (int h = 4 * cg; h < 4 * cg + 4; ++h) {    // the four warps of a quadrant split the 16 column blocks
            uint32_t r[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const uint4 x = __ldg(wrow + h * 4 + q4);
                r[4 * q4] = x.x; r[4 * q4 + 1] = x.y; r[4 * q4 + 2] = x.z; r[4 * q4 + 3] = x.w;
            }
            tmem_st16(lane_addr + (uint32_t)(h * 16), r);
        }
        if (want_lo && Cfg::kKT) {
            const uint4* lrow = reinterpret_cast<const uint4*>(P.Wlo[net] + (size_t)row * REC_H);
#pragma unroll 1
            for (int h = cg; h < Cfg::kKT / 32; h += 4) {
                uint32_t r[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint4 x = __ldg(lrow + h * 4 + q4);
                    r[4 * q4] = x.x; r[4 * q4 + 1] = x.y; r[4 * q4 + 2] = x.z; r[4 * q4 + 3] = x.w;
                }
                tmem_st16(lane_addr + (uint32_t)(Cfg::kWloCol + h * 16), r);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // every CTA's barriers exist before the first remote completion
    tc_fence_after();

    if (warp < R2_EW) {
        // ------------------------------------------------------------------ epilogue / cell warps: one cell per thread and stream
        const int j = 8 * q + (lane >> 2), g = lane & 3;           // unit inside the CTA, gate row (i, f, g, o)
        const int unit = 32 * (int)c + j;
        const int row0 = cg * 4 + g;                               // this thread's sequence inside a stream
        int bq[NSTR], blen[NSTR];
        float c_reg[NSTR];
        uint32_t hb[NSTR], lb[NSTR];                               // current h as bf16 bits (hi, lo)
        float4 xp[NSTR];
        const float* xp_base = P.XP[net] + 4 * unit;
        auto load_xp = [&](int st, int t) {
            xp[st] = (blen[st] >= 0 && t < T) ? ld_nc_f4(xp_base + ((size_t)t * B + bq[st]) * REC_G4) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // publish: own tile -> staging of (stream, parity), then one bulk copy per destination CTA
        auto publish = [&](int st, int par, bool send) {
            uint8_t* sp = stage_ptr + (st * 2 + par) * TILE;
            const uint32_t o = sw64_off(row0, j);
            *reinterpret_cast<uint16_t*>(sp + o) = (uint16_t)hb[st];
            *reinterpret_cast<uint16_t*>(sp + PLANE + o) = (uint16_t)lb[st];
            fence_proxy_async_smem();
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(32 * R2_EW) : "memory");
            if (send && tid < R2_CL) {
                const uint32_t dst = mapa_u32(sH + (uint32_t)(((st * 2 + par) * R2_CL + (int)c) * TILE), (uint32_t)tid);
                const uint32_t bar = mapa_u32(bar_full(st, par, (int)c / R2_GRP), (uint32_t)tid);
                bulk_copy_to_cluster(dst, sS + (st * 2 + par) * TILE, TILE, bar);
            }
        };
#pragma unroll
        for (int st = 0; st < NSTR; ++st) {
            bq[st] = NSTR * NS * quarter + NS * st + row0;
            const bool ok = bq[st] < B;
            blen[st] = ok ? P.len[bq[st]] : -1;                    // -1: no such sequence (never live, nothing stored)
            c_reg[st] = 0.f; hb[st] = 0u; lb[st] = 0u;
            if (ok) {
                c_reg[st] = P.c0[(size_t)bq[st] * P.ld_c0 + unit];
                hb[st] = __bfloat16_as_ushort(P.Hhi[net][(size_t)bq[st] * REC_H + unit]);      // HsX block 0 = stored h0 (split)
                lb[st] = __bfloat16_as_ushort(P.Hlo[net][(size_t)bq[st] * REC_H + unit]);
            }
            publish(st, 0, true);                                  // h_{-1}
            load_xp(st, 0);
        }

        for (int t = 0; t < T; ++t) {
            const int par = t & 1;
#pragma unroll
            for (int st = 0; st < NSTR; ++st) {
                const bool tr = P.trace && blockIdx.x == 0 && tid == 0 && st == 0;
                if (tr) P.trace[t * 8 + 0] = gtime();
                mbar_wait(bar_acc(st, par), ((uint32_t)t >> 1) & 1u);
                tc_fence_after();
                if (tr) P.trace[t * 8 + 1] = gtime();
                // accumulator: columns [0, 16) = W_hi.h_hi + W_lo.h_hi, [16, 32) = W_hi.h_lo; this warp reads its 4 sequences
                float v[4];
                {
                    const uint32_t a0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg::kAccCol + 32 * (st * 2 + par) + cg * 4);
                    uint32_t ra[4], rb[4];
                    tmem_ld_n_issue<4>(a0, ra);
                    tmem_ld_n_issue<4>(a0 + NS, rb);
                    tmem_ld_n_wait<4>(ra, rb);
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(ra[i]) + (want_lo ? __uint_as_float(rb[i]) : 0.f);   // fast mode never writes the second half
                }
                // 4-lane transpose: lane (unit, gate g) ends with the four gates of its own sequence
                float rv[3];
#pragma unroll
                for (int off = 1; off < 4; ++off) {
                    const int gd = (g - off) & 3;                  // the lane that reads from me wants its own sequence
                    rv[off - 1] = __shfl_sync(0xffffffffu, pickq<1>(v, gd, 0), (lane & ~3) | ((g + off) & 3));
                }
                float gt[4];
                const float own = pickq<1>(v, g, 0);
#pragma unroll
                for (int G = 0; G < 4; ++G) {
                    const int off = (G - g) & 3;                   // gate G sits `off` lanes further in the quad
                    gt[G] = off == 0 ? own : off == 1 ? rv[0] : off == 2 ? rv[1] : rv[2];
                }
                const float gi = fast_sigmoid(gt[0] + xp[st].x);
                const float gf = fast_sigmoid(gt[1] + xp[st].y);
                const float gg = fast_tanh(gt[2] + xp[st].z);
                const float go = fast_sigmoid(gt[3] + xp[st].w);
                const float cn = gf * c_reg[st] + gi * gg;
                const float hn = go * fast_tanh(cn);
                if (t < blen[st]) {
                    c_reg[st] = cn;
                    const bf16 hh = __float2bfloat16_rn(hn);
                    const bf16 ll = __float2bfloat16_rn(hn - __bfloat162float(hh));
                    hb[st] = __bfloat16_as_ushort(hh);
                    lb[st] = __bfloat16_as_ushort(ll);
                }
                if (tr) P.trace[t * 8 + 2] = gtime();
                publish(st, par ^ 1, t + 1 < T);       // h_t is the input of step t+1
                if (tr) P.trace[t * 8 + 3] = gtime();
                // ---- off the critical path: everything that goes to global memory, and the next step's input projection.
                // (fence.proxy.async above is a MEMBAR: it waits for this thread's outstanding global accesses, so they are
                // issued AFTER it and have a whole step to complete.)
                if (blen[st] >= 0) {
                    const size_t row = (size_t)t * B + bq[st];
                    if (P.Gs[net]) *reinterpret_cast<float4*>(P.Gs[net] + row * REC_G4 + 4 * unit) = make_float4(gi, gf, gg, go);
                    P.Cs[net][row * REC_H + unit] = c_reg[st];
                }
                load_xp(st, t + 1);
                // h_t -> HsX block t+1 (for the heads and the backward pass): 16-byte chunks of the staged tile
                if (tid < NS * 8) {
                    const int plane = tid / (NS * 4), row = (tid >> 2) % NS, ch = tid & 3;
                    const int b = NSTR * NS * quarter + NS * st + row;
                    if (b < B) {
                        const uint4 x = *reinterpret_cast<const uint4*>(stage_ptr + (st * 2 + (par ^ 1)) * TILE + plane * PLANE + (row >> 3) * 512 + (row & 7) * 64 +
                                                                       (((ch ^ (row >> 1)) & 3) << 4));
                        bf16* dstp = (plane ? P.Hlo[net] : P.Hhi[net]) + ((size_t)(t + 1) * B + b) * REC_H + 32 * (int)c + ch * 8;
                        *reinterpret_cast<uint4*>(dstp) = x;
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issue (whole warp loops, one lane issues)
        constexpr uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
        constexpr uint32_t idesc2 = idesc_base | ((uint32_t)(2 * NS >> 3) << 17), idesc1 = idesc_base | ((uint32_t)(NS >> 3) << 17);
        const bool leader = elect_one();
        const uint32_t uW = __shfl_sync(0xffffffffu, sW, 0), uH = __shfl_sync(0xffffffffu, sH, 0);
        const uint32_t uT = __shfl_sync(0xffffffffu, tmem_base, 0);
        for (int t = 0; t < T; ++t) {
            const int par = t & 1;
            const uint32_t ph = ((uint32_t)t >> 1) & 1u;
#pragma unroll
            for (int st = 0; st < NSTR; ++st) {
                if (leader)
                    for (int gr = 0; gr < NGRP; ++gr) mbar_arrive_expect_tx(bar_full(st, par, gr), R2_GRP * TILE);
                __syncwarp();
                const uint32_t acc = uT + (uint32_t)(Cfg::kAccCol + 32 * (st * 2 + par));
                const uint32_t hbase = uH + (uint32_t)((st * 2 + par) * R2_CL * TILE);
#pragma unroll
                for (int gr = 0; gr < NGRP; ++gr) {
                    mbar_wait(bar_full(st, par, gr), ph);
                    tc_fence_after();
                    if (P.trace && blockIdx.x == 0 && leader && gr == 0 && st == 0) P.trace[t * 8 + 6] = gtime();
                    if (leader) {
#pragma unroll
                        for (int ss = 0; ss < R2_GRP; ++ss) {
                            const int s = gr * R2_GRP + ss;
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const int kk = 32 * s + 16 * k;                                               // first reduction index of this MMA
                                const uint64_t b_hl = umma_desc_sw64(hbase + (uint32_t)(s * TILE + k * 32));   // rows [0, 16) = hi plane, [16, 32) = lo plane
                                const uint32_t a_t = uT + (uint32_t)(kk >> 1);
                                if (want_lo) {
                                    umma_bf16_ts(acc, a_t, b_hl, idesc2, (s | k) ? 1u : 0u);                  // W_hi . [h_hi | h_lo]
                                    if (kk < Cfg::kKT) umma_bf16_ts(acc, uT + (uint32_t)(Cfg::kWloCol + (kk >> 1)), b_hl, idesc1, 1u);     // W_lo . h_hi
                                    else umma_bf16(acc, umma_desc_sw128(uW + (uint32_t)(((kk - Cfg::kKT) >> 6) * 16384 + (kk & 63) * 2)), b_hl, idesc1, 1u);
                                } else {
                                    umma_bf16_ts(acc, a_t, b_hl, idesc1, (s | k) ? 1u : 0u);
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                if (leader) umma_commit(bar_acc(st, par));
                if (P.trace && blockIdx.x == 0 && leader && st == 0) P.trace[t * 8 + 7] = gtime();
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // nobody exits while a peer may still copy from its staging tile
    if (warp == R2_EW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <int NSTR>
static inline cudaLaunchConfig_t rec2_config(int clusters, cudaStream_t s, cudaLaunchAttribute* attr) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * R2_CL);
    cfg.blockDim = dim3(R2_THREADS);
    cfg.dynamicSmemBytes = Rec2Cfg<NSTR>::kSmem;
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = R2_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}

// how many 16-CTA clusters of the recurrence kernel the device keeps resident at once (0: clusters of 16 are not schedulable)
template <int NSTR>
static inline int rec2_max_active_clusters() {
    static int cached = -1;
    if (cached >= 0) return cached;
    static unsigned long long configured = 0;
    if (ensure_dynamic_smem(rec2_fwd_kernel<NSTR>, Rec2Cfg<NSTR>::kSmem, &configured) != cudaSuccess) return 0;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = rec2_config<NSTR>(8, nullptr, attr);
    int n = 0;
    cudaError_t e = cudaFuncSetAttribute(rec2_fwd_kernel<NSTR>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, rec2_fwd_kernel<NSTR>, &cfg);
    if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    return cached = n;
}

extern int g_rec2_ns;     // 0 = automatic, 16 / 32 = forced sequences per cluster (1 / 2 interleaved streams of 16)

// Returns cudaErrorNotSupported when clusters of 16 CTAs with this much shared memory cannot be scheduled on the device.
static inline cudaError_t launch_rec2_fwd(const RecFwdParams& P, int nets, cudaStream_t s) {
    cudaLaunchAttribute attr[1];
    const int cap1 = rec2_max_active_clusters<1>(), cap2 = rec2_max_active_clusters<2>();
    const int need1 = nets * ((P.B + 15) / 16), need2 = nets * ((P.B + 31) / 32);
    // one stream per cluster has the shorter step; use it when all its clusters are resident at once (or a batch leaves the
    // second stream empty anyway), otherwise interleave two streams per cluster
    int nstr = (cap1 >= need1 || cap2 == 0 || P.B <= 16) ? 1 : 2;
    if (g_rec2_ns == 16) nstr = 1;
    if (g_rec2_ns == 32) nstr = 2;
    if ((nstr == 1 ? cap1 : cap2) < 1) return cudaErrorNotSupported;
    if (nstr == 1) {
        cudaLaunchConfig_t cfg = rec2_config<1>(need1, s, attr);
        return cudaLaunchKernelEx(&cfg, rec2_fwd_kernel<1>, P, (P.B + 15) / 16);
    }
    cudaLaunchConfig_t cfg = rec2_config<2>(need2, s, attr);
    return cudaLaunchKernelEx(&cfg, rec2_fwd_kernel<2>, P, (P.B + 31) / 32);
}

// ================================================================================================
// Cluster-resident BPTT recurrence (online network): dh_{t-1} = dgates_t . W_hh without leaving the SMs.
//
//   One cluster of 16 CTAs per 16 sequences.  CTA c owns hidden units [32c, 32c+32) = gate rows [128c, 128c+128):
//   * pointwise role (512 threads = 32 units x 16 sequences): dh_t = dH[t] + sum of the 16 partials received for step
//     t+1 (fixed order: deterministic), LSTM cell backward -> dgates_t (4 per cell), carried dc in a register;
//     dgates_t goes to global memory (split, for the weight-gradient GEMMs) and, as bf16 hi|lo, into shared memory as
//     the B operand [16 sequences][128 own gate rows];
//   * GEMM role: partial[512 units][16] = W_hh[own 128 gate rows, :]^T . dgates_t  -- the K-slice of the product that
//     this CTA can compute from its OWN dgates.  A = W_hh^T block [512 units][128 k]: hi plane resident in TENSOR
//     MEMORY (4 M-tiles x 64 columns), lo plane in shared memory (128 KB); 64 MMAs (M=128) per step;
//   * reduce-scatter: rows [32d, 32d+32) of the partial belong to CTA d: each epilogue thread pushes its 64 bytes
//     straight from registers into CTA d's receive slot with st.async (distributed shared memory, completes on CTA d's
//     mbarrier).  No global memory, flags or fences on the critical path.
// ================================================================================================
constexpr int RB2_WLO = 4 * 2 * 16384;             // W^T lo plane: [M-tile][k-block][128 units][64 k]
constexpr int RB2_BOP = 2 * 4096;                  // dgates as B operand: [k-block][hi 16 rows | lo 16 rows][128 B]
constexpr int RB2_SLOT = 2048;                     // one source's partial for my 32 units: [32][16] fp32
constexpr int RB2_RECV = 2 * R2_CL * RB2_SLOT;     // [parity][source]
constexpr int RB2_SMEM = RB2_WLO + RB2_BOP + RB2_RECV + 1024 + 256;
constexpr int RB2_THREADS = 32 * R2_EW + 128;     // 16 pointwise / epilogue warps + 4 MMA-issue warps
constexpr int RB2_ACC_COL = 256;                   // TMEM: W^T hi in columns [0,256), accumulators 4 x 32 columns at 256

__device__ __forceinline__ void st_async_v4(uint32_t dst, uint32_t mbar, float a, float b, float c, float d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(dst), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(mbar) : "memory");
}
__device__ __forceinline__ float ld_nc_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(RB2_THREADS, 1) rec2_bwd_kernel(const RecBwdParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    const uint32_t sW = raw + pad;                     // W^T lo plane
    const uint32_t sB = sW + RB2_WLO;                  // dgates B operand
    const uint32_t sR = sB + RB2_BOP;                  // receive slots [parity][source]
    uint8_t* bop_ptr = smem + RB2_WLO;
    const float* recv_ptr = reinterpret_cast<const float*>(smem + RB2_WLO + RB2_BOP);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RB2_WLO + RB2_BOP + RB2_RECV);       // recv_full[2] | accf[4] (one per M-tile)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t c = cluster_ctarank();
    const int quarter = blockIdx.x / R2_CL;
    const int B = P.B, T = P.T;
    const bool want_lo = !P.fast;

    if (blockIdx.x == 0 && tid == 0 && P.started != nullptr) atomicAdd(P.started, 1u);     // "the clusters are resident from here on"
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == R2_EW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // W^T lo plane -> shared memory: tile (m, kb) = units [128m, +128) x own gate rows [64 kb, +64), K-major SWIZZLE_128B
    if (want_lo) {
        for (int u = tid; u < 4 * 2 * 128 * 8; u += RB2_THREADS) {
            const int jj = u & 7, row = (u >> 3) & 127, tile = u >> 10, m = tile >> 1, kb = tile & 1;
            const uint32_t dst = (uint32_t)(tile * 16384 + (row >> 3) * 1024 + (row & 7) * 128 + ((jj ^ (row & 7)) << 4));
            cp_async16(sW + dst, P.WTlo + (size_t)(128 * m + row) * REC_G4 + 128 * (int)c + 64 * kb + 8 * jj, true);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int q = warp & 3, m = warp >> 2;             // epilogue role: TMEM lane quadrant, M-tile  (destination CTA = 4 m + q)
    if (warp < R2_EW) {                                // W^T hi plane -> tensor memory: lane = unit (within M-tile m), column = 64 m + k / 2
        const uint4* wrow = reinterpret_cast<const uint4*>(P.WThi + (size_t)(128 * m + 32 * q + lane) * REC_G4 + 128 * (int)c);
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
            uint32_t r[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const uint4 x = __ldg(wrow + h * 4 + q4);
                r[4 * q4] = x.x; r[4 * q4 + 1] = x.y; r[4 * q4 + 2] = x.z; r[4 * q4 + 3] = x.w;
            }
            tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 * m + 16 * h), r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();

    if (warp < R2_EW) {
        // pointwise ownership: sequence pb (fastest: conflict-free reads of the receive slots), unit pj
        const int pb = tid & 15, pj = tid >> 4;
        const int unit = 32 * (int)c + pj;
        const int b = 16 * quarter + pb;
        const bool own = b < B;
        const int len = own ? P.len[b] : 0;
        float dcrec = 0.f;
        // B-operand position of this thread's 4 gate values: row pb, k = 4 pj .. 4 pj + 3 of the own 128 gate rows
        const uint32_t bop_off = (uint32_t)((pj >> 4) * 4096 + (pb >> 3) * 1024 + (pb & 7) * 128 + (((((pj & 15) >> 1)) ^ (pb & 7)) << 4) + (pj & 1) * 8);
        // operands of the cell backward that do not depend on the recurrence, fetched one step ahead
        float4 g_n = make_float4(0.f, 0.f, 0.f, 0.f);
        float dh_n = 0.f, ct_n = 0.f, cp_n = 0.f;
        auto prefetch = [&](int t) {
            if (own && t >= 0 && t < len) {
                const size_t row = (size_t)t * B + b;
                g_n = ld_nc_f4(P.Gs + row * REC_G4 + 4 * unit);
                dh_n = ld_nc_f1(P.dH + row * REC_H + unit);
                ct_n = ld_nc_f1(P.Cs + row * REC_H + unit);
                cp_n = t ? ld_nc_f1(P.Cs + ((size_t)(t - 1) * B + b) * REC_H + unit) : ld_nc_f1(P.c0 + (size_t)b * P.ld_c0 + unit);
            }
        };
        prefetch(T - 1);
        for (int t = T - 1, step = 0; t >= 0; --t, ++step) {
            const float4 g = g_n;
            float dh = dh_n;
            const float ct = ct_n, cp = cp_n;
            const bool live = own && t < len;
            const bool tr = P.trace && blockIdx.x == 0 && tid == 0;
            if (tr) P.trace[step * 8 + 0] = gtime();
            if (step > 0) {                            // partials of dgates_{t+1} . W_hh for my units have landed (parity of t+1)
                const int par = (t + 1) & 1;
                if (tid == 0) mbar_arrive_expect_tx(smem_u32(&bars[par]), R2_CL * RB2_SLOT);
                mbar_wait(smem_u32(&bars[par]), ((uint32_t)(step - 1) >> 1) & 1u);
                const float* rp = recv_ptr + (size_t)par * R2_CL * (RB2_SLOT / 4) + pj * 16 + pb;
                float acc = 0.f;
#pragma unroll
                for (int s = 0; s < R2_CL; ++s) acc += rp[s * (RB2_SLOT / 4)];
                dh += acc;
            }
            if (tr) P.trace[step * 8 + 1] = gtime() + (unsigned long long)(dh == 12345.f);
            float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
            if (live) {
                const float tc = fast_tanh(ct);
                const float dc = dcrec + dh * g.w * (1.f - tc * tc);
                o0 = dc * g.z * g.x * (1.f - g.x);
                o1 = dc * cp * g.y * (1.f - g.y);
                o2 = dc * g.x * (1.f - g.z * g.z);
                o3 = dh * tc * g.w * (1.f - g.w);
                dcrec = dc * g.y;
            }
            uint32_t h2[2], l2[2];
            split2(o0, o1, h2[0], l2[0]);
            split2(o2, o3, h2[1], l2[1]);
            if (t > 0) {
                *reinterpret_cast<uint2*>(bop_ptr + bop_off) = make_uint2(h2[0], h2[1]);
                *reinterpret_cast<uint2*>(bop_ptr + 2048 + bop_off) = make_uint2(l2[0], l2[1]);
                fence_proxy_async_smem();
            }
            tc_fence_before();
            if (tr) P.trace[step * 8 + 2] = gtime();
            asm volatile("bar.sync 2, %0;" ::"n"(RB2_THREADS) : "memory");       // B operand complete -> MMA warps
            if (tr) P.trace[step * 8 + 3] = gtime();
            // ---- off the critical path: dgates_t to global memory, operands of the next step
            if (own) {
                const size_t o = ((size_t)t * B + b) * REC_G4 + 4 * unit;
                *reinterpret_cast<uint2*>(P.DGhi + o) = make_uint2(h2[0], h2[1]);
                *reinterpret_cast<uint2*>(P.DGlo + o) = make_uint2(l2[0], l2[1]);
            }
            if (t == 0) break;                          // dh_{-1} is not needed
            prefetch(t - 1);
            // ---- reduce-scatter of this step's partial: warp (q, m) holds units [128 m + 32 q, +32) = CTA 4 m + q
            mbar_wait(smem_u32(&bars[2 + m]), (uint32_t)step & 1u);
            tc_fence_after();
            if (tr) P.trace[step * 8 + 4] = gtime();
            float v[16];
            {
                const uint32_t a0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(RB2_ACC_COL + 32 * m);
                uint32_t ra[16], rb[16];
                tmem_ld16_issue(a0, ra);
                tmem_ld16_issue(a0 + 16, rb);
                tmem_ld_wait(ra);
                tmem_ld_wait(rb);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]) + (want_lo ? __uint_as_float(rb[i]) : 0.f);
            }
            {
                const int par = t & 1;
                const uint32_t d = (uint32_t)(4 * m + q);
                const uint32_t dst = mapa_u32(sR + (uint32_t)((par * R2_CL + (int)c) * RB2_SLOT + lane * 64), d);
                const uint32_t bar = mapa_u32(smem_u32(&bars[par]), d);
#pragma unroll
                for (int i = 0; i < 4; ++i) st_async_v4(dst + 16 * i, bar, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (tr) P.trace[step * 8 + 5] = gtime();
            tc_fence_before();
        }
    } else {
        // ------------------------------------------------------------------ MMA issue: FOUR warps, one per M-tile of 128 units (a single
        // issuing thread needs ~1 us for the step's 64 MMAs -- instruction-issue bound -- on a 4 us critical path)
        constexpr uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
        constexpr uint32_t idesc2 = idesc_base | ((32u >> 3) << 17), idesc1 = idesc_base | ((16u >> 3) << 17);
        const int mm = warp - R2_EW;
        const bool leader = elect_one();
        const uint32_t uW = __shfl_sync(0xffffffffu, sW, 0), uB = __shfl_sync(0xffffffffu, sB, 0);
        const uint32_t uT = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t acc = uT + (uint32_t)(RB2_ACC_COL + 32 * mm);
        for (int t = T - 1; t >= 0; --t) {
            asm volatile("bar.sync 2, %0;" ::"n"(RB2_THREADS) : "memory");
            if (t == 0) break;
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint64_t b_hl = umma_desc_sw128(uB + (uint32_t)((kk >> 2) * 4096 + (kk & 3) * 32));      // rows 0-15 hi, 16-31 lo
                    const uint32_t a_t = uT + (uint32_t)(64 * mm + 8 * kk);
                    if (want_lo) {
                        umma_bf16_ts(acc, a_t, b_hl, idesc2, kk ? 1u : 0u);
                        umma_bf16(acc, umma_desc_sw128(uW + (uint32_t)((mm * 2 + (kk >> 2)) * 16384 + (kk & 3) * 32)), b_hl, idesc1, 1u);
                    } else {
                        umma_bf16_ts(acc, a_t, b_hl, idesc1, kk ? 1u : 0u);
                    }
                }
                umma_commit(smem_u32(&bars[2 + mm]));
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == R2_EW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

static inline cudaError_t launch_rec2_bwd(const RecBwdParams& P, cudaStream_t s) {
    static unsigned long long configured = 0;
    static int usable = -1;
    {
        cudaError_t e = ensure_dynamic_smem(rec2_bwd_kernel, RB2_SMEM, &configured);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(((P.B + 15) / 16) * R2_CL);
    cfg.blockDim = dim3(RB2_THREADS);
    cfg.dynamicSmemBytes = RB2_SMEM;
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = R2_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (usable < 0) {
        cudaError_t e = cudaFuncSetAttribute(rec2_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int n = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, rec2_bwd_kernel, &cfg);
        usable = (e == cudaSuccess && n >= 1) ? 1 : 0;
        (void)cudaGetLastError();
    }
    if (!usable) return cudaErrorNotSupported;
    return cudaLaunchKernelEx(&cfg, rec2_bwd_kernel, P);
}

}  // namespace r2d2
