// K4: HBM-resident replay block store: ingest from pinned staging and batch gather.
//
// Replaces the storage half of the reference's ReplayBuffer (worker.py:141-240): `add` keeps a ring of
// Blocks (worker.py:23-35) and `sample_batch` slices [start-burn_in : start+learning+forward] windows out of
// them in a Python loop and zero-pads with pad_sequence.  Here every block is ONE fixed-size blob in HBM
// (same arrays, fixed offsets, 256-byte aligned sections) so that
//   * ingest is a single cudaMemcpyAsync of the actor's block from a pinned staging slot, on a side stream;
//   * gather is two launches: a one-CTA metadata pass (window bounds, ragged row offsets) and a
//     (T x B)-CTA copy pass moving 16-byte chunks (a C x 84 x 84 u8 frame is 441*C chunks).
// Bound: HBM.  Algorithmic bytes per sampled sequence: 2 * (b+l+f) * 7056 * C  (read + write) + ~10 KB side data.
#include <cuda_bf16.h>

#include "common.cuh"

namespace r2d2 {

struct ReplayLayout {
    int64_t obs, last_action, last_reward, action, n_step_reward, gamma, hidden, burn, learn, fwd, num_seq, total;
};

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct SeqDesc {            // one sampled sequence, produced by the metadata pass
    int64_t blob;           // byte offset of its block
    int start;              // first frame of the window inside the block's frame array
    int b, l, f;            // burn-in / learning / forward counts
    int row_off;            // first row in the ragged (sequence-major) arrays
    int seq;                // sequence index inside the block
    int act_off;            // first learning step inside the block's per-step arrays
};

}  // namespace r2d2

struct r2d2_replay {
    int num_blocks, block_len, burn_in, learning, forward, C, A, H, spb, frames_per_block;
    int64_t frame_bytes;
    r2d2::ReplayLayout lay;
    uint8_t* store;          // [num_blocks][lay.total]
    r2d2::SeqDesc* desc;     // [desc_cap]
    int desc_cap;
    int copy_smem;           // dynamic shared memory the gather's copy CTAs ask for without using it (r2d2_replay_set_copy_smem)
};

namespace r2d2 {

__global__ void replay_meta_kernel(const uint8_t* __restrict__ store, ReplayLayout lay, int spb, int num_blocks, const int64_t* __restrict__ idx,
                                   int B, SeqDesc* __restrict__ desc, uint8_t* __restrict__ burn_out, uint8_t* __restrict__ learn_out,
                                   uint8_t* __restrict__ fwd_out, int32_t* __restrict__ rows_out) {
    extern __shared__ int s_l[];
    for (int n = threadIdx.x; n < B; n += blockDim.x) {
        int64_t slot = idx[n];
        const bool valid = slot >= 0 && slot < (int64_t)num_blocks * spb;     // a stale / corrupt index yields an EMPTY sequence, never an
        if (!valid) slot = 0;                                                  // out-of-bounds read (the reference raises IndexError here)
        const int64_t blk = slot / spb;
        const int s = (int)(slot - blk * spb);
        const uint8_t* blob = store + blk * lay.total;
        const uint8_t* burn = blob + lay.burn;
        const uint8_t* learn = blob + lay.learn;
        int before = 0;                                  // worker.py:186,193: sum(learning_steps[:seq])
        for (int i = 0; i < s; ++i) before += learn[i];
        SeqDesc d;
        d.blob = blk * lay.total;
        d.b = valid ? burn[s] : 0; d.l = valid ? learn[s] : 0; d.f = valid ? blob[lay.fwd + s] : 0;
        d.start = (int)burn[0] + before - d.b;           // window [start-b, start+l+f) with start = burn[0] + before
        d.seq = s;
        d.act_off = before;
        d.row_off = 0;
        desc[n] = d;
        s_l[n] = d.l;
        burn_out[n] = (uint8_t)d.b; learn_out[n] = (uint8_t)d.l; fwd_out[n] = (uint8_t)d.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int n = 0; n < B; ++n) { desc[n].row_off = run; run += s_l[n]; }
        *rows_out = run;
    }
}

__global__ void __launch_bounds__(256) replay_copy_kernel(const uint8_t* __restrict__ store, ReplayLayout lay, const SeqDesc* __restrict__ desc,
                                                          const float* __restrict__ isw, int T, int C, int A, int H, int64_t frame_bytes,
                                                          uint8_t* __restrict__ obs, __nv_bfloat16* __restrict__ s2d,
                                                          uint8_t* __restrict__ last_action,
                                                          float* __restrict__ last_reward, float* __restrict__ hidden,
                                                          uint8_t* __restrict__ action, float* __restrict__ nsr, float* __restrict__ gam,
                                                          float* __restrict__ isw_rows) {
    const int t = blockIdx.x, n = blockIdx.y;
    const SeqDesc d = desc[n];
    const uint8_t* blob = store + d.blob;
    const int len = d.b + d.l + d.f;
    const int64_t chunks = frame_bytes / 16;
    uint4* dst = obs ? reinterpret_cast<uint4*>(obs + ((int64_t)n * T + t) * frame_bytes) : nullptr;
    // fused space-to-depth output (the layout conv1 consumes, net.cu): [Y][X][c*16 + r*4 + q] bf16, pixel (c, 4Y+r, 4X+q)
    __nv_bfloat16* sdst = s2d ? s2d + ((int64_t)n * T + t) * (441 * 16 * C) : nullptr;
    if (sdst) {
        // item = (c, Y, X): four coalesced 32-bit reads (rows 4Y..4Y+3, pixels 4X..4X+3) -> one 32-byte sector of bf16
        const int items = C * 441;
        const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(blob + lay.obs + (int64_t)(d.start + (t < len ? t : 0)) * frame_bytes);
        // two items per thread and iteration: eight independent loads in flight, one 256-bit store per item
        for (int i0 = threadIdx.x; i0 < items; i0 += 2 * blockDim.x) {
            uint32_t w[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = i0 + u * blockDim.x;
                const int X = i % 21, Y = (i / 21) % 21, c = i / 441;
#pragma unroll
                for (int r = 0; r < 4; ++r) w[u][r] = (t < len && i < items) ? __ldg(wsrc + (c * 84 + 4 * Y + r) * 21 + X) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = i0 + u * blockDim.x;
                if (i >= items) break;
                const int X = i % 21, Y = (i / 21) % 21, c = i / 441;
                uint32_t o[8];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const __nv_bfloat162 p0 = __floats2bfloat162_rn((float)(w[u][r] & 255u), (float)((w[u][r] >> 8) & 255u));
                    const __nv_bfloat162 p1 = __floats2bfloat162_rn((float)((w[u][r] >> 16) & 255u), (float)(w[u][r] >> 24));
                    o[2 * r] = *reinterpret_cast<const uint32_t*>(&p0);
                    o[2 * r + 1] = *reinterpret_cast<const uint32_t*>(&p1);
                }
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(sdst + ((Y * 21 + X) * 16 * C + c * 16)), "r"(o[0]),
                             "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
            }
        }
    }
    if (t < len) {
        const uint4* src = reinterpret_cast<const uint4*>(blob + lay.obs + (int64_t)(d.start + t) * frame_bytes);
        if (dst) for (int64_t i = threadIdx.x; i < chunks; i += blockDim.x) dst[i] = __ldg(src + i);
        if (threadIdx.x < A) last_action[((int64_t)n * T + t) * A + threadIdx.x] = blob[lay.last_action + (int64_t)(d.start + t) * A + threadIdx.x];
        if (threadIdx.x == 0)
            last_reward[(int64_t)n * T + t] = reinterpret_cast<const float*>(blob + lay.last_reward)[d.start + t];
    } else {                                           // pad_sequence zero padding at the END (worker.py:212-214)
        const uint4 z = make_uint4(0, 0, 0, 0);
        if (dst) for (int64_t i = threadIdx.x; i < chunks; i += blockDim.x) dst[i] = z;
        if (threadIdx.x < A) last_action[((int64_t)n * T + t) * A + threadIdx.x] = 0;
        if (threadIdx.x == 0) last_reward[(int64_t)n * T + t] = 0.f;
    }
    if (t == 0) {                                      // per-sequence side data (worker.py:193-198,216)
        const float* hsrc = reinterpret_cast<const float*>(blob + lay.hidden) + (int64_t)d.seq * 2 * H;
        for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) hidden[(int64_t)n * 2 * H + i] = hsrc[i];
        const float w = isw[n];
        for (int i = threadIdx.x; i < d.l; i += blockDim.x) {
            action[d.row_off + i] = blob[lay.action + d.act_off + i];
            nsr[d.row_off + i] = reinterpret_cast<const float*>(blob + lay.n_step_reward)[d.act_off + i];
            gam[d.row_off + i] = reinterpret_cast<const float*>(blob + lay.gamma)[d.act_off + i];
            isw_rows[d.row_off + i] = w;
        }
    }
}

}  // namespace r2d2

using namespace r2d2;

extern "C" {

/* ReplayBuffer storage (worker.py:43-48,70): num_blocks ring slots of one Block each.
 * Block arrays (worker.py:23-35) inside a slot, at the byte offsets r2d2_replay_layout reports:
 *   obs u8 [burn_in+block_len+1][C][84][84] | last_action u8 [frames][A] | last_reward f32 [frames] |
 *   action u8 [block_len] | n_step_reward f32 [block_len] | gamma f32 [block_len] |
 *   hidden f32 [seq_per_block][2][H] | burn/learn/fwd u8 [seq_per_block] each | num_sequences i32 */
int r2d2_replay_create(int num_blocks, int block_len, int burn_in, int learning, int forward, int C, int A, int H,
                       r2d2_replay** out) {
    R2D2_REQUIRE(out && num_blocks >= 1 && block_len >= 1 && learning >= 1 && block_len % learning == 0 && burn_in >= 0 &&
                     forward >= 1 && C >= 1 && A >= 1 && A <= 255 && H >= 1,
                 "bad replay shape");
    r2d2_replay* r = new r2d2_replay();
    memset(r, 0, sizeof(*r));
    r->num_blocks = num_blocks; r->block_len = block_len; r->burn_in = burn_in; r->learning = learning; r->forward = forward;
    r->C = C; r->A = A; r->H = H; r->spb = block_len / learning;
    r->frames_per_block = burn_in + block_len + 1;
    r->frame_bytes = (int64_t)C * 84 * 84;
    ReplayLayout& l = r->lay;
    int64_t o = 0;
    l.obs = o; o = align_up(o + r->frames_per_block * r->frame_bytes, 256);
    l.last_action = o; o = align_up(o + (int64_t)r->frames_per_block * A, 256);
    l.last_reward = o; o = align_up(o + (int64_t)r->frames_per_block * 4, 256);
    l.action = o; o = align_up(o + block_len, 256);
    l.n_step_reward = o; o = align_up(o + (int64_t)block_len * 4, 256);
    l.gamma = o; o = align_up(o + (int64_t)block_len * 4, 256);
    l.hidden = o; o = align_up(o + (int64_t)r->spb * 2 * H * 4, 256);
    l.burn = o; o += r->spb; l.learn = o; o += r->spb; l.fwd = o; o = align_up(o + r->spb, 16);
    l.num_seq = o; o = align_up(o + 4, 256);
    l.total = o;
    R2D2_CUDA_CHECK(cudaMalloc(&r->store, (size_t)num_blocks * l.total));
    R2D2_CUDA_CHECK(cudaMemset(r->store, 0, (size_t)num_blocks * l.total));
    r->desc_cap = 0; r->desc = nullptr;
    R2D2_CUDA_CHECK(cudaDeviceSynchronize());
    *out = r;
    return R2D2_OK;
}

int r2d2_replay_destroy(r2d2_replay* r) {
    if (!r) return R2D2_OK;
    cudaFree(r->store);
    if (r->desc) cudaFree(r->desc);
    delete r;
    return R2D2_OK;
}

/* offsets_out[12]: obs, last_action, last_reward, action, n_step_reward, gamma, hidden, burn, learn, fwd, num_seq, total */
int r2d2_replay_layout(const r2d2_replay* r, int64_t* offsets_out) {
    R2D2_REQUIRE(r && offsets_out, "bad arguments");
    const ReplayLayout& l = r->lay;
    const int64_t v[12] = {l.obs, l.last_action, l.last_reward, l.action, l.n_step_reward, l.gamma, l.hidden, l.burn, l.learn, l.fwd,
                           l.num_seq, l.total};
    for (int i = 0; i < 12; ++i) offsets_out[i] = v[i];
    return R2D2_OK;
}

/* ReplayBuffer.add storage half (worker.py:154): one async H2D copy of a packed block from PINNED host memory into
 * ring slot `block_idx`.  `nbytes` <= layout total (the tail of a short block need not be sent). */
int r2d2_replay_ingest(r2d2_replay* r, int block_idx, const void* host_blob, int64_t nbytes, void* stream) {
    R2D2_REQUIRE(r && host_blob && block_idx >= 0 && block_idx < r->num_blocks && nbytes > 0 && nbytes <= r->lay.total, "bad arguments");
    R2D2_CUDA_CHECK(cudaMemcpyAsync(r->store + (size_t)block_idx * r->lay.total, host_blob, (size_t)nbytes, cudaMemcpyHostToDevice,
                                    as_stream(stream)));
    return R2D2_OK;
}

/* ReplayBuffer.sample_batch slicing half (worker.py:172-238) for B sampled slots (device int64) and their IS weights.
 * Outputs (device): obs u8 [B][T][C][84][84], last_action u8 [B][T][A], last_reward f32 [B][T], hidden f32 [B][2][H],
 * action u8 [rows], n_step_reward f32 [rows], gamma f32 [rows], burn/learn/fwd u8 [B], is_weights f32 [rows] (per-sequence
 * weight repeated over its learning steps), rows_out i32[1].  T >= burn_in+learning+forward; ragged arrays need capacity
 * B*learning. */
static int replay_gather_impl(r2d2_replay* r, const int64_t* idx, const float* isw, int B, int T, uint8_t* obs, void* s2d,
                              uint8_t* last_action, float* last_reward, float* hidden, uint8_t* action, float* n_step_reward,
                              float* gamma, uint8_t* burn, uint8_t* learn, uint8_t* fwd, float* is_weights_rows, int32_t* rows_out,
                              void* stream) {
    R2D2_REQUIRE(r && idx && isw && (obs || s2d) && last_action && last_reward && hidden && action && n_step_reward && gamma && burn && learn &&
                     fwd && is_weights_rows && rows_out && B >= 1 && T >= r->burn_in + r->learning + r->forward,
                 "bad arguments");
    if (B > r->desc_cap) {
        if (r->desc) cudaFree(r->desc);
        R2D2_CUDA_CHECK(cudaMalloc(&r->desc, (size_t)B * sizeof(SeqDesc)));
        r->desc_cap = B;
    }
    cudaStream_t s = as_stream(stream);
    replay_meta_kernel<<<1, 256, B * sizeof(int), s>>>(r->store, r->lay, r->spb, r->num_blocks, idx, B, r->desc, burn, learn, fwd, rows_out);
    replay_copy_kernel<<<dim3(T, B), 256, r->copy_smem, s>>>(r->store, r->lay, r->desc, isw, T, r->C, r->A, r->H, r->frame_bytes, obs,
                                                  (__nv_bfloat16*)s2d, last_action,
                                                  last_reward, hidden, action, n_step_reward, gamma, is_weights_rows);
    R2D2_LAUNCH_CHECK();
    return R2D2_OK;
}

/* Placement control for a gather that runs next to a learner update (worker.py:309-316 prefetches its batches the same way):
 * with `bytes` of (unused) dynamic shared memory per copy CTA, the CTAs do not fit on an SM that holds a ~200 KB GEMM /
 * recurrence CTA, so the copy only takes SMs the update leaves idle (r2d2_net_shadow_gate); 0 = anywhere (default). */
int r2d2_replay_set_copy_smem(r2d2_replay* r, int bytes) {
    R2D2_REQUIRE(r && bytes >= 0 && bytes <= 48 * 1024, "bad arguments");
    static unsigned long long configured = 0;
    int dev = 0;
    R2D2_CUDA_CHECK(cudaGetDevice(&dev));
    if (!(configured & (1ull << (dev & 63)))) {
        R2D2_CUDA_CHECK(cudaFuncSetAttribute(replay_copy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        R2D2_CUDA_CHECK(cudaFuncSetAttribute(replay_meta_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured |= 1ull << (dev & 63);
    }
    r->copy_smem = bytes;
    return R2D2_OK;
}

int r2d2_replay_gather(r2d2_replay* r, const int64_t* idx, const float* isw, int B, int T, uint8_t* obs, uint8_t* last_action,
                       float* last_reward, float* hidden, uint8_t* action, float* n_step_reward, float* gamma, uint8_t* burn,
                       uint8_t* learn, uint8_t* fwd, float* is_weights_rows, int32_t* rows_out, void* stream) {
    return replay_gather_impl(r, idx, isw, B, T, obs, nullptr, last_action, last_reward, hidden, action, n_step_reward, gamma, burn,
                              learn, fwd, is_weights_rows, rows_out, stream);
}

/* Same gather with the frames written directly in the network's space-to-depth bf16 staging layout
 * (s2d_out = r2d2_net_s2d_buffer(net), [B*T][21][21][16C]) instead of as raw u8 frames: the learner then calls
 * r2d2_net_forward_pair with obs == NULL and one 2 x 85 x 7056 x C byte round trip through HBM per sequence disappears. */
int r2d2_replay_gather_s2d(r2d2_replay* r, const int64_t* idx, const float* isw, int B, int T, void* s2d_out, uint8_t* last_action,
                           float* last_reward, float* hidden, uint8_t* action, float* n_step_reward, float* gamma, uint8_t* burn,
                           uint8_t* learn, uint8_t* fwd, float* is_weights_rows, int32_t* rows_out, void* stream) {
    R2D2_REQUIRE(s2d_out, "null s2d buffer");
    return replay_gather_impl(r, idx, isw, B, T, nullptr, s2d_out, last_action, last_reward, hidden, action, n_step_reward, gamma,
                              burn, learn, fwd, is_weights_rows, rows_out, stream);
}

}  // extern "C"
