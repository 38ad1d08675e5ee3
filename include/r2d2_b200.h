/*
 * r2d2_b200 -- C ABI of the B200-native R2D2 learner hot path.
 *
 * The reference (ZiyuanMa/R2D2) is pure Python and has no FFI layer; its drop-in
 * boundary is the Python class surface train.py imports.
 * This header is the flat C boundary our Python mirror of that surface
 * (r2d2_b200/{priority_tree,worker,model}.py) binds with ctypes; every entry
 * point cites the reference interface it replaces (file:line in the upstream
 * repository).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; r2d2_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - all data pointers are DEVICE pointers unless the name says host; buffers are
 *     caller-owned (e.g. torch tensors' data_ptr()); handles are library-owned
 *     and freed by the matching *_destroy;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *     nothing synchronises the host unless documented;
 *   - one caller thread per handle.
 */
#ifndef R2D2_B200_H
#define R2D2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R2D2_OK 0
#define R2D2_ERR_ARG (-1)
#define R2D2_ERR_CUDA (-2)
#define R2D2_ERR_STATE (-3)

const char* r2d2_last_error(void);
/* ABI version of this header (bumped on any signature change). */
int r2d2_abi_version(void);
/* Compute capability check: 0 when the current device can run the sm_100a kernels. */
int r2d2_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * K3  GPU sum tree  <->  priority_tree.PriorityTree (priority_tree.py:4-45)
 * The tree is the reference's array-backed binary tree of float64 partial sums
 * (2^L - 1 nodes, leaves at 2^(L-1) - 1) resident in HBM.
 * ---------------------------------------------------------------------------------------- */
typedef struct r2d2_tree r2d2_tree;

/* PriorityTree.__init__ (priority_tree.py:5-13). */
int r2d2_tree_create(int64_t capacity, double prio_exponent, double is_exponent, r2d2_tree** out);
int r2d2_tree_destroy(r2d2_tree* t);
int r2d2_tree_num_layers(const r2d2_tree* t);
int64_t r2d2_tree_num_nodes(const r2d2_tree* t);
/* device pointer to the float64 node array (`PriorityTree.ptree`). */
double* r2d2_tree_nodes(r2d2_tree* t);

/* PriorityTree.update (priority_tree.py:15-24) fused with the stale-index mask of
 * ReplayBuffer.update_priorities (worker.py:247-256).
 *   idx[n] int64 leaf-relative slots, td[n] float32.  leaf <- (float)td^alpha evaluated in
 *   float32 like NumPy does for an f32 array, stored as float64; every ancestor is recomputed
 *   from its two children in float64.  Duplicate slots: the LAST occurrence wins.
 *   Masking: pass old_ptr < 0 to disable.  Otherwise slots whose block
 *   (slot / seq_per_block) was overwritten between old_ptr and cur_ptr are skipped. */
int r2d2_tree_update(r2d2_tree* t, const int64_t* idx, const float* td, int64_t n,
                     int64_t old_ptr, int64_t cur_ptr, int64_t seq_per_block, void* stream);
/* Write already-exponentiated float64 leaf values (test/restore path; same ancestor rebuild). */
int r2d2_tree_set_leaves(r2d2_tree* t, const int64_t* idx, const double* leaf, int64_t n, void* stream);

/* PriorityTree.sample (priority_tree.py:26-45): stratified prefix-sum descent in float64.
 *   unit_uniforms[n] float64 in [0,1): when NULL they are drawn on the device from a
 *   Philox4x32-10 stream keyed by (seed, call counter).
 *   idx_out[n] int64 leaf-relative; isw_out_f32[n] = (p/min p)^-beta as float32 (the cast of
 *   worker.py:234); isw_out_f64 optional (may be NULL). */
int r2d2_tree_sample(r2d2_tree* t, int64_t n, const double* unit_uniforms, uint64_t seed,
                     int64_t* idx_out, float* isw_out_f32, double* isw_out_f64, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  fused TD kernel  <->  Learner.run body worker.py:346-359 + calculate_mixed_td_errors
 * (worker.py:268-276) + value_rescale / inverse_value_rescale (worker.py:383-390).
 *   q, qn_online, qn_target: float32 [rows, A] sequence-major rows; action u8[rows];
 *   n_step_reward, n_step_gamma, is_weights float32[rows]; learning_steps u8[B].
 *   Outputs: td[rows], priorities[B], loss_sum[1] (sum of is_w*(q_a-target)^2 over rows, NOT yet
 *   divided by rows), rows_out[1] int32 (= sum learning_steps), dq[rows, A] = d(loss_sum)/dq
 *   (the caller scales by 1/rows, possibly after a cross-rank reduction). dq may be NULL.
 * ---------------------------------------------------------------------------------------- */
int r2d2_td_loss(const float* q, const float* qn_online, const float* qn_target, const uint8_t* action,
                 const float* n_step_reward, const float* n_step_gamma, const float* is_weights,
                 const uint8_t* learning_steps, int B, int A, float* td_out, float* prio_out,
                 float* loss_sum_out, int32_t* rows_out, float* dq_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1 / K1b  sequence-unroll network  <->  model.Network (model.py:27-150) as used by the learner
 * (worker.py:346,347,352) and loss.backward() (worker.py:363).
 * Parameters live in ONE caller-owned flat float32 buffer holding the reference's 20 state_dict
 * tensors (model.py:39-63) in state_dict order, each in its PyTorch layout, each start 16-byte
 * aligned; r2d2_net_param_layout returns the 21 offsets (in floats; [20] = total length).
 * ---------------------------------------------------------------------------------------- */
typedef struct r2d2_net r2d2_net;

int r2d2_net_param_layout(int action_dim, int in_channels, int64_t* offsets_out /* [21] */);
/* Workspace for a fixed batch shape: B sequences of T frames (C,84,84), at most Lmax learning
 * steps per sequence, max_forward = config.forward_steps (model.py:37). */
int r2d2_net_create(int B, int T, int C, int action_dim, int Lmax, int max_forward, r2d2_net** out);
int r2d2_net_destroy(r2d2_net* n);
/* Row capacity of the Q outputs ([rows_capacity][A], >= B*Lmax). */
int r2d2_net_rows_capacity(const r2d2_net* n);
int r2d2_net_ku(const r2d2_net* n);
/* Device pointer of the frame staging buffer conv1 reads: bf16 [B*T][21][21][16*C] (frames after space-to-depth by 4). */
void* r2d2_net_s2d_buffer(r2d2_net* n);
/* Two staging buffers so that the gather of batch i+1 can run while update i still reads batch i (the reference keeps its
 * batches prefetched in a queue, worker.py:124-139,309-316): _at(idx) is the address to gather into (buffer 1 is allocated on
 * first use, not during a stream capture), select makes idx the one the following forward / backward calls read. */
void* r2d2_net_s2d_buffer_at(r2d2_net* n, int idx);
int r2d2_net_select_s2d(r2d2_net* n, int idx);
/* cudaEventRecord for an event that a stream OUTSIDE a CUDA graph waits on: an external event-record node while `stream` is
 * being captured, a plain record otherwise. */
int r2d2_event_record(void* cuda_event, void* stream);
/* One-warp kernel on `stream` that returns once the backward recurrence of the update it is paired with is executing (or
 * after ~1 s; placement only): work enqueued behind it lands on the 84 SMs that kernel leaves idle (see r2d2_replay_set_copy_smem).
 * Pairing: one gate per update; _reset (stream-ordered, before the first update of such a pipeline) re-bases the count. */
int r2d2_net_shadow_gate(r2d2_net* n, void* stream);
int r2d2_net_shadow_gate_reset(r2d2_net* n, void* stream);
/* Re-lay out `params` for slot `which` (0 = online, 1 = target).  Call after every change of that
 * slot's parameters (optimizer step, target sync; worker.py:365,376-377). */
int r2d2_net_pack(r2d2_net* n, int which, const float* params, void* stream);
/* Network.calculate_q_ / calculate_q (model.py:81-150) in one unroll of b+l+f steps.
 *   obs u8 [B][T][C][84][84] (raw frames; the /255 of worker.py:342 is folded in);
 *   last_action u8/bool [B][T][A]; last_reward f32 [B][T]; hidden f32 [B][2][512] (Block.hidden
 *   rows, worker.py:198: [b][0] = h0, [b][1] = c0); burn/learn/fwd u8 [B] (worker.py:229-231).
 *   q_learn_out [rows_capacity][A]: Q(h_{b+t}) rows (calculate_q), may be NULL.
 *   q_shift_out [rows_capacity][A]: Q(h_{min(b+F+t, b+l+f-1)}) rows (calculate_q_), may be NULL.
 *   Rows are sequence-major; the first sum(learn) are valid.  For slot 0 the activations are kept
 *   for r2d2_net_backward; obs/hidden must stay alive until then. */
int r2d2_net_forward(r2d2_net* n, int which, const float* params, const uint8_t* obs, const uint8_t* last_action,
                     const float* last_reward, const float* hidden, const uint8_t* burn, const uint8_t* learn,
                     const uint8_t* fwd, float* q_learn_out, float* q_shift_out, void* stream);
/* cuda_event: a cudaEvent_t (or NULL to clear).  r2d2_net_backward records it on its stream once the gradients of
 * feature.7.weight and every later tensor of the flat layout (FC, LSTM, heads: 98 % of the bytes) are final, before the
 * conv layers' backward; a data-parallel learner waits on it from a side stream to overlap the gradient all-reduce
 * (the one exchange step this path adds to worker.py:362-365) with the remaining backward kernels. */
int r2d2_net_set_dense_grads_event(r2d2_net* n, void* cuda_event);
/* (h, c) of slot `which` after time step t of the last r2d2_net_forward*, as f32 [B][2][512] (the layout of `hidden`).
 * Replaces the state Network.forward returns to an actor (model.py:65-79, worker.py:533-541): on a T = 1 net,
 * r2d2_net_forward + r2d2_net_state_after is one batched environment step of B actors. */
int r2d2_net_state_after(r2d2_net* n, int which, int t, float* hidden_out, void* stream);
/* The three Q tensors of one learner update (worker.py:346,347,352) in one call: both slots are unrolled on the
 * same batch and the two recurrences advance together in shared launches. */
/* obs may be NULL when the frames were already staged by r2d2_replay_gather_s2d. */
int r2d2_net_forward_pair(r2d2_net* n, const float* params_online, const float* params_target, const uint8_t* obs,
                          const uint8_t* last_action, const float* last_reward, const float* hidden, const uint8_t* burn,
                          const uint8_t* learn, const uint8_t* fwd, float* q_learn_out, float* qn_online_out,
                          float* qn_target_out, void* stream);
/* loss.backward() (worker.py:363) for the online slot: BPTT through all b+l steps (burn-in
 * included) and the encoder.  dq [rows_capacity][A]; grads: flat buffer in the parameter layout,
 * fully overwritten (alignment gaps are left untouched and must be zero). */
int r2d2_net_backward(r2d2_net* n, const float* params, const float* dq, float* grads, void* stream);
/* GEMM scheduling of the plain-matrix contractions of K1/K1b (FC layer, LSTM input projection, their data and weight
 * gradients): 1 (default) = CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles, TMA-fed: csrc/umma3.cuh); 0 = single-CTA
 * 128 x 128 tiles fed by cp.async (csrc/umma2.cuh).  Same results to fp32 rounding.  Returns the previous value. */
int r2d2_set_pair_gemm(int on);
/* Incremented by every r2d2_set_* call: a caller that caches launch sequences (CUDA graphs of an update) keys them on it. */
int r2d2_config_epoch(void);
/* Forward recurrence: 1 (default) = one thread-block cluster of 16 CTAs per (network, 16 sequences): W_hh resident in tensor
 * memory + shared memory, h_t exchanged over distributed shared memory (csrc/recurrence2.cuh); 0 = the persistent kernel
 * that exchanges h_t through L2 flags (csrc/recurrence.cuh).  Returns the previous value. */
int r2d2_set_cluster_recurrence(int on);
/* Diagnostics: how many such clusters the current device can keep resident at once (< 0: the query failed). */
int r2d2_debug_cluster_capacity(void);
/* Recurrence scheduling: 1 (default) = all T LSTM steps of both networks in ONE persistent cooperative kernel
 * (W_hh slices resident in shared memory, per-network step barriers) for B <= 64; 0 = one launch per step. */
int r2d2_set_persistent_recurrence(int on);
/* Debug: device buffer of T*8 uint64 that receives per-step globaltimer stamps of CTA 0 of the persistent recurrence
 * (poll start, flag acquired, tile staged, MMAs done, epilogue done, released); NULL detaches. */
int r2d2_debug_rec_trace(void* device_buffer);
/* The same for the cluster BPTT kernel (stamps: step start, partials received, dgates staged, barrier passed, accumulator
 * ready, partial pushed). */
int r2d2_debug_rec_trace_bwd(void* device_buffer);
/* Test/debug access to device intermediates (see net.cu for the names). */
void* r2d2_net_debug_ptr(r2d2_net* n, int which, const char* name);

/* ------------------------------------------------------------------------------------------
 * K4  HBM-resident replay block store  <->  ReplayBuffer.add / sample_batch storage halves
 * (worker.py:141-161, 163-240) and the Block wire format (worker.py:23-35).
 * ---------------------------------------------------------------------------------------- */
typedef struct r2d2_replay r2d2_replay;

/* Ring of num_blocks slots (worker.py:43-48,70); each slot holds one packed Block at fixed byte offsets. */
int r2d2_replay_create(int num_blocks, int block_len, int burn_in, int learning, int forward, int C, int action_dim, int H,
                       r2d2_replay** out);
int r2d2_replay_destroy(r2d2_replay* r);
/* offsets_out[12] (bytes inside a slot): obs, last_action, last_reward, action, n_step_reward, gamma, hidden,
 * burn, learn, fwd, num_seq, total.  Array shapes: obs u8 [burn_in+block_len+1][C][84][84]; last_action u8
 * [frames][A] (one-hot); last_reward f32 [frames]; action u8 [block_len]; n_step_reward, gamma f32 [block_len];
 * hidden f32 [block_len/learning][2][H]; burn/learn/fwd u8 [block_len/learning]; num_seq i32. */
int r2d2_replay_layout(const r2d2_replay* r, int64_t* offsets_out);
/* ReplayBuffer.add, storage half (worker.py:154): async H2D copy of one packed block from PINNED host memory. */
int r2d2_replay_ingest(r2d2_replay* r, int block_idx, const void* host_blob, int64_t nbytes, void* stream);
/* ReplayBuffer.sample_batch, slicing half (worker.py:172-238), for B sampled slots idx (device int64) with their
 * IS weights isw (device f32).  Outputs are the device-side 14-tuple fields (see r2d2_net_forward for layouts);
 * is_weights_rows holds each sequence's weight repeated over its learning steps (worker.py:216); rows_out i32[1]. */
int r2d2_replay_gather(r2d2_replay* r, const int64_t* idx, const float* isw, int B, int T, uint8_t* obs, uint8_t* last_action,
                       float* last_reward, float* hidden, uint8_t* action, float* n_step_reward, float* gamma, uint8_t* burn,
                       uint8_t* learn, uint8_t* fwd, float* is_weights_rows, int32_t* rows_out, void* stream);

/* bytes of unused dynamic shared memory per copy CTA of the gathers (0 = default): with 32 KB the CTAs cannot share an SM with
 * a ~200 KB GEMM / recurrence CTA, so a gather that runs next to an update only takes idle SMs. */
int r2d2_replay_set_copy_smem(r2d2_replay* r, int bytes);
/* The same gather writing the frames straight into the network's space-to-depth bf16 staging buffer
 * (s2d_out = r2d2_net_s2d_buffer(net) or _at(net, idx)); r2d2_net_forward_pair is then called with obs == NULL. */
int r2d2_replay_gather_s2d(r2d2_replay* r, const int64_t* idx, const float* isw, int B, int T, void* s2d_out, uint8_t* last_action,
                           float* last_reward, float* hidden, uint8_t* action, float* n_step_reward, float* gamma, uint8_t* burn,
                           uint8_t* learn, uint8_t* fwd, float* is_weights_rows, int32_t* rows_out, void* stream);

/* Precision mode of the tensor-core path: 0 = strict (bf16x3 split products everywhere, default), 1 = fast (plain
 * bf16 products), 2 = balanced (hi+lo only for the weight operands of the encoder contractions; recurrence, input
 * projection and dueling head stay strict).  Returns the previous mode. */
int r2d2_set_fast_math(int mode);
/* Test entry for the v2 kernel: operands as bf16 hi/lo planes; major 1 = [K][rows] storage (MN-major descriptors). */
int r2d2_debug_gemm2(int ubn, int a_major, int b_major, int M, int N, int K, const void* a_hi, const void* a_lo,
                     const void* b_hi, const void* b_lo, float* C, int splits, void* stream);

/* Test entry for the v3 kernel (CTA pairs with cta_group::2 MMAs, TMA-fed): same operand convention; C [splits][M][N] must be zeroed. */
int r2d2_debug_gemm3(int a_major, int b_major, int M, int N, int K, const void* a_hi, const void* a_lo, const void* b_hi,
                     const void* b_lo, float* C, int splits, void* stream);

/* Hardware probe (tests/tools only): D[128][32] = A[shift .. shift+128)[64] . B[32][64]^T with a K-major
 * SWIZZLE_128B descriptor whose start address is shifted by `shift` rows inside one staged buffer (A bf16 [144][64]);
 * mode 1 also sets the descriptor's base_offset field to shift & 7. */
int r2d2_debug_shift_probe(const void* A, const void* B, float* D, int shift, int mode, void* stream);
/* hardware probe: D[128][16] = A[128][64] . B[16][64]^T (bf16) with A read from tensor memory (tcgen05.st) and B from
 * K-major SWIZZLE_64B shared-memory tiles -- the operand forms of the cluster recurrence (csrc/recurrence2.cuh). */
int r2d2_debug_ts_probe(const void* A, const void* B, float* D, void* stream);
/* hardware probe: cycles for reps*4 back-to-back tcgen05.mma (M x N x 16, bf16, operands in shared memory) on each of
 * `ctas` CTAs; mode bit 0 alternates two accumulators, bit 1 reads A MN-major.  cycles[0] <- clock64 delta of CTA 0. */
int r2d2_debug_mma_rate(int M, int N, int reps, int mode, int ctas, long long* cycles, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5  clip_grad_norm_(max_norm) + Adam(lr, eps).step()  (worker.py:289,364-365) on the flat
 * buffers.  grad_scale: optional device float multiplied into the gradients first;
 * partial_ws: device double[592] scratch; step: 1-based update count; norm_out: optional device
 * float receiving the pre-clip global norm.
 * ---------------------------------------------------------------------------------------- */
int r2d2_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                   const float* grad_scale, double* partial_ws, float max_norm, float lr, float beta1, float beta2,
                   float eps, int64_t step, float* norm_out, void* stream);
/* Same update with nothing host-variable in the argument list (CUDA-graph replay): the 1-based update count is read from
 * device memory (step_dev, int64, incremented by the caller on the same stream) and the gradient scale may come from a
 * device row count (rows_dev, int32: scale = 1 / rows, overrides grad_scale; NULL = use grad_scale). */
int r2d2_clip_adam_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const float* grad_scale, const int32_t* rows_dev, double* partial_ws, float max_norm, float lr, float beta1,
                       float beta2, float eps, const int64_t* step_dev, float* norm_out, void* stream);
/* The same with the increment of the update count done on the device as well (step_dev += 1 before the Adam kernel reads it). */
int r2d2_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* grad_scale, const int32_t* rows_dev, double* partial_ws, float max_norm, float lr, float beta1,
                        float beta2, float eps, int64_t* step_dev, float* norm_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel exchange step over NVLink peer memory (the reference has ONE learner,
 * worker.py:363-365 is a local backward + optimizer step, so these have no upstream counterpart:
 * they sit between r2d2_net_backward and r2d2_clip_adam_dev when there is one learner per GPU).
 * Every rank maps every rank's gradient buffer and a small control block (symmetric memory:
 * grad_ptrs[i] / ctl_ptrs[i] are THIS process's addresses of rank i's copies; mc_grad_ptr is the
 * NVSwitch multicast address of the gradient buffers or 0).  Control blocks are
 * r2d2_dp_ctl_bytes() bytes, zeroed on every rank before the first call.
 *  r2d2_dp_allreduce   in-place SUM over ranks of grads[off, off+len) (floats, multiples of 4).
 *      Rank r reduces slice r (multimem.ld_reduce in the switch, or P2P loads) and writes it to
 *      all ranks.  rows_dev / rows_slot / grad_scale_dev (all or none): the local int32 row count
 *      is written to grads[rows_slot] (a padding element inside the range) before the reduction;
 *      afterwards grad_scale = 1 / (global rows) and the slot is zero again.  channel 0/1: two
 *      reductions may be in flight on different streams.  ctas x threads: launch shape (threads a
 *      multiple of 32, <= 512); no shared memory, so the CTAs fit next to resident GEMM CTAs.
 *  r2d2_dp_is_post / r2d2_dp_is_apply   importance weights of one global prioritized sampler
 *      (priority_tree.py:39-41 over all shards).  post, right after sampling: min over the sampled
 *      leaves / root from the tree's node array -> every peer.  apply, before K2 of the same update:
 *      is_weights[0, rows) *= ((min_local/root_local) / min over ranks)^-beta; factor_out optional.
 *      Exactly one post and one apply per update; they may run on different streams (apply waits
 *      for every rank's flag, this rank's included).
 * All calls are stream-ordered and graph-replayable; a rank that waits ~60 s for a peer traps
 * (r2d2_dp_error then reports which barrier).
 * ---------------------------------------------------------------------------------------- */
size_t r2d2_dp_ctl_bytes(void);
int r2d2_dp_create(int rank, int world, const unsigned long long* grad_ptrs, unsigned long long mc_grad_ptr,
                   const unsigned long long* ctl_ptrs, void** handle);
void r2d2_dp_destroy(void* handle);
int r2d2_dp_allreduce(void* handle, long long off, long long len, int channel, const int32_t* rows_dev, long long rows_slot,
                      float* grad_scale_dev, int ctas, int threads, int use_multicast, void* stream);
int r2d2_dp_is_post(void* handle, const double* nodes, long long leaf_base, const long long* idx, int n, void* stream);
int r2d2_dp_is_apply(void* handle, double beta, float* is_weights, int rows, float* factor_out, void* stream);
int r2d2_dp_error(void* handle, unsigned int* out);

#ifdef __cplusplus
}
#endif
#endif /* R2D2_B200_H */
