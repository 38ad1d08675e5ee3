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
 *   (slot / seq_per_block) was overwritten between old_ptr and cur_ptr are skipped.
 *   cur_ptr_dev (optional, may be NULL): when non-NULL the current block pointer is read
 *   from this device int32 instead of cur_ptr (HBM-resident replay keeps it on device). */
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

#ifdef __cplusplus
}
#endif
#endif /* R2D2_B200_H */
