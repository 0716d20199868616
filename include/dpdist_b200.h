/*
 * dpdist_b200 -- C ABI of the B200-native DPDist hot path (libdpdist_b200.so).
 *
 * The reference (dahliau/DPDist) has no FFI: the path lives behind Python functions that build a
 * TF1 graph.  This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md); each entry point names the reference code it replaces, paths relative to the
 * reference repo root.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (16-byte aligned, contiguous,
 *     row-major); h_* is HOST memory.  The library never allocates, frees or retains them.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden syncs.
 *   - return 0 = ok, <0 = invalid argument / unsupported configuration (DPD_E_*), >0 = cudaError_t.
 *     dpd_last_error() returns a thread-local message for the last non-zero return.
 *   - "cloud" = one point set; a DPDist pair (A,B) is two clouds.  The host concatenates
 *     [A-clouds | B-clouds] on the leading axis, mirroring tf.concat([net, netB], 0)
 *     (utils/dpdist_util.py:511).
 *   - grid tables are built on the HOST exactly as the reference builds them (numpy fp64 -> fp32):
 *     h_centers[G] = float32(np.linspace(-1,1,G,False)+1/G)  (utils/dpdist_util.py:42,50)
 *     h_lo[G], h_hi[G] = float32(c) -/+ float32(gs), c from get_grid_centers (:982-992),
 *     gs = |C[0].z-C[1].z|/2 (:468), so the float comparisons are the reference's own.
 *   - flat Gaussian / voxel index g = i0*G*G + i1*G + i2 has centre (x=l[i1], y=l[i0], z=l[i2])
 *     (np.meshgrid 'xy' indexing, utils/dpdist_util.py:47-48, 990-992).
 */
#ifndef DPDIST_B200_H
#define DPDIST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPD_ABI_VERSION 4
#define DPD_MAX_GRID 16
#define DPD_FV_CHANNELS_FULL 20
#define DPD_FV_CHANNELS_SMALL 7

#define DPD_E_INVALID (-1)     /* null pointer, non-positive size, misalignment */
#define DPD_E_UNSUPPORTED (-2) /* configuration outside what the kernels implement */
#define DPD_E_WORKSPACE (-3)   /* workspace too small */

/* head implementation selector (flags argument of dpd_head_forward) */
#define DPD_HEAD_AUTO 0
#define DPD_HEAD_SIMT 1 /* fp32 FFMA GEMMs (sanity path) */
#define DPD_HEAD_TC 2      /* tcgen05 tensor-core GEMMs, fp16x3 split precision with power-of-two scaling */
#define DPD_HEAD_TC_TF32 3 /* tcgen05 tensor-core GEMMs, 3xTF32 split precision */
#define DPD_HEAD_TRAIN 0x10 /* OR-ed into flags: keep the activations for dpd_head_backward (one row chunk only) */
#define DPD_HEAD_INPUT_GRAD 0x20 /* OR-ed with DPD_HEAD_TRAIN: also keep what dpd_head_backward_inputs needs */

/* stages of dpd_head_backward: ALL, or one layer at a time (4 -> 1) so that the caller can start the
 * gradient all-reduce of a layer while the next one is still being computed */
#define DPD_BWD_ALL 0
#define DPD_BWD_L4 1
#define DPD_BWD_L3 2
#define DPD_BWD_L2 3
#define DPD_BWD_L1 4

int dpd_version(void);
const char* dpd_last_error(void);

/* 3DmFV encoding: replaces get_3dmfv_tf (utils/dpdist_util.py:22-141; second copy
 * pcrnet-registration/models/ipcr_model.py:53-172).
 *   d_points [n_clouds, n_points, 3] fp32
 *   d_fv     flatten=0: [n_clouds, G^3, C]   (C = 20 if full_fv else 7; order pi(mean[,max]),
 *                        mu(mean xyz[,max xyz,min xyz]), sigma(same))          (:134-137)
 *            flatten=1: [n_clouds, C*G^3]    channel-major                      (:129-132)     */
int dpd_fv_forward(const float* d_points, int n_clouds, int n_points, int G, const float* h_centers,
                   float sigma, int full_fv, int flatten, float* d_fv, void* stream);

/* Gradient of the 3DmFV encoding w.r.t. the points: replaces what tf.gradients builds over get_3dmfv_tf
 * (utils/dpdist_util.py:54-137) when the DPDist graph is used as a loss for another network and gradients flow
 * into input1 / input2 (pcrnet-registration/iterative_PCRNet_ours.py:229-257; train_multi_gpu_pc_compare_dist.py:433-463).
 * TF semantics: reduce_max / reduce_min gradients are split evenly among ties; d/dx sign(x) sqrt(max(|x|,1e-12)) is 0
 * below the clamp; l2_normalize is a constant scale below its clamp.  G <= 10 (shared-memory limit).
 *   d_grad_fv     same layout as dpd_fv_forward's d_fv (flatten = 0 or 1)
 *   d_grad_points [n_clouds, n_points, 3] out                                                    */
int dpd_fv_backward(const float* d_points, int n_clouds, int n_points, int G, const float* h_centers,
                    float sigma, int full_fv, int flatten, const float* d_grad_fv, float* d_grad_points,
                    void* stream);

/* Voxel assignment of query points: replaces DPDist.get_pc_grid_binary_mask_from_centers +
 * the mask / offset gathers of get_emb_and_concat (utils/dpdist_util.py:459-492, 434-447).
 *   d_query  [n_clouds, n_query, 3] fp32
 *   d_idx    [n_clouds, n_query]    int32 flat voxel index (first match; 0 if none)   (:490)
 *   d_mask   [n_clouds, n_query]    fp32 1/0 in-cube flag (binary_vect at idx)        (:436-440)
 *   d_offset [n_clouds, n_query, 3] fp32 query - centre(idx)                          (:443-447, 491)
 * any output pointer may be NULL.                                                               */
int dpd_voxel_assign(const float* d_query, int n_clouds, int n_query, int G, const float* h_centers,
                     const float* h_lo, const float* h_hi, int32_t* d_idx, float* d_mask,
                     float* d_offset, void* stream);

/* Materialised local patches: replaces local_z / local_z_3d (utils/dpdist_util.py:850-854,
 * 911-960).  Only for callers that ask for embedding_set; the head never materialises this.
 *   d_fv [n_clouds, G^3, C] -> d_patches [n_clouds, G^3, k^3*C], element order (a0,a1,a2,c),
 *   zero outside the grid, SAME padding ((k-1)/2 before, k/2 after).                           */
int dpd_local_patches(const float* d_fv, int n_clouds, int G, int C, int k, float* d_patches,
                      void* stream);

/* Implicit distance head: replaces DPDist conv_version 1 (utils/dpdist_util.py:412-544, 688-700)
 * + tf_util.conv2d x4 (utils/tf_util.py:161-228), evaluated per cloud:
 *   out[c, q, :] = mask * relu6(MLP([query - centre(idx) | patch_k(fv[c], idx)])) / 3
 * Weights are given in the reference's own layouts (HWIO flattened, row-major [K_in, K_out]):
 *   d_w1 [3 + k^3*C, H] (offset rows FIRST, :455), d_w2 [H,H], d_w3 [H,H], d_w4 [H,3], biases [H],[H],[H],[3].
 * dpd_head_pack_weights re-lays them out for the kernels (call again whenever they change).      */
typedef struct dpd_head_config {
  int n_clouds; /* 2*B for a DPDist batch of B pairs */
  int n_query;  /* queries per cloud (NP) */
  int G;        /* voxels per axis */
  int C;        /* FV channels per voxel (20) */
  int k;        /* local patch edge */
  int H;        /* MLP width (1024) */
  int flags;    /* DPD_HEAD_* */
} dpd_head_config;

size_t dpd_head_packed_bytes(const dpd_head_config* cfg);
size_t dpd_head_workspace_bytes(const dpd_head_config* cfg);

int dpd_head_pack_weights(const dpd_head_config* cfg, const float* d_w1, const float* d_b1,
                          const float* d_w2, const float* d_b2, const float* d_w3,
                          const float* d_b3, const float* d_w4, const float* d_b4, void* d_packed,
                          void* stream);

/*   d_fv    [n_clouds, G^3, C]      fp32 (output of dpd_fv_forward, flatten=0)
 *   d_query [n_clouds, n_query, 3]  fp32
 *   d_out   [n_clouds, n_query, 3]  fp32
 *   d_idx   optional [n_clouds, n_query] int32 out (voxel index actually used)                  */
int dpd_head_forward(const dpd_head_config* cfg, const float* d_fv, const float* d_query,
                     const float* h_centers, const float* h_lo, const float* h_hi,
                     const void* d_packed, float* d_out, int32_t* d_idx, void* d_workspace,
                     size_t workspace_bytes, void* stream);

/* Whole forward of the hot path in one call: replaces the body of get_model for k > 0, conv_version 1
 * (models/dpdist_and_aue.py:31-86): get_3dmfv_tf of every cloud (:56-61) -> local_z (:64-65, never
 * materialised) -> DPDist (:69-75).  Same results as dpd_fv_forward followed by dpd_head_forward; the
 * library additionally lets the 3DmFV kernel emit the tensor-core operand copy of the FV tensor and uses
 * the bound |fv| <= 1 (per-channel L2 normalisation, utils/dpdist_util.py:124-126) instead of measuring it.
 *   d_points [n_clouds, n_points, 3]  clouds to encode, rows [A-clouds (+noise, :45) | B-clouds]
 *   d_query  [n_clouds, n_query, 3]   rows [pcB | pcA]: A's field is queried at B's points (:494-500)
 *   d_fv     [n_clouds, G^3, C]  out  the 3DmFV tensor (what embedding_set is built from), C = cfg->C in {20, 7}
 *   d_out    [n_clouds, n_query, 3] out  rows [pred_AB | pred_BA]
 *   h_fv_centers[G]: Gaussian axis centres (:42); h_centers / h_lo / h_hi: voxel tables (see top).     */
int dpd_model_forward(const dpd_head_config* cfg, const float* d_points, int n_points, float sigma,
                      const float* h_fv_centers, const float* d_query, const float* h_centers,
                      const float* h_lo, const float* h_hi, const void* d_packed, float* d_fv,
                      float* d_out, int32_t* d_idx, void* d_workspace, size_t workspace_bytes,
                      void* stream);

/* Gradients of the head's 8 variables: replaces what optimizer.compute_gradients(total_loss_samples,
 * vars in scope 'pc_compare') builds in the reference trainer (train_multi_gpu_pc_compare_dist.py:274-277)
 * over utils/dpdist_util.py:494-544,688-698.  Must follow a dpd_head_forward with the SAME cfg (flags
 * including DPD_HEAD_TRAIN), packed weights and workspace; uses the activations left there.
 *   d_grad_out [n_clouds, n_query, 3]  dLoss/d(out); row blocks that are identically zero are skipped
 *   d_gw1 [3+k^3*C, H] (reference row order, offset rows first), d_gw2/3 [H,H], d_gw4 [H,3], d_gb* biases
 * stage = DPD_BWD_ALL, or DPD_BWD_L4, L3, L2, L1 in that order (each writes only its layer's gradients).
 * A NULL d_gw<i> skips that layer's weight / bias gradient (frozen DPDist used as a loss): the chain of
 * activation gradients is still propagated for dpd_head_backward_inputs.
 * With the fp16x3 tensor-core head (the default) every product of the backward pass runs on the tensor cores as well
 * (dX = dZ.W^T with the ReLU' bit masks the forward left in the workspace, dW = A^T.dZ split over the active rows);
 * the stages must then be issued in order starting with DPD_BWD_L4 (it resets the per-layer |dZ|max slots).
 * Deterministic: all reductions run in a fixed order.                                                      */
int dpd_head_backward(const dpd_head_config* cfg, const float* d_fv, const void* d_packed,
                      const float* d_grad_out, int stage, float* d_gw1, float* d_gb1, float* d_gw2,
                      float* d_gb2, float* d_gw3, float* d_gb3, float* d_gw4, float* d_gb4,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* Gradients of the head w.r.t. its INPUTS: the 3DmFV tensor and the query points (the path PCRNet-ours and the AUE
 * task differentiate through, pcrnet-registration/iterative_PCRNet_ours.py:229-257).  Must follow
 * dpd_head_backward stages L4, L3, L2 (or ALL) of the same forward; cfg.flags carry DPD_HEAD_TRAIN | DPD_HEAD_INPUT_GRAD.
 *   d_grad_fv    [n_clouds, G^3, C]     out: sum over the cloud's queries of the patch part of dL/d(layer-1 input),
 *                                       i.e. the transpose of the gather of get_emb_and_concat (:449-453) and of
 *                                       extract_volume_patches (:922-930)
 *   d_grad_query [n_clouds, n_query, 3] out: the offset part (offset = query - centre, :491; the voxel index and the
 *                                       in-cube mask are piecewise constant and carry no gradient)            */
int dpd_head_backward_inputs(const dpd_head_config* cfg, const void* d_packed, float* d_grad_fv,
                             float* d_grad_query, void* d_workspace, size_t workspace_bytes, void* stream);

/* Adam with tf.train.AdamOptimizer semantics (train_multi_gpu_pc_compare_dist.py:216, 301):
 * lr_t = lr*sqrt(1-beta2^step)/(1-beta1^step); var -= lr_t * m / (sqrt(v) + eps).  step is 1-based. */
int dpd_adam_step(float* d_param, const float* d_grad, float* d_m, float* d_v, size_t n, float lr,
                  float beta1, float beta2, float eps, int step, void* stream);

/* The same update with the bias-corrected rate lr_t = dpd_adam_lr_t(lr, beta1, beta2, step) read from DEVICE memory:
 * lets a captured CUDA graph of the whole training step (forward, backward, Adam) be replayed while the host only
 * rewrites that scalar -- the reference's batch of 16 pairs (train_multi_gpu_pc_compare_dist.py:57) is launch-bound. */
float dpd_adam_lr_t(float lr, float beta1, float beta2, int step);
int dpd_adam_step_dev(float* d_param, const float* d_grad, float* d_m, float* d_v, size_t n, const float* d_lr_t,
                      float beta1, float beta2, float eps, void* stream);

/* Test hook for the tensor-core GEMM used by layers 2-3 of the head:
 *   d_out[M,N] = relu(d_a[M,K] . d_w[K,N] + d_bias[N]),  K % 64 == 0, N % 256 == 0,
 * computed with the split-precision tcgen05 kernel (f16 != 0: fp16x3, else 3xTF32).
 * d_scratch >= 8*(M*K + N*K) + 256 bytes.                                                      */
int dpd_debug_tc_gemm(const float* d_a, int M, int K, const float* d_w, int N, const float* d_bias,
                      float* d_out, void* d_scratch, size_t scratch_bytes, int f16, void* stream);

/* Test hook (host only, no launch): the physical order of the layer-1 operand row in the fp16 tensor-core path.
 *   h_out[k] = position in the reference's row [patch (taps*C, tap-major) | offsets (3) | padding] of the element the
 * packed row of length Kp (a multiple of 64) holds at k.  The reference concatenates [offset | patch]
 * (utils/dpdist_util.py:455); the packed order is patch-first, channel-split (16 + 4 channels at C = 20) with its
 * 64-element blocks permuted; W1 is packed and the dW1 partials are mapped back with this map. */
int dpd_debug_tc_operand_order(int taps, int C, int Kp, int* h_out);

/* ---- layer-by-layer fp32 path for training-mode batch norm (--BN 1) --------------------------------------------------
 * Replaces, for `bn` truthy and is_training True, what utils/tf_util.py:213-227 builds per conv layer: conv2d + bias_add
 * (dpd_layer_forward), tf.contrib.layers.batch_norm with batch statistics (utils/tf_util.py:558-577; dpd_bn_forward) and
 * the activation, plus the gradient graph over them (dpd_bn_backward, dpd_layer_backward).  Batch statistics need a whole
 * layer's pre-activations before it can be normalised, so this configuration cannot use the fused head; it is off at the
 * reference defaults (train_multi_gpu_pc_compare_dist.py:61,105).  All arrays fp32, row-major, device memory.
 *
 * dpd_layer_forward:  z[rows,N] = act(x[rows,K] . w[K,N] + b[N]), act 0 = none, 1 = relu.
 *   d_fv != NULL selects the gathered layer 1: x is virtual, row r = [patch_k(fv[cloud(r)], idx[r]) | offset[r] | 0 pad],
 *   cloud(r) = r / n_query, and w must be packed the same way (patch rows, 3 offset rows, zero rows; K = rows of w, a
 *   multiple of 16, >= k^3*C + 3); d_x is ignored.  N <= 4 (the output layer) uses a narrow kernel, otherwise N % 4 == 0
 *   and K % 16 == 0.
 * dpd_layer_backward: gw[K',N] = x^T . dz, gb[N] = column sums of dz, dx[rows,K] = dz . w^T (d_dx may be NULL).  For the
 *   gathered layer d_dx is the gradient w.r.t. fv, [rows / n_query, G^3, C] (the transpose of the patch gather, without
 *   atomics); the workspace must then hold, beyond dpd_layer_workspace_bytes, rows*12 bytes and at least one cloud's
 *   n_query * K * 4 bytes (more = fewer passes).  For the gathered layer gw is returned in the REFERENCE row order (offset first,
 *   utils/dpdist_util.py:455) with K' = k^3*C + 3 rows.  Needs N % 128 == 0 unless N <= 4.
 * dpd_bn_forward:  mean[c], var[c] (biased) over the rows, y = act(gamma * (z - mean) * rsqrt(var + eps) + beta),
 *   act 0 = none, 1 = relu.  Two-pass statistics, fixed-order reductions (deterministic).
 * dpd_bn_backward: dz, dgamma, dbeta from dy (the gradient w.r.t. y; the activation's gate is re-derived from z).
 * Workspace for all four: dpd_layer_workspace_bytes(rows, K, N) bytes (use the layer's K and N; K = N for the bn calls). */
size_t dpd_layer_workspace_bytes(int rows, int K, int N);
int dpd_layer_forward(const float* d_x, int rows, int K, const float* d_w, const float* d_b, int N, int act, float* d_z,
                      const float* d_fv, const int32_t* d_idx, const float* d_offset, int n_query, int G, int C, int k,
                      void* stream);
int dpd_layer_backward(const float* d_x, int rows, int K, const float* d_w, int N, const float* d_dz, float* d_gw,
                       float* d_gb, float* d_dx, const float* d_fv, const int32_t* d_idx, const float* d_offset,
                       int n_query, int G, int C, int k, void* d_workspace, size_t workspace_bytes, void* stream);
/* Pieces the conv_version 3 head (utils/dpdist_util.py:640-687: 1x1x1 and 3x3x3 conv3d over the k^3 patch, residual blocks,
 * FC) is assembled from, together with dpd_layer_forward / dpd_layer_backward: a 3x3x3 SAME conv3d over the [k,k,k,Cin]
 * volume of every query IS the gathered layer with fv = the volumes, G = k, patch edge 3, n_query = k^3, idx = 0..k^3-1.
 *   dpd_gather_rows:   out[r,:] = [offset (3) | patch (k^3*C)], the row get_emb_and_concat builds (:434-457)
 *   dpd_relu_backward: dy[i] = 0 where y[i] <= 0 (in place)        dpd_add_inplace: a += b (residual connections) */
int dpd_gather_rows(const float* d_fv, const int32_t* d_idx, const float* d_offset, int rows, int n_query, int G, int C, int k,
                    float* d_out, void* stream);
int dpd_relu_backward(float* d_dy, const float* d_y, size_t n, void* stream);
int dpd_add_inplace(float* d_a, const float* d_b, size_t n, void* stream);
int dpd_bn_forward(const float* d_z, int rows, int N, const float* d_gamma, const float* d_beta, float eps, int act,
                   float* d_y, float* d_mean, float* d_var, void* d_workspace, size_t workspace_bytes, void* stream);
int dpd_bn_backward(const float* d_z, const float* d_dy, int rows, int N, const float* d_gamma, const float* d_beta,
                    const float* d_mean, const float* d_var, float eps, int act, float* d_dz, float* d_dgamma,
                    float* d_dbeta, void* d_workspace, size_t workspace_bytes, void* stream);

/* Ground-truth distances of the dataset generator: replaces scipy cdist(point_set, neg_set).min(0)
 * (dataset_sample_with_gt.py:87-91, 116-117), brute force in fp32 on d^2 = dx^2 + dy^2 + dz^2 (no |a|^2+|b|^2-2ab).
 *   d_surface [n_clouds, n_surface, 3], d_query [n_clouds, n_query, 3]
 *   d_dist    [n_clouds, n_query]  out: Euclidean distance to the nearest surface point
 *   d_arg     [n_clouds, n_query]  out, optional (NULL): index of that point (first minimum)             */
int dpd_nearest_distance(const float* d_surface, int n_clouds, int n_surface, const float* d_query, int n_query,
                         float* d_dist, int32_t* d_arg, void* stream);

/* Batch assembly of the DPDist trainer fused with the dataset augmentation: replaces train_one_epoch_3d's numpy
 * slicing (train_multi_gpu_pc_compare_dist.py:749-766) and ModelNetDataset._augment_batch_data
 * (modelnet_dataset.py:82-95 = provider.rotate_point_cloud :32-50 + provider.shift_point_cloud :200-211).
 *   d_data  [bsize, 3*npoints, 3] = surface | close | far points of each item (modelnet_dataset.py:136-139)
 *   d_label [bsize, 2*npoints]    = GT distances of close | far
 *   d_angle [bsize] rotation about the up (y) axis in radians, d_shift [bsize, 3]; either may be NULL (no augmentation)
 *   d_pcA [bsize, num_point, 3] = S_A[:num_point];  d_pcB [bsize, num_point, 3] = S_B[:h] | close[:q] | far[q:h];
 *   d_labels_ab [bsize, num_point] = 0 x h | gt_close[:q] | gt_far[q:h],  h = num_point/2, q = h/2,
 *   S_A / S_B = the two halves of the surface points.                                                        */
int dpd_assemble_batch(const float* d_data, const float* d_label, int bsize, int npoints, int num_point,
                       const float* d_angle, const float* d_shift, float* d_pcA, float* d_pcB,
                       float* d_labels_ab, void* stream);

/* CRC32C (Castagnoli) of a HOST buffer: the per-variable checksum of TensorFlow V2 checkpoints, used when a
 * reference-trained model.ckpt (train_multi_gpu_pc_compare_dist.py:311; consumed by
 * pcrnet-registration/iterative_PCRNet_ours.py:229) is loaded under its TF variable names. */
uint32_t dpd_crc32c(const void* h_data, size_t n);

/* Measurement hooks (used by bench.py; no effect on results).
 * dpd_launch_count : kernels this library has launched in this process (cumulative).
 * dpd_profile_enable(1) brackets every kernel launch with CUDA events on the launching stream;
 * dpd_profile_read synchronises them, writes per-kernel totals (device ms, launches) and
 * returns the number of entries written (reset != 0 clears the totals afterwards).            */
typedef struct dpd_profile_entry {
  char name[48];
  double ms;
  long long launches;
} dpd_profile_entry;

long long dpd_launch_count(void);
int dpd_profile_enable(int on);
int dpd_profile_read(dpd_profile_entry* h_entries, int max_entries, int reset);

#ifdef __cplusplus
}
#endif
#endif /* DPDIST_B200_H */
