// Backward of the implicit distance head w.r.t. its 8 variables (fp32 SIMT kernels).
// Replaces what tf.gradients builds for optimizer.compute_gradients(total_loss_samples, vars in
// scope 'pc_compare') in the reference trainer (train_multi_gpu_pc_compare_dist.py:274-277).
#pragma once
#include "head_simt.cuh"

namespace dpd {

constexpr int BWD_SLICES = 8;   // split of the row (reduction) axis of the weight-gradient GEMMs

// active[b] = 1 if any of grad_out[128*b .. 128*b+127, 0..2] is non-zero.  Row blocks whose upstream
// gradient is identically zero (the whole B->A half in DPDist training, :967) are skipped everywhere.
int launch_row_active(const float* grad_out, int M, int* active, cudaStream_t st);

// dZ3[r,:] = (sum_j dz4[r,j] W4[:,j]) * (H3[r,:] > 0),   dz4 = grad_out * mask * relu6'(z4) / 3,
// gw4 = H3^T dz4, gb4 = sum_r dz4.
// Block formulation (one CTA per 64 rows, two phases): partial4 holds one (3H + 3) record per 64-row
// block; launch_reduce_out_blocks sums the records of the active blocks in a fixed order.
int launch_out_backward_blocks(const float* h3, const float* w4, const float* b4, const float* mask, const float* grad_out,
                               const int* active, float* dz3, float* partial4, int M, int H, cudaStream_t st,
                               unsigned* absmax_bits = nullptr);
int launch_reduce_out_blocks(const float* partial4, const int* active, int M, int H, float* gw4, float* gb4, cudaStream_t st);

struct TnParams {
  const float* A;      // dense [M, lda] (ignored when gathering)
  int lda;
  const float* B;      // upstream gradient dZ [M, N]
  int M, N, Kp;        // Kp = columns of A (multiple of 128 after padding for the launch grid)
  const int* active;   // per 128-row block
  float* partial;      // [BWD_SLICES][Kp][N]
  float* partial_bias; // [BWD_SLICES][N] column sums of B (written by the blockIdx.y == 0 CTAs)
  GatherDesc g;
};
int launch_simt_gemm_tn(const TnParams& p, bool gather, cudaStream_t st);

// grad[row_map(k)][n] = sum_s partial[s][k][n]; unpermute != 0 maps packed layer-1 rows (patch | offset | pad)
// back to the reference's (offset | patch) order and drops the padding rows; with tc_taps > 0 the patch part of the
// partials is in the channel-split operand order of the fp16 tensor-core gather (common.cuh tc_k_to_patch_k).
int launch_reduce_partials(const float* partial, const float* partial_bias, int Kp, int K_valid, int N, int E, int unpermute,
                           float* gw, float* gb, cudaStream_t st, int nslices = BWD_SLICES, int tc_taps = 0, int tc_C = 0);

int launch_add_inplace(float* a, const float* b, size_t n, cudaStream_t st);

// Input gradients of layer 1 for the clouds [cloud0, cloud0 + n_clouds): dx1 [rows, ldx] = dZ1 . W1p^T holds, per query row,
// the gradient of its virtual operand row [patch (E) | offset (3) | pad].  grad_fv[cloud, v, :] gathers the patch
// parts of all queries of the cloud whose k^3 neighbourhood contains v (fixed query order: deterministic, no
// atomics); grad_query[row, :] = the offset part (offset = query - centre, utils/dpdist_util.py:491).
// idx / active are indexed by global row; dx1 row 0 is global row cloud0 * n_query.
int launch_patch_scatter(const float* dx1, int ldx, const int32_t* idx, const int* active, int cloud0, int n_clouds, int n_query,
                         int G, int C, int k, float* grad_fv, float* grad_query, cudaStream_t st);
int launch_transpose(const float* w, int K, int N, float* wt, cudaStream_t st);

}  // namespace dpd
