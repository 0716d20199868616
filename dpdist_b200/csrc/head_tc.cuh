// tcgen05 tensor-core path of the implicit distance head (layers 1-3): interface used by head.cu.
// f16 = true: fp16x3 operands with power-of-two scaling (DPD_HEAD_TC); false: 3xTF32 (DPD_HEAD_TC_TF32).
#pragma once
#include "head_simt.cuh"

namespace dpd {

bool tc_supported(const dpd_head_config& c);
size_t tc_packed_bytes(const dpd_head_config& c, bool f16);
size_t tc_workspace_bytes(const dpd_head_config& c, bool f16, size_t rows);

// w1p [Kp1_src,H] (permuted, padded), w2/w3 [H,H] row-major [K_in,K_out] -> packed tensor-core operands
int tc_pack_weights(const dpd_head_config& c, bool f16, int Kp1_src, const float* w1p, const float* w2, const float* w3,
                    const float* b1, const float* b2, void* tc_blob, cudaStream_t st);

// once per head call: hi/lo split of the whole FV tensor (and, for fp16, the activation scales) into the tc workspace.
// mode 0: foreign fv (|fv|max is measured on the device)   1: |fv| <= 1 guaranteed (output of the 3DmFV kernels)
//      2: as 1, and the (hi, lo) pair was already written to tc_fv_split_ptrs() scaled by TC_FV_UNIT_SCALE
constexpr float TC_FV_UNIT_SCALE = 32768.0f;   // pow2_floor_scale(1): largest power of two s with s*1 <= 32768
int tc_prepare_fv(const dpd_head_config& c, bool f16, const float* fv, const void* tc_blob, void* tc_ws, size_t ws_rows,
                  cudaStream_t st, int mode = 0);
// The fp16 copy is CHANNEL-SPLIT: X part [cloud][voxel][C & ~7] at element 0 of hi / lo, Y part [cloud][voxel][C - (C & ~7)]
// at element *y_off (head_tc_kernel2.cuh, gather role).
void tc_fv_split_ptrs(const dpd_head_config& c, bool f16, void* tc_ws, size_t ws_rows, void** hi, void** lo, long long* y_off);

// runs layers 1..3 for `rows` chunk-local rows; *h3 points at the fp32 [rows,H] layer-3 activations.
// If fused_out != nullptr and the configuration allows it (fp16x3, 2-CTA kernel, inference) the output layer is
// fused into layer 3: fused_out [rows,3] receives the final masked distances and *h3 is set to nullptr.
int tc_head_layers(const dpd_head_config& c, bool f16, const GatherDesc& g, const float* mask, int rows, size_t ws_rows,
                   const void* tc_blob, const float* b1, const float* b2, const float* b3, float* ha, float* hb,
                   float* h3_out, void* tc_ws, const float** h3, cudaStream_t st, const float* w4 = nullptr,
                   const float* b4 = nullptr, float* fused_out = nullptr);

// backward support: fp32 layer-1 / layer-2 activations into ha / hb from the (hi, lo) pairs of the forward
int tc_merge_activations(const dpd_head_config& c, bool f16, void* tc_ws, size_t ws_rows, int rows, float* ha, float* hb,
                         cudaStream_t st);

// tensor-core backward (fp16x3, 2-CTA kernel, training configuration)
bool tc_backward_supported(const dpd_head_config& c, bool f16);
// start of a backward pass: resets the |dZ|max slots and returns the one the layer-4 backward kernel must fill
int tc_backward_begin(const dpd_head_config& c, void* tc_ws, size_t ws_rows, unsigned** slot3, cudaStream_t st);
// one layer (3, 2 or 1) of the backward pass: gw / gb (skipped when gw == nullptr) and, for layers 3 and 2, dz_next
int tc_backward_layer(const dpd_head_config& c, int layer, const void* tc_blob, void* tc_ws, size_t ws_rows, int rows,
                      const GatherDesc* g, const float* dz, float* dz_next, const int* active, float* gw, float* gb,
                      cudaStream_t st);

// input gradients: dX1 = dZ1 . W1p^T (leading dimension tc_kp1(c)) for a group of rows, after one prepare per backward
int tc_kp1(const dpd_head_config& c);
int tc_backward_inputs_prepare(const dpd_head_config& c, const void* tc_blob, void* tc_ws, size_t ws_rows, int rows,
                               const float* dz1, const int* active, cudaStream_t st);
int tc_backward_inputs_rows(const dpd_head_config& c, const void* tc_blob, void* tc_ws, size_t ws_rows, size_t r0, int nrows,
                            const int* active, float* dx1, cudaStream_t st);

int tc_debug_gemm(const float* a, int M, int K, const float* w, int N, const float* bias, float* out, void* scratch,
                  size_t scratch_bytes, int f16, cudaStream_t st);

}  // namespace dpd
