// tcgen05 tensor-core path of the implicit distance head (layers 1-3).
#pragma once
#include "head_simt.cuh"

namespace dpd {

bool tc_supported(const dpd_head_config& c);
size_t tc_packed_bytes(const dpd_head_config& c, int Kp1);
size_t tc_workspace_bytes(const dpd_head_config& c, size_t rows);

// w1p [Kp1,H] (permuted, padded), w2/w3 [H,H] row-major [K_in,K_out] -> packed tensor-core operands
int tc_pack_weights(const dpd_head_config& c, int Kp1, const float* w1p, const float* w2, const float* w3,
                    void* tc_blob, cudaStream_t st);

// once per head call: hi/lo split of the whole FV tensor into the tc workspace (laid out for ws_rows)
int tc_prepare_fv(const dpd_head_config& c, const float* fv, void* tc_ws, size_t ws_rows, cudaStream_t st);

// runs layers 1..3 for `rows` chunk-local rows; *h3 points at the fp32 [rows,H] layer-3 activations
int tc_head_layers(const dpd_head_config& c, int Kp1, const GatherDesc& g, int rows, size_t ws_rows, const void* tc_blob,
                   const float* b1, const float* b2, const float* b3, float* ha, float* hb, float* h3_out,
                   void* tc_ws, const float** h3, cudaStream_t st);

// backward support: where the (hi, lo) halves of the layer-2 activations live in the tc workspace
void tc_h2_buffers(const dpd_head_config& c, void* tc_ws, size_t ws_rows, float** hi, float** lo);

int tc_debug_gemm(const float* a, int M, int K, const float* w, int N, const float* bias, float* out, void* scratch,
                  size_t scratch_bytes, cudaStream_t st);

}  // namespace dpd
