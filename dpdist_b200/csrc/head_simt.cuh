// fp32 SIMT (FFMA) kernels of the implicit distance head: the sanity path and the pieces the
// tensor-core path shares (layer 4 + output activation).
#pragma once
#include "common.cuh"

namespace dpd {

// Describes how rows of the layer-1 operand are assembled on the fly from the FV tensor:
// row r = [patch_k(fv[cloud(r)], idx[r]) (E floats, order (a0,a1,a2,ch)) | offset[r] (3) | 0-pad]
// (reference order is offset FIRST, utils/dpdist_util.py:455; the packed W1 is permuted to match).
struct GatherDesc {
  const float* fv;         // [n_clouds, V, C]
  const int32_t* idx;      // [rows] chunk-local
  const float* offset;     // [rows, 3] chunk-local
  long long row0;          // global row index of chunk-local row 0 (cloud = (row0 + m) / n_query)
  int n_query, G, C, k, E; // E = k^3*C
};

struct SimtGemmParams {
  const float* A;     // dense A [M, lda] (ignored when gathering)
  int lda;
  const float* B;     // [Kp, N] row-major
  const float* bias;  // [N]
  float* Cout;        // [M, N]
  int M, N, Kp;
  int relu;
  GatherDesc g;
};

int launch_simt_gemm(const SimtGemmParams& p, bool gather, cudaStream_t st);

// out[r, j] = mask[r] * relu6(h[r,:] . W4[:,j] + b4[j]) / 3   (utils/dpdist_util.py:539-544, 690-698)
int launch_head_out(const float* h, int ldh, const float* w4, const float* b4, const float* mask,
                    float* out, int M, int H, cudaStream_t st);

}  // namespace dpd
