// Shared declarations of the 3DmFV kernels.
#pragma once
#include "common.cuh"

namespace dpd {

struct FvParams {
  const float* points;  // [n_clouds, N, 3]
  float* fv;            // [n_clouds, V, C] or [n_clouds, C*V]
  int n_clouds, N, G, V, C;
  int full_fv, flatten;
  float sigma;
  float c[DPD_MAX_GRID];
};

// sign(x) * pow(max(|x|, 1e-12), 0.5)   (reference utils/dpdist_util.py:118-121); sign(0) = 0
__device__ __forceinline__ float power_norm(float x) {
  if (x == 0.f) return 0.f;
  return copysignf(sqrtf(fmaxf(fabsf(x), 1e-12f)), x);
}

// fv_g8.cu: specialised kernel for G = 8, full FV.  Returns 1 if the configuration is not covered.
int fv_forward_optimized(const FvParams& p, cudaStream_t stream);

}  // namespace dpd
