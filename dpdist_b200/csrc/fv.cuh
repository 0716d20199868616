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
  // optional second output for the tensor-core head (flatten = 0 only): fv * split_scale as an fp16 (hi, lo) pair,
  // hi = fp16(s*x), lo = fp16(s*x - hi).  |fv| <= 1 after the per-channel L2 normalisation, so s = 2^15 cannot overflow.
  // The copy is channel-split for the head's gather: channels 0 .. CX-1 (CX = C & ~7) as [cloud][voxel][CX] at element 0,
  // channels CX .. C-1 as [cloud][voxel][C - CX] at element split_y_off.
  void* fv_hi;
  void* fv_lo;
  float split_scale;
  long long split_y_off;
};

// sign(x) * pow(max(|x|, 1e-12), 0.5)   (reference utils/dpdist_util.py:118-121); sign(0) = 0
__device__ __forceinline__ float power_norm(float x) {
  if (x == 0.f) return 0.f;
  return copysignf(sqrtf(fmaxf(fabsf(x), 1e-12f)), x);
}

inline void fill_fv_params(FvParams& p, const float* points, int n_clouds, int n_points, int G, const float* h_centers,
                           float sigma, int full_fv, int flatten, float* fv) {
  p.points = points; p.fv = fv; p.n_clouds = n_clouds; p.N = n_points; p.G = G; p.V = G * G * G;
  p.full_fv = full_fv ? 1 : 0; p.flatten = flatten ? 1 : 0;
  p.C = full_fv ? DPD_FV_CHANNELS_FULL : DPD_FV_CHANNELS_SMALL;
  p.sigma = sigma;
  for (int i = 0; i < DPD_MAX_GRID; ++i) p.c[i] = i < G ? h_centers[i] : 0.f;
  p.fv_hi = nullptr; p.fv_lo = nullptr; p.split_scale = 0.f; p.split_y_off = 0;
}

// fv_g8.cu: specialised kernel for G = 8, full FV.  Returns 1 if the configuration is not covered.
int fv_forward_optimized(const FvParams& p, cudaStream_t stream);

// fv_ws.cu: warp-specialised kernel for G = 8, full FV (builder / accumulator / finaliser roles, bulk-copy staging).
// Returns 1 if the configuration is not covered.
int fv_forward_ws(const FvParams& p, cudaStream_t stream);

// fv.cu: validated dispatch used by dpd_fv_forward and dpd_model_forward.  *split_done tells whether the kernel
// that ran also produced fv_hi / fv_lo (only the G = 8 kernel does; otherwise the caller splits afterwards).
int fv_forward_dispatch(const FvParams& p, cudaStream_t stream, bool* split_done);

}  // namespace dpd
