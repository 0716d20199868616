// Adam update with TensorFlow-1 semantics (tf.train.AdamOptimizer defaults, used by the reference
// trainer at train_multi_gpu_pc_compare_dist.py:216):
//   lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   var -= lr_t * m / (sqrt(v) + eps)
#include <math.h>

#include "common.cuh"

namespace dpd {
namespace {
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr_t, float b1, float b2, float eps) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}
// same update with lr_t read from device memory, so that a captured CUDA graph of the whole training step can be
// replayed with a new step-dependent learning rate (the host rewrites the scalar before each replay)
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                size_t n, const float* __restrict__ lr_t_ptr, float b1, float b2, float eps) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr_t = *lr_t_ptr;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}
}  // namespace
}  // namespace dpd

extern "C" float dpd_adam_lr_t(float lr, float beta1, float beta2, int step) {
  return (float)((double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step)));
}

extern "C" int dpd_adam_step_dev(float* d_param, const float* d_grad, float* d_m, float* d_v, size_t n, const float* d_lr_t,
                                 float beta1, float beta2, float eps, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_param && d_grad && d_m && d_v && d_lr_t, DPD_E_INVALID, "dpd_adam_step_dev: null pointer");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  DPD_LAUNCH("adam", st, adam_dev_kernel<<<(unsigned)ceil_div<size_t>(n, 256), 256, 0, st>>>(d_param, d_grad, d_m, d_v, n, d_lr_t, beta1, beta2, eps));
  DPD_CUDA_CHECK_LAUNCH("adam_dev_kernel");
  return 0;
}

extern "C" int dpd_adam_step(float* d_param, const float* d_grad, float* d_m, float* d_v, size_t n, float lr,
                             float beta1, float beta2, float eps, int step, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_param && d_grad && d_m && d_v, DPD_E_INVALID, "dpd_adam_step: null pointer");
  DPD_REQUIRE(step >= 1, DPD_E_INVALID, "dpd_adam_step: step is 1-based");
  if (n == 0) return 0;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  cudaStream_t st = (cudaStream_t)stream;
  DPD_LAUNCH("adam", st, adam_kernel<<<(unsigned)ceil_div<size_t>(n, 256), 256, 0, st>>>(d_param, d_grad, d_m, d_v, n, (float)lr_t, beta1, beta2, eps));
  DPD_CUDA_CHECK_LAUNCH("adam_kernel");
  return 0;
}
