// Shared host/device helpers for libdpdist_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dpdist_b200.h"

namespace dpd {

// thread-local error message, surfaced by dpd_last_error()
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define DPD_REQUIRE(cond, code, ...)                     \
  do {                                                   \
    if (!(cond)) return ::dpd::set_error((code), __VA_ARGS__); \
  } while (0)

#define DPD_CUDA_CHECK_LAUNCH(what)                                                         \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) return ::dpd::set_error((int)e__, "%s: %s", what, cudaGetErrorString(e__)); \
  } while (0)

#define DPD_CUDA_CALL(expr)                                                                  \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) return ::dpd::set_error((int)e__, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grid tables passed to kernels by value (G <= DPD_MAX_GRID).
struct GridTables {
  float c[DPD_MAX_GRID];   // centres l[i]
  float lo[DPD_MAX_GRID];  // c - gs  (fp32, computed on the host the reference's way)
  float hi[DPD_MAX_GRID];  // c + gs
};

inline void fill_tables(GridTables& t, int G, const float* c, const float* lo, const float* hi) {
  memset(&t, 0, sizeof(t));
  for (int i = 0; i < G; ++i) {
    t.c[i] = c[i];
    if (lo) t.lo[i] = lo[i];
    if (hi) t.hi[i] = hi[i];
  }
}

int num_sms();

// cudaFuncSetAttribute is per device: remember per (call site, device) whether the opt-in shared-memory size was set.
// One process per GPU is the deployment model, but several devices in one process must work too.
struct PerDeviceOnce {
  bool done[64] = {};
  bool need() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// Counts every kernel launch; when profiling is on, brackets it with CUDA events on its stream.
struct ProfScope {
  ProfScope(const char* name, cudaStream_t st);
  ~ProfScope();
  const char* name_;
  cudaStream_t st_;
  cudaEvent_t e0_, e1_;
  bool on_;
};
#define DPD_LAUNCH(name, st, ...)            \
  do {                                       \
    ::dpd::ProfScope prof_scope__(name, st); \
    __VA_ARGS__;                             \
  } while (0)

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }

// Reference voxel assignment for one point (utils/dpdist_util.py:474-490): per axis, the FIRST
// cell whose half-open interval (lo, hi] contains p; the min flat index of a product set is the
// tuple of per-axis minima, so this equals argmax over the flat binary mask.  Flat index
// g = i0*G*G + i1*G + i2 with centre (x=l[i1], y=l[i0], z=l[i2]).
__device__ __forceinline__ int first_cell(const GridTables& t, int G, float p) {
  int r = -1;
#pragma unroll 1
  for (int i = G - 1; i >= 0; --i)
    if (p > t.lo[i] && p <= t.hi[i]) r = i;
  return r;
}

struct VoxelHit {
  int i0, i1, i2;  // valid only if inside
  int idx;         // flat index, 0 if !inside (tf.math.argmax of an all-zero row)
  bool inside;
};

__device__ __forceinline__ VoxelHit assign_voxel(const GridTables& t, int G, float x, float y, float z) {
  VoxelHit h;
  h.i1 = first_cell(t, G, x);
  h.i0 = first_cell(t, G, y);
  h.i2 = first_cell(t, G, z);
  h.inside = (h.i0 >= 0) && (h.i1 >= 0) && (h.i2 >= 0);
  if (!h.inside) h.i0 = h.i1 = h.i2 = 0;
  h.idx = (h.i0 * G + h.i1) * G + h.i2;
  return h;
}

}  // namespace dpd
