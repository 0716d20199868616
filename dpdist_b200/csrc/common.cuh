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

// Operand order of layer 1 in the fp16 tensor-core path (head_tc_kernel2.cuh).  Logical order = channel-split:
// [taps x CX | taps x (C - CX) | 3 offsets | padding], CX = C & ~7.  The first nX = taps * CX / 64 K-blocks (64 elements)
// are pure X (16-byte bypass copies, fast to gather), the remaining nY hold the last X units, the Y part, the offsets and
// the padding (8-byte copies, about 2.5 x slower to issue).  Physical order = the K-blocks permuted so that the slow ones
// are spread out, one after every s - 1 fast ones, instead of forming a tail the 4-deep gather ring cannot absorb; the
// last logical K-block (the one that may be partially skipped, KernelArgs::last_ks) stays last.
__host__ __device__ inline int tc_kb_logical(int b, int nkb, int nX) {
  const int nY = nkb - nX;
  if (nY <= 1 || nX <= 0 || b == nkb - 1) return b;
  const int s = (nkb - 1) / (nY - 1) > 1 ? (nkb - 1) / (nY - 1) : 1;    // a slow K-block at b = s - 1, 2 s - 1, ...
  const int slots_upto = ((b + 1) / s) < (nY - 1) ? ((b + 1) / s) : (nY - 1);     // slow slots among positions 0 .. b
  if ((b + 1) % s == 0 && (b + 1) / s <= nY - 1) return nX + (b + 1) / s - 1;
  return b - slots_upto;
}
// position k of the channel-split (logical) order -> position tap * C + c of the reference's tap-major patch order;
// positions from taps * C on (offsets, padding) are unchanged
__host__ __device__ inline int tc_split_to_patch_k(int k, int taps, int C) {
  const int cx = C & ~7, nx = taps * cx;
  if (k < nx) return (k / cx) * C + k % cx;
  if (k < taps * C) { const int i = k - nx, cy = C - cx; return (i / cy) * C + cx + i % cy; }
  return k;
}
// physical position k of the packed operand row (length Kp, a multiple of 64) -> reference patch-order position
__host__ __device__ inline int tc_k_to_patch_k(int k, int taps, int C, int Kp) {
  const int lb = tc_kb_logical(k / 64, Kp / 64, taps * (C & ~7) / 64);
  return tc_split_to_patch_k(lb * 64 + k % 64, taps, C);
}

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
