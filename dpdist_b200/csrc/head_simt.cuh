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
  const float* gate = nullptr;   // backward mode: out = acc * (gate[m,n] > 0), no bias
  const int* active = nullptr;   // optional per-128-row-block skip flags
  GatherDesc g;
};

// ---- on-the-fly layer-1 operand (shared by the forward and the weight-gradient kernels) ----
struct RowInfo {
  long long base;  // element offset of the row's cloud in fv, or -1 for a row past M
  int i0, i1, i2;
  float off[3];
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// one element of the virtual layer-1 operand
__device__ __forceinline__ float gather_elem(const GatherDesc& g, const RowInfo& r, int kk) {
  if (r.base < 0) return 0.f;
  if (kk >= g.E) return (kk < g.E + 3) ? r.off[kk - g.E] : 0.f;
  const int j = kk / g.C, ch = kk - j * g.C;
  const int pb = (g.k - 1) >> 1;
  const int a2 = j % g.k, a1 = (j / g.k) % g.k, a0 = j / (g.k * g.k);
  const int n0 = r.i0 + a0 - pb, n1 = r.i1 + a1 - pb, n2 = r.i2 + a2 - pb;
  if ((unsigned)n0 >= (unsigned)g.G || (unsigned)n1 >= (unsigned)g.G || (unsigned)n2 >= (unsigned)g.G) return 0.f;
  return g.fv[r.base + (long long)((n0 * g.G + n1) * g.G + n2) * g.C + ch];
}

__device__ __forceinline__ float4 gather_chunk(const GatherDesc& g, const RowInfo& r, int kk, bool vec) {
  if (vec) {  // C % 4 == 0: a 4-float chunk never straddles a voxel record
    if (r.base < 0) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (kk >= g.E) return (kk == g.E) ? make_float4(r.off[0], r.off[1], r.off[2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int j = kk / g.C, ch = kk - j * g.C;
    const int pb = (g.k - 1) >> 1;
    const int a2 = j % g.k, a1 = (j / g.k) % g.k, a0 = j / (g.k * g.k);
    const int n0 = r.i0 + a0 - pb, n1 = r.i1 + a1 - pb, n2 = r.i2 + a2 - pb;
    if ((unsigned)n0 >= (unsigned)g.G || (unsigned)n1 >= (unsigned)g.G || (unsigned)n2 >= (unsigned)g.G)
      return make_float4(0.f, 0.f, 0.f, 0.f);
    return ld4(g.fv + r.base + (long long)((n0 * g.G + n1) * g.G + n2) * g.C + ch);
  }
  return make_float4(gather_elem(g, r, kk), gather_elem(g, r, kk + 1), gather_elem(g, r, kk + 2),
                     gather_elem(g, r, kk + 3));
}


__device__ __forceinline__ RowInfo make_row_info(const GatherDesc& g, int m, int M) {
  RowInfo r;
  if (m < M) {
    const long long cloud = (g.row0 + m) / g.n_query;
    r.base = cloud * (long long)(g.G * g.G * g.G) * g.C;
    const int v = g.idx[m];
    r.i2 = v % g.G; r.i1 = (v / g.G) % g.G; r.i0 = v / (g.G * g.G);
    r.off[0] = g.offset[(size_t)m * 3 + 0]; r.off[1] = g.offset[(size_t)m * 3 + 1]; r.off[2] = g.offset[(size_t)m * 3 + 2];
  } else {
    r.base = -1; r.i0 = r.i1 = r.i2 = 0; r.off[0] = r.off[1] = r.off[2] = 0.f;
  }
  return r;
}


int launch_simt_gemm(const SimtGemmParams& p, bool gather, cudaStream_t st);

// out[r, j] = mask[r] * relu6(h[r,:] . W4[:,j] + b4[j]) / 3   (utils/dpdist_util.py:539-544, 690-698)
int launch_head_out(const float* h, int ldh, const float* w4, const float* b4, const float* mask,
                    float* out, int M, int H, cudaStream_t st);

}  // namespace dpd
