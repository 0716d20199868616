// tcgen05 tensor-core path of the implicit distance head (layers 1-3), sm_100a: host side.
//
// Precision: the reference computes these convolutions in fp32 (tf.nn.conv2d on fp32 tensors,
// utils/tf_util.py:213) and parity is 1e-4, which a single TF32/BF16/FP16 pass cannot hold at K = 2503.
// Every operand is therefore split x = hi + lo and each K-step issues three MMAs, Al*Bh + Ah*Bl + Ah*Bh
// (the dropped Al*Bl is 2^-22 relative), into segment accumulators that are promoted in fp32 registers
// (head_tc_kernel.cuh).  Two operand formats:
//   DPD_HEAD_TC      fp16x3: hi = fp16(s*x), lo = fp16(s*x - hi) with s a power of two chosen from a
//                    rigorous bound of the tensor (|fv| measured; activations bounded by column L1 norms of
//                    the weights), so nothing can overflow; 11+11 mantissa bits like TF32, but the fp16 pipe
//                    runs at twice the TF32 rate and the operands are half the bytes.
//   DPD_HEAD_TC_TF32 3xTF32: hi = tf32(x), lo = tf32(x - hi); no scaling needed (fp32 exponent range).
#include <vector>
#include <cstdio>
#include "head_bwd.cuh"
#include "head_tc.cuh"
#include "head_tc_kernel.cuh"
#include "head_tc_kernel2.cuh"
#include <stdlib.h>

namespace dpd {
namespace tc {

// ---------------------------------------------------------------------------------------------
// helper kernels: TF32 format
// ---------------------------------------------------------------------------------------------
// Wt_hi/lo[n][k] from W[k][n] (row-major [K,N]); 32x32 smem transpose
__global__ void transpose_split_kernel(const float* __restrict__ w, int K, int N, float* __restrict__ t_hi,
                                       float* __restrict__ t_lo) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? w[(size_t)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) {
      const float x = tile[threadIdx.x][i];
      const float hi = tf32_rna(x);
      t_hi[(size_t)n * K + k] = hi;
      t_lo[(size_t)n * K + k] = tf32_rna(x - hi);
    }
  }
}

__global__ void split_kernel(const float* __restrict__ x, size_t n, float* __restrict__ hi, float* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const float h = tf32_rna(v);
  hi[i] = h;
  lo[i] = tf32_rna(v - h);
}

__global__ void split_off4_kernel(const float* __restrict__ off, int rows, float* __restrict__ hi, float* __restrict__ lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float4 h, l;
  const float a = off[(size_t)i * 3], b = off[(size_t)i * 3 + 1], c = off[(size_t)i * 3 + 2];
  h.x = tf32_rna(a); h.y = tf32_rna(b); h.z = tf32_rna(c); h.w = 0.f;
  l.x = tf32_rna(a - h.x); l.y = tf32_rna(b - h.y); l.z = tf32_rna(c - h.z); l.w = 0.f;
  reinterpret_cast<float4*>(hi)[i] = h;
  reinterpret_cast<float4*>(lo)[i] = l;
}

// ---------------------------------------------------------------------------------------------
// helper kernels: scaled FP16 format
// ---------------------------------------------------------------------------------------------
// scale slots (floats): per-forward values live in the workspace, per-pack values in the blob
enum { S_A1 = 0, S_A2, S_A3, S_ACC1, S_ACC2, S_ACC3, S_FVMAX_BITS, S_COUNT = 8 };          // workspace
enum { P_W1 = 0, P_W2, P_W3, P_W1MAX_BITS, P_W2MAX_BITS, P_W3MAX_BITS, P_C1_BITS, P_C2_BITS, P_B1MAX_BITS, P_B2MAX_BITS, P_COUNT = 12 };  // blob

__device__ __forceinline__ float pow2_floor_scale(float bound) {
  // largest power of two s with s*bound <= 32768 (half of the fp16 maximum), clamped for degenerate bounds
  const float b = fmaxf(bound, 1e-30f);
  float e = floorf(log2f(32768.0f / b));
  e = fminf(fmaxf(e, -60.f), 60.f);
  return exp2f(e);
}

__global__ void absmax_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ out_bits) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order like their bits
}

// max over columns n of sum_k |w[k][n]|   (w row-major [K, N]).  Block = 32 columns x 32 row lanes; the row lanes of a
// column are summed in a fixed order (deterministic), then the block's 32 column sums are max-reduced.
__global__ void __launch_bounds__(1024) col_l1_max_kernel(const float* __restrict__ w, int K, int N, unsigned* __restrict__ out_bits) {
  __shared__ float part[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (n < N)
    for (int k = ty; k < K; k += 32) s += fabsf(w[(size_t)k * N + n]);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float c = 0.f;
    for (int j = 0; j < 32; ++j) c += part[j][tx];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, o));
    if (tx == 0) atomicMax(out_bits, __float_as_uint(c));
  }
}

__global__ void weight_scales_kernel(float* __restrict__ ps) {   // one thread
  unsigned* pb = reinterpret_cast<unsigned*>(ps);
  ps[P_W1] = pow2_floor_scale(__uint_as_float(pb[P_W1MAX_BITS]));
  ps[P_W2] = pow2_floor_scale(__uint_as_float(pb[P_W2MAX_BITS]));
  ps[P_W3] = pow2_floor_scale(__uint_as_float(pb[P_W3MAX_BITS]));
}

// per forward: activation scales from |fv|max and the weight column norms (rigorous bounds, no overflow possible):
//   |A1| <= in1 = max(|fv|max, 1)            (offsets are < 1: query - centre of its own voxel, masked rows zeroed)
//   |H1| <= B1 = in1*c1 + max|b1|,  |H2| <= B2 = B1*c2 + max|b2|     (c_l = max column L1 norm of W_l)
__global__ void activation_scales_kernel(float* __restrict__ ws, const float* __restrict__ ps, bool unit_bound) {   // one thread
  const unsigned* wb = reinterpret_cast<const unsigned*>(ws);
  const unsigned* pb = reinterpret_cast<const unsigned*>(ps);
  const float in1 = unit_bound ? 1.0f : fmaxf(__uint_as_float(wb[S_FVMAX_BITS]), 1.0f);
  const float B1 = in1 * __uint_as_float(pb[P_C1_BITS]) + __uint_as_float(pb[P_B1MAX_BITS]);
  const float B2 = B1 * __uint_as_float(pb[P_C2_BITS]) + __uint_as_float(pb[P_B2MAX_BITS]);
  const float a1 = pow2_floor_scale(in1), a2 = pow2_floor_scale(B1), a3 = pow2_floor_scale(B2);
  ws[S_A1] = a1; ws[S_A2] = a2; ws[S_A3] = a3;
  ws[S_ACC1] = 1.0f / (a1 * ps[P_W1]); ws[S_ACC2] = 1.0f / (a2 * ps[P_W2]); ws[S_ACC3] = 1.0f / (a3 * ps[P_W3]);
}

__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

// (hi, lo) = fp16 split of x * scale for active 128-row blocks, zeros for inactive ones
__global__ void split_f16_active_kernel(const float* __restrict__ x, int rows, int H, const int* __restrict__ active,
                                        const float* __restrict__ scale, __half* __restrict__ hi, __half* __restrict__ lo) {
  const int blk = blockIdx.x;
  const bool on = !active || active[blk];
  const float s = *scale;
  const size_t b0 = (size_t)blk * 128 * H, b1 = min((size_t)rows, (size_t)(blk + 1) * 128) * H;
  for (size_t i = b0 + threadIdx.x * 4; i < b1; i += blockDim.x * 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) v = *reinterpret_cast<const float4*>(x + i);
    __half h[4], l[4];
    split_half(v.x * s, h[0], l[0]); split_half(v.y * s, h[1], l[1]);
    split_half(v.z * s, h[2], l[2]); split_half(v.w * s, h[3], l[3]);
    *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l);
  }
}

// One pass over the upstream gradient dz [rows, H] (fp32): the scaled fp16 (hi, lo) pair (operand of both products of
// the layer: dX = dZ . W^T and, MN-major, dW = A^T . dZ), zeros for inactive 128-row blocks, and the per-64-row-block
// column sums that make up the bias gradient.
__global__ void split_colsum_f16_kernel(const float* __restrict__ dz, int rows, int H, const int* __restrict__ active,
                                        const float* __restrict__ scale, __half* __restrict__ hi, __half* __restrict__ lo,
                                        const int* __restrict__ extent, float* __restrict__ col_partial) {
  __shared__ float csum[8][64];
  const int m0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  if (extent && m0 >= *extent) return;       // past the last active block: neither product reads these rows
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool on = !active || active[m0 >> 7];
  const float s = *scale;
  float cs0 = 0.f, cs1 = 0.f;               // column sums of this thread's 8 rows (bias gradient, fixed order)
  for (int r = ty; r < 64; r += 8) {
    const int m = m0 + r;
    float2 v = make_float2(0.f, 0.f);
    if (on && m < rows) v = *reinterpret_cast<const float2*>(dz + (size_t)m * H + c0 + 2 * tx);
    cs0 += v.x; cs1 += v.y;
    __half h0, l0, h1, l1;
    split_half(v.x * s, h0, l0); split_half(v.y * s, h1, l1);
    if (m < rows) {
      *reinterpret_cast<__half2*>(hi + (size_t)m * H + c0 + 2 * tx) = __halves2half2(h0, h1);
      *reinterpret_cast<__half2*>(lo + (size_t)m * H + c0 + 2 * tx) = __halves2half2(l0, l1);
    }
  }
  if (col_partial == nullptr) return;
  csum[ty][2 * tx] = cs0; csum[ty][2 * tx + 1] = cs1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a += csum[j][threadIdx.x];
    col_partial[(size_t)blockIdx.x * H + c0 + threadIdx.x] = a;
  }
}

// gb[c] = sum over the 64-row blocks below the active extent of col_partial[b][c].  CTA = 64 columns x 16 block lanes;
// lane j sums blocks j, j+16, ... in order, then the 16 lane sums are added in order (deterministic).
__global__ void __launch_bounds__(1024) colsum_blocks_final_kernel(const float* __restrict__ col_partial, const int* __restrict__ extent,
                                                                   int H, float* __restrict__ gb) {
  __shared__ float red[16][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + tx;
  const int nb = (*extent + 63) / 64;
  float a = 0.f;
  for (int b = ty; b < nb; b += 16) a += col_partial[(size_t)b * H + c];
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) t += red[j][tx];
    gb[c] = t;
  }
}

// Per query row: {index of its own voxel record = cloud * V + voxel (or -1 past the end), validity bits of the k
// taps per axis (bit a: i0 + a - pb in range, bit 8 + a: i1, bit 16 + a: i2)} -- what the forward gather producers
// compute per tile, prepared once per backward pass for the gather producers of the dW1 product.
__global__ void rowinfo_kernel(const int32_t* __restrict__ idx, long long row0, int n_query, int G, int Cc, int k, int rows,
                               int2* __restrict__ out, int2* __restrict__ lut, int lut_chunks, int E) {
  if (lut != nullptr && blockIdx.x == 0) build_gather_lut(lut, lut_chunks, Cc, k, G, E, threadIdx.x, blockDim.x);
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  const int V = G * G * G, pb = (k - 1) >> 1;
  const long long cloud = (row0 + m) / n_query;
  const int v = idx[m];
  const int i0 = v / (G * G), i1 = (v / G) % G, i2 = v % G;
  uint32_t mk = 0;
  for (int a = 0; a < k; ++a) {
    mk |= ((unsigned)(i0 + a - pb) < (unsigned)G ? 1u : 0u) << a;
    mk |= ((unsigned)(i1 + a - pb) < (unsigned)G ? 1u : 0u) << (8 + a);
    mk |= ((unsigned)(i2 + a - pb) < (unsigned)G ? 1u : 0u) << (16 + a);
  }
  out[m] = make_int2((int)(cloud * V) + v, (int)mk);
}

// rows covered by the active blocks: (index of the last active 128-row block + 1) * 128, clipped to `rows`
__global__ void active_extent_kernel(const int* __restrict__ active, int nblk, int rows, int* __restrict__ out) {   // one warp
  int last = -1;
  for (int b = threadIdx.x; b < nblk; b += 32) if (active[b]) last = b;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (threadIdx.x == 0) *out = min(rows, (last + 1) * 128);
}

// backward products: sc[0] = |dZ|max bits -> sc[1] = s_g, sc[2] = 1 / (s_g * *other_scale)
// and sc[3] = 1 / (s_g * *act_scale) for the weight-gradient product
__global__ void bwd_scales_kernel(float* __restrict__ sc, const float* __restrict__ other_scale, const float* __restrict__ act_scale,
                                  const unsigned* __restrict__ amax_bits) {   // one thread
  const float m = __uint_as_float(amax_bits ? *amax_bits : reinterpret_cast<const unsigned*>(sc)[0]);
  const float sg = m > 0.f ? pow2_floor_scale(m) : 1.0f;
  sc[1] = sg;
  sc[2] = other_scale ? 1.0f / (sg * *other_scale) : 0.f;
  sc[3] = 1.0f / (sg * *act_scale);
}

__global__ void debug_scales_kernel(float* __restrict__ sc) {   // one thread (tc_debug_gemm)
  const unsigned* b = reinterpret_cast<const unsigned*>(sc);
  sc[0] = pow2_floor_scale(__uint_as_float(b[4]));
  sc[1] = pow2_floor_scale(__uint_as_float(b[5]));
  sc[2] = 1.0f / (sc[0] * sc[1]);
}


// Wt_hi/lo[n][k] (fp16, [N, Kp], zero beyond K) from W[k][n] times *scale.  With taps > 0 (layer 1 of the 2-CTA kernel)
// column k of the result is row tc_k_to_patch_k(k) of W: the physical operand order of the gather producers (common.cuh).
__global__ void transpose_split_f16_kernel(const float* __restrict__ w, int K, int Kp, int N, const float* __restrict__ scale,
                                           __half* __restrict__ t_hi, __half* __restrict__ t_lo, int taps, int C) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const float s = *scale;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    const int ks = (taps > 0 && k < Kp) ? tc_k_to_patch_k(k, taps, C, Kp) : k;
    tile[i][threadIdx.x] = (ks < K && n < N) ? w[(size_t)ks * N + n] * s : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < Kp) {
      __half hi, lo;
      split_half(tile[threadIdx.x][i], hi, lo);
      t_hi[(size_t)n * Kp + k] = hi;
      t_lo[(size_t)n * Kp + k] = lo;
    }
  }
}

__global__ void split_f16_kernel(const float* __restrict__ x, size_t n, const float* __restrict__ scale, __half* __restrict__ hi,
                                 __half* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  split_half(x[i] * *scale, hi[i], lo[i]);
}

// 3DmFV tensor [records, C] fp32 -> scaled fp16 (hi, lo) copy in the channel-split layout of the 2-CTA gather:
// X part [records, CX] (CX = C & ~7) at element 0, Y part [records, C - CX] at element y_off
__global__ void split_fv_f16_kernel(const float* __restrict__ x, size_t n, int C, long long y_off, const float* __restrict__ scale,
                                    __half* __restrict__ hi, __half* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t rec = i / C;
  const int c = (int)(i - rec * C), cx = C & ~7;
  const size_t o = c < cx ? rec * cx + c : (size_t)y_off + rec * (C - cx) + (c - cx);
  split_half(x[i] * *scale, hi[o], lo[o]);
}

// offsets of rows outside the unit cube are zeroed: their output is masked to 0 anyway (:697-698) and an
// arbitrary far-away query must not be able to overflow fp16
__global__ void split_off4_f16_kernel(const float* __restrict__ off, const float* __restrict__ mask, int rows,
                                      const float* __restrict__ scale, __half* __restrict__ hi, __half* __restrict__ lo,
                                      const int32_t* __restrict__ idx, long long row0, int n_query, int G, int Cc, int k,
                                      int2* __restrict__ rowinfo, int2* __restrict__ lut, int lut_chunks, int E) {
  if (lut != nullptr && blockIdx.x == 0) build_gather_lut(lut, lut_chunks, Cc, k, G, E, threadIdx.x, blockDim.x);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  if (rowinfo != nullptr) {     // the gather producers' per-row words, same contents as rowinfo_kernel
    const int V = G * G * G, pb = (k - 1) >> 1;
    const long long cloud = (row0 + i) / n_query;
    const int v = idx[i];
    const int i0 = v / (G * G), i1 = (v / G) % G, i2 = v % G;
    uint32_t mk = 0;
    for (int a = 0; a < k; ++a) {
      mk |= ((unsigned)(i0 + a - pb) < (unsigned)G ? 1u : 0u) << a;
      mk |= ((unsigned)(i1 + a - pb) < (unsigned)G ? 1u : 0u) << (8 + a);
      mk |= ((unsigned)(i2 + a - pb) < (unsigned)G ? 1u : 0u) << (16 + a);
    }
    rowinfo[i] = make_int2((int)(cloud * V) + v, (int)mk);
  }
  const float s = mask[i] != 0.f ? *scale : 0.f;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    __half h = __float2half_rn(0.f), l = h;
    if (d < 3 && s != 0.f) split_half(off[(size_t)i * 3 + d] * s, h, l);
    hi[(size_t)i * 4 + d] = h;
    lo[(size_t)i * 4 + d] = l;
  }
}

// fp32 activations for the backward pass: (hi + lo) / scale
__global__ void merge_f16_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, const float* __restrict__ scale,
                                 float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (__half2float(hi[i]) + __half2float(lo[i])) / *scale;
}

// output layer finish for the fused layer-3/4 kernel: sums the per-(N-tile, half) partial dot products in
// fixed order, adds b4, relu6(x)/3 and the in-cube mask (utils/dpdist_util.py:690-691, 697-698).
__global__ void head_out_finish_kernel(const float4* __restrict__ part4, int nslots, const float* __restrict__ b4,
                                       const float* __restrict__ mask, float* __restrict__ out, int M) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int i = 0; i < nslots; ++i) {
    const float4 v = part4[(size_t)row * nslots + i];
    s0 += v.x; s1 += v.y; s2 += v.z;
  }
  const int orow = row;
  const float m = mask[orow];
  out[(size_t)orow * 3 + 0] = fminf(fmaxf(s0 + b4[0], 0.f), 6.f) / 3.0f * m;
  out[(size_t)orow * 3 + 1] = fminf(fmaxf(s1 + b4[1], 0.f), 6.f) / 3.0f * m;
  out[(size_t)orow * 3 + 2] = fminf(fmaxf(s2 + b4[2], 0.f), 6.f) / 3.0f * m;
}


// ---------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// [rows, cols] row-major (pitch = cols) of fp32 or fp16, box = {128 bytes of columns, box_rows}, 128-byte swizzle
static int make_tmap(CUtensorMap* m, const void* base, bool f16, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  DPD_REQUIRE(fn != nullptr, DPD_E_UNSUPPORTED, "cuTensorMapEncodeTiled entry point not found");
  const int elem = f16 ? 2 : 4;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * (cuuint64_t)elem};
  cuuint32_t box[2] = {(cuuint32_t)(ROW_BYTES / elem), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DPD_REQUIRE(r == CUDA_SUCCESS, DPD_E_INVALID, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r,
              (unsigned long long)rows, (unsigned long long)cols);
  return 0;
}

template <bool GATHER, bool F16>
static int launch_t(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi, const CUtensorMap& tb_lo,
                    const KernelArgs& ka, int grid, size_t smem, cudaStream_t st) {
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    DPD_CUDA_CALL(cudaFuncSetAttribute(tc_gemm_kernel<GATHER, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
  }
  const char* name = GATHER ? (F16 ? "tc_gemm_gather_l1_f16" : "tc_gemm_gather_l1_tf32") : (F16 ? "tc_gemm_dense_f16" : "tc_gemm_dense_tf32");
  DPD_LAUNCH(name, st, tc_gemm_kernel<GATHER, F16><<<grid, GATHER ? 512 : 384, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, ka));
  DPD_CUDA_CHECK_LAUNCH("tc_gemm_kernel");
  return 0;
}

template <int GM>
static int launch2_t(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi, const CUtensorMap& tb_lo,
                     const KernelArgs& ka, int grid, size_t smem, cudaStream_t st) {
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    DPD_CUDA_CALL(cudaFuncSetAttribute(tc_gemm2_kernel<GM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
  }
  DPD_LAUNCH(GM ? (ka.mn_major ? "tc_gemm2_bwd_dw1_gather_f16" : "tc_gemm2_gather_l1_f16") : (ka.part4 ? "tc_gemm2_dense_l3_l4_f16" : (ka.mode == 1 ? "tc_gemm2_bwd_dx_f16" : (ka.mode == 2 ? "tc_gemm2_bwd_dw_f16" : "tc_gemm2_dense_f16"))), st,
             tc_gemm2_kernel<GM><<<grid, GM ? THREADS2_GATHER : 384, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, ka));
  DPD_CUDA_CHECK_LAUNCH("tc_gemm2_kernel");
  return 0;
}

// The fp16x3 path always runs on the 2-CTA (cta_group::2) kernel, 256x256 tiles on CTA pairs: the fp16 copy of the 3DmFV
// tensor is stored in the layout its gather producers read.  The single-CTA kernel serves the 3xTF32 format.
static bool use_2cta() { return true; }

// DPD_TC_FUSE_L4=0 keeps the separate output-layer kernel (A/B measurements)
static bool fuse_l4() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DPD_TC_FUSE_L4"); v = e ? (atoi(e) != 0) : 1; }
  return v != 0;
}

// optional backward-pass behaviour of launch2 (see KernelArgs)
struct BwdExtras {
  int mode = 0;
  const uint4* relu_bits_in = nullptr;
  const int* active = nullptr;
  int slices = 1;
  const int* k_limit = nullptr;
  long long slice_stride = 0;
  unsigned* absmax_bits = nullptr;
  int lut_chunks = 0;
  int mn_major = 0;          // A [K, M] and B [K, N] row-major (K = reduction): D = A^T . B
  unsigned long long mn_rows = 0;   // valid rows of A and B in that case
  const char* name = nullptr;
};

static int launch2(bool gather, const void* a_hi, const void* a_lo, int M, int K, const void* bt_hi, const void* bt_lo, int N,
                   const float* bias, void* out0, void* out1, int split, const float* acc_scale, const float* out_scale,
                   const GatherArgs* g, cudaStream_t st, const float* w4 = nullptr, float* part4 = nullptr,
                   const BwdExtras* bx = nullptr, uint4* relu_bits_out = nullptr, bool long_head = false) {
  DPD_REQUIRE(K % 64 == 0 && N % BN == 0 && M > 0, DPD_E_UNSUPPORTED, "tc gemm2: need K %% 64 == 0, N %% 256 == 0 (K=%d N=%d)", K, N);
  DPD_REQUIRE(!gather || (bx && bx->mn_major ? bx->lut_chunks : K / 4) <= MAX_LUT, DPD_E_UNSUPPORTED, "tc gemm2: operand row too long");
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  if (bx && bx->mn_major) {
    // operands as stored: A [K rows, M columns], B [K rows, N columns]; boxes of 64 columns x 64 reduction rows.
    // K here is the VALID row count of both arrays (rows beyond it are never read: TMA fills zeros).  With `gather`
    // the A operand is assembled by the gather warps instead (dW1).
    DPD_REQUIRE(M % 64 == 0, DPD_E_UNSUPPORTED, "tc gemm2 (MN-major): M %% 64 != 0");
    if ((rc = make_tmap(&tb_hi, bt_hi, true, bx->mn_rows, N, 64))) return rc;
    if ((rc = make_tmap(&tb_lo, bt_lo, true, bx->mn_rows, N, 64))) return rc;
    if (!gather) {
      if ((rc = make_tmap(&ta_hi, a_hi, true, bx->mn_rows, M, 64))) return rc;
      if ((rc = make_tmap(&ta_lo, a_lo, true, bx->mn_rows, M, 64))) return rc;
    } else {
      ta_hi = tb_hi; ta_lo = tb_lo;
    }
  } else {
    if ((rc = make_tmap(&tb_hi, bt_hi, true, N, K, BN / 2))) return rc;
    if ((rc = make_tmap(&tb_lo, bt_lo, true, N, K, BN / 2))) return rc;
    if (!gather) {
      if ((rc = make_tmap(&ta_hi, a_hi, true, M, K, BM))) return rc;
      if ((rc = make_tmap(&ta_lo, a_lo, true, M, K, BM))) return rc;
    } else {
      ta_hi = tb_hi; ta_lo = tb_lo;
    }
  }
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.M = M; ka.N = N; ka.num_kb = K / 64; ka.bias = bias; ka.out0 = out0; ka.out1 = out1; ka.split = split;
  ka.acc_scale = acc_scale; ka.out_scale = out_scale; ka.w4 = w4; ka.part4 = part4; ka.relu_bits_out = relu_bits_out;
  const bool gather_mn = gather && bx && bx->mn_major;
  // Longer first promotion segments (see seg_end in the kernel): inference forward only.  Training keeps the short
  // segments everywhere: the gradient tests compare ReLU gates with the fp64 oracle's, and the backward is a chain of products.
  { const char* e = getenv("DPD_TC_SEG_HEAD"); ka.seg_head = (bx || !long_head) ? 0 : (e ? atoi(e) : (gather ? 6 : 5)); }
  if (g) {
    ka.g = *g;
    if (gather && !gather_mn) {   // valid operand length E + 3; everything from there to K is zero padding
      const int tail = g->E + 3 - (K - 64);
      ka.last_ks = tail >= 64 ? 0 : (tail <= 0 ? 1 : ceil_div(tail, 16));
    }
  }
  int tiles = ceil_div(M, 2 * BM) * (N / BN);
  if (bx) {
    DPD_REQUIRE((!gather || gather_mn) && !split && part4 == nullptr, DPD_E_UNSUPPORTED,
                "tc gemm2: backward modes need the fp32-output kernel (gather only with MN-major operands)");
    ka.lut_chunks = bx->lut_chunks; ka.g_rows = (int)bx->mn_rows;
    ka.mode = bx->mode; ka.relu_bits_in = bx->relu_bits_in; ka.active = bx->active; ka.k_limit = bx->k_limit;
    ka.slices = bx->slices; ka.slice_stride = bx->slice_stride; ka.absmax_bits = bx->absmax_bits; ka.mn_major = bx->mn_major;
    // K-blocks per slice: a multiple of the promotion segment so that segments never straddle slices
    ka.kb_per_slice = round_up(ceil_div(K / 64, bx->slices > 1 ? bx->slices : 1), 4);
    tiles *= bx->slices > 1 ? bx->slices : 1;
  }
  const int clusters = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
  const size_t smem = 1024 + (gather ? (size_t)GATHER_RING_BYTES : (size_t)STAGES2 * STAGE2_BYTES) + sizeof(SharedCtl2);
  const int gm = gather ? 1 : 0;
  const size_t smem_use = smem;
  auto go = [&]() -> int {
    return gm == 1 ? launch2_t<1>(ta_hi, ta_lo, tb_hi, tb_lo, ka, 2 * clusters, smem_use, st)
                   : launch2_t<0>(ta_hi, ta_lo, tb_hi, tb_lo, ka, 2 * clusters, smem_use, st);
  };
  // DPD_TC_TRACE=<prefix>: timeline of cluster 0 (clock64 stamps per warp role) of launches 20..22 of every kernel flavour,
  // written to <prefix>.<flavour>.<n>.bin after a stream synchronise (tools/tc_trace.py reads them).  Measurement aid only.
  if (const char* tp = getenv("DPD_TC_TRACE")) {
    static int counter[2] = {0, 0};
    static unsigned long long* buf = nullptr;
    const int n = counter[gather ? 1 : 0]++;
    if (n >= 20 && n < 23 && !bx) {
      const size_t bytes = 4 * 65536 * sizeof(unsigned long long);
      if (!buf) DPD_CUDA_CALL(cudaMalloc(&buf, bytes));
      DPD_CUDA_CALL(cudaMemsetAsync(buf, 0, bytes, st));
      ka.trace = buf;
      const int rc2 = go();
      DPD_CUDA_CALL(cudaStreamSynchronize(st));
      std::vector<unsigned long long> h(4 * 65536);
      DPD_CUDA_CALL(cudaMemcpy(h.data(), buf, bytes, cudaMemcpyDeviceToHost));
      char path[512];
      snprintf(path, sizeof(path), "%s.%s.K%d.%d.bin", tp, (gather ? "gather" : (part4 ? "dense_l3" : "dense")), K, n);
      if (FILE* f = fopen(path, "wb")) { fwrite(h.data(), 1, bytes, f); fclose(f); }
      return rc2;
    }
  }
  return go();
}

// D[M,N] = relu(acc_scale * A[M,K] * B[N,K]^T + bias), operands pre-split (hi, lo); K % kb == 0, N % 256 == 0
static int launch(bool gather, bool f16, const void* a_hi, const void* a_lo, int M, int K, const void* bt_hi, const void* bt_lo,
                  int N, const float* bias, void* out0, void* out1, int split, const float* acc_scale, const float* out_scale,
                  const GatherArgs* g, cudaStream_t st, uint4* relu_bits_out = nullptr, bool long_head = false) {
  if (f16)
    return launch2(gather, a_hi, a_lo, M, K, bt_hi, bt_lo, N, bias, out0, out1, split, acc_scale, out_scale, g, st, nullptr, nullptr,
                   nullptr, relu_bits_out, long_head);
  const int kb = 32;
  DPD_REQUIRE(K % kb == 0 && N % BN == 0 && M > 0, DPD_E_UNSUPPORTED, "tc gemm: need K %% %d == 0, N %% 256 == 0 (K=%d N=%d)", kb, K, N);
  DPD_REQUIRE(K / 4 <= MAX_LUT, DPD_E_UNSUPPORTED, "tc gemm: K=%d too large", K);
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  if ((rc = make_tmap(&tb_hi, bt_hi, f16, N, K, BN))) return rc;
  if ((rc = make_tmap(&tb_lo, bt_lo, f16, N, K, BN))) return rc;
  if (!gather) {
    if ((rc = make_tmap(&ta_hi, a_hi, f16, M, K, BM))) return rc;
    if ((rc = make_tmap(&ta_lo, a_lo, f16, M, K, BM))) return rc;
  } else {
    ta_hi = tb_hi; ta_lo = tb_lo;   // unused
  }
  KernelArgs ka;
  memset(&ka, 0, sizeof(ka));
  ka.M = M; ka.N = N; ka.num_kb = K / kb; ka.bias = bias; ka.out0 = out0; ka.out1 = out1; ka.split = split;
  ka.acc_scale = acc_scale; ka.out_scale = out_scale;
  if (g) ka.g = *g;
  const int tiles = ceil_div(M, BM) * (N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  const size_t smem = 1024 + (size_t)STAGES * STAGE_BYTES + sizeof(SharedCtl) + 2 * (size_t)(K / 4) * sizeof(uint32_t);
  // the single-CTA kernel serves the 3xTF32 format only (its F16 flavour read the interleaved fp16 copy that no longer exists)
  if (gather) return launch_t<true, false>(ta_hi, ta_lo, tb_hi, tb_lo, ka, grid, smem, st);
  return launch_t<false, false>(ta_hi, ta_lo, tb_hi, tb_lo, ka, grid, smem, st);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// interface used by head.cu
// ---------------------------------------------------------------------------------------------
bool tc_supported(const dpd_head_config& c) {
  return c.H % tc::BN == 0 && c.C % 4 == 0 && c.G <= 255 && c.k <= 8 &&
         (long long)c.n_clouds * c.G * c.G * c.G * c.C < (1ll << 31);
}

namespace {
size_t up256(size_t x) { return round_up<size_t>(x, 256); }
int kp1_of(const dpd_head_config& c, bool f16) { return round_up(c.k * c.k * c.k * c.C + 3, f16 ? 64 : 32); }

bool tc_train(const dpd_head_config& c) { return (c.flags & DPD_HEAD_TRAIN) != 0; }
// tensor-core backward: fp16x3 2-CTA kernel only (DPD_TC_BWD=0 keeps the fp32 SIMT backward for A/B measurements)
bool tc_bwd_env() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DPD_TC_BWD"); v = e ? (atoi(e) != 0) : 1; }
  return v != 0;
}

struct TcBlob { size_t w1h, w1l, w2h, w2l, w3h, w3l, w2nh, w2nl, w3nh, w3nl, w1nh, w1nl, scales, total; };
TcBlob tc_blob_layout(const dpd_head_config& c, bool f16) {
  TcBlob b; size_t o = 0; const size_t H = c.H, e = f16 ? 2 : 4, Kp1 = kp1_of(c, f16);
  b.w1h = o; o += up256(H * Kp1 * e); b.w1l = o; o += up256(H * Kp1 * e);
  b.w2h = o; o += up256(H * H * e);   b.w2l = o; o += up256(H * H * e);
  b.w3h = o; o += up256(H * H * e);   b.w3l = o; o += up256(H * H * e);
  b.w2nh = b.w2nl = b.w3nh = b.w3nl = o;
  if (f16 && tc_train(c)) {   // W2, W3 as stored ([K_in, K_out] = the K-major B operand of dX = dZ . W^T)
    b.w2nh = o; o += up256(H * H * e); b.w2nl = o; o += up256(H * H * e);
    b.w3nh = o; o += up256(H * H * e); b.w3nl = o; o += up256(H * H * e);
  }
  b.w1nh = b.w1nl = o;
  if (f16 && tc_train(c) && (c.flags & DPD_HEAD_INPUT_GRAD)) {   // W1p as stored ([Kp1, H]): B operand of dX1 = dZ1 . W1p^T
    b.w1nh = o; o += up256(H * Kp1 * e); b.w1nl = o; o += up256(H * Kp1 * e);
  }
  b.scales = o; o += up256(tc::P_COUNT * 4);
  b.total = o; return b;
}
constexpr int TC_DW_SLICES = 9;       // dW2 / dW3: 16 tiles x 9 slices = 144 work items on 74 clusters (1.95 waves)
constexpr int TC_DW1_SLICES = 11;     // dW1: 40 tiles x 11 = 440 (5.95 waves)
struct TcWs { size_t fvh, fvl, o4h, o4l, xh, xl, yh, yl, gh, gl, rinfo, lut, part, cpart, rb1, rb2, bsc, scales, total; };
TcWs tc_ws_layout(const dpd_head_config& c, bool f16, size_t rows) {
  TcWs w; size_t o = 0; const size_t e = f16 ? 2 : 4;
  const size_t nfv = (size_t)c.n_clouds * c.G * c.G * c.G * c.C;
  w.fvh = o; o += up256(nfv * e); w.fvl = o; o += up256(nfv * e);
  w.o4h = o; o += up256(rows * 4 * e); w.o4l = o; o += up256(rows * 4 * e);
  w.xh = o; o += up256(rows * (size_t)c.H * e); w.xl = o; o += up256(rows * (size_t)c.H * e);
  w.yh = w.yl = o;
  if (f16) { w.yh = o; o += up256(rows * (size_t)c.H * e); w.yl = o; o += up256(rows * (size_t)c.H * e); }
  w.rinfo = o;
  if (f16) o += up256(rows * 8);      // per-row gather info (forward: split_off4_f16_kernel, backward: rowinfo_kernel)
  w.lut = o;
  if (f16) o += up256((size_t)tc::MAX_LUT * 8);   // gather LUT of the 2-CTA kernel, {code, delta} per 4-element chunk
  w.gh = w.gl = w.part = w.cpart = w.rb1 = w.rb2 = w.bsc = o;
  if (f16 && tc_train(c)) {   // backward: (hi, lo) of the upstream gradient, partials
    const size_t act = up256(rows * (size_t)c.H * e), Kp1 = kp1_of(c, true);
    w.gh = o; o += act; w.gl = o; o += act;
    const size_t p1 = (size_t)TC_DW1_SLICES * Kp1 * c.H * 4, p2 = (size_t)TC_DW_SLICES * c.H * c.H * 4;
    w.part = o; o += up256(p1 > p2 ? p1 : p2);
    w.cpart = o; o += up256((rows / 64 + 1) * (size_t)c.H * 4);      // per-64-row-block column sums of dZ
    // ReLU' bit masks of H1 / H2: one uint4 per (256x256 tile, CTA of the pair, epilogue warp, lane)
    const size_t nbits = ((rows + 255) / 256) * (size_t)(c.H / tc::BN) * 2 * tc::NUM_EPI_WARPS * 32 * 16;
    w.rb1 = o; o += up256(nbits); w.rb2 = o; o += up256(nbits);
    w.bsc = o; o += up256(64 * 4);
  }
  w.scales = o; o += up256(tc::S_COUNT * 4);
  w.total = o; return w;
}
}  // namespace

// element offset of the Y part (channels C & ~7 .. C-1) in the channel-split fp16 copy of the 3DmFV tensor
long long tc_fv_y_off(const dpd_head_config& c) { return (long long)c.n_clouds * c.G * c.G * c.G * (c.C & ~7); }

size_t tc_packed_bytes(const dpd_head_config& c, bool f16) { return tc_blob_layout(c, f16).total; }
size_t tc_workspace_bytes(const dpd_head_config& c, bool f16, size_t rows) { return tc_ws_layout(c, f16, rows).total; }

int tc_pack_weights(const dpd_head_config& c, bool f16, int Kp1_src, const float* w1p, const float* w2, const float* w3,
                    const float* b1, const float* b2, void* tc_blob, cudaStream_t st) {
  const TcBlob b = tc_blob_layout(c, f16);
  char* base = (char*)tc_blob;
  const int H = c.H;
  dim3 blk(32, 8);
  if (!f16) {
    DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_kernel<<<dim3(ceil_div(H, 32), ceil_div(Kp1_src, 32)), blk, 0, st>>>(
        w1p, Kp1_src, H, (float*)(base + b.w1h), (float*)(base + b.w1l)));
    DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_kernel<<<dim3(ceil_div(H, 32), ceil_div(H, 32)), blk, 0, st>>>(
        w2, H, H, (float*)(base + b.w2h), (float*)(base + b.w2l)));
    DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_kernel<<<dim3(ceil_div(H, 32), ceil_div(H, 32)), blk, 0, st>>>(
        w3, H, H, (float*)(base + b.w3h), (float*)(base + b.w3l)));
    DPD_CUDA_CHECK_LAUNCH("transpose_split_kernel");
    return 0;
  }
  const int Kp1 = kp1_of(c, true);
  float* ps = (float*)(base + b.scales);
  unsigned* pb = (unsigned*)ps;
  DPD_CUDA_CALL(cudaMemsetAsync(ps, 0, tc::P_COUNT * 4, st));
  const size_t n1 = (size_t)Kp1_src * H, n2 = (size_t)H * H;
  DPD_LAUNCH("tc_pack_stats", st, tc::absmax_kernel<<<256, 256, 0, st>>>(w1p, n1, pb + tc::P_W1MAX_BITS));
  DPD_LAUNCH("tc_pack_stats", st, tc::absmax_kernel<<<256, 256, 0, st>>>(w2, n2, pb + tc::P_W2MAX_BITS));
  DPD_LAUNCH("tc_pack_stats", st, tc::absmax_kernel<<<256, 256, 0, st>>>(w3, n2, pb + tc::P_W3MAX_BITS));
  DPD_LAUNCH("tc_pack_stats", st, tc::absmax_kernel<<<4, 256, 0, st>>>(b1, (size_t)H, pb + tc::P_B1MAX_BITS));
  DPD_LAUNCH("tc_pack_stats", st, tc::absmax_kernel<<<4, 256, 0, st>>>(b2, (size_t)H, pb + tc::P_B2MAX_BITS));
  DPD_LAUNCH("tc_pack_stats", st, tc::col_l1_max_kernel<<<ceil_div(H, 32), 1024, 0, st>>>(w1p, Kp1_src, H, pb + tc::P_C1_BITS));
  DPD_LAUNCH("tc_pack_stats", st, tc::col_l1_max_kernel<<<ceil_div(H, 32), 1024, 0, st>>>(w2, H, H, pb + tc::P_C2_BITS));
  DPD_LAUNCH("tc_pack_stats", st, tc::weight_scales_kernel<<<1, 1, 0, st>>>(ps));
  // W1: K-major rows in the channel-split operand order of the gather producers
  DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_f16_kernel<<<dim3(ceil_div(H, 32), ceil_div(Kp1, 32)), blk, 0, st>>>(
      w1p, Kp1_src, Kp1, H, ps + tc::P_W1, (__half*)(base + b.w1h), (__half*)(base + b.w1l), c.k * c.k * c.k, c.C));
  DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_f16_kernel<<<dim3(ceil_div(H, 32), ceil_div(H, 32)), blk, 0, st>>>(
      w2, H, H, H, ps + tc::P_W2, (__half*)(base + b.w2h), (__half*)(base + b.w2l), 0, 0));
  DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_f16_kernel<<<dim3(ceil_div(H, 32), ceil_div(H, 32)), blk, 0, st>>>(
      w3, H, H, H, ps + tc::P_W3, (__half*)(base + b.w3h), (__half*)(base + b.w3l), 0, 0));
  if (tc_train(c) && (c.flags & DPD_HEAD_INPUT_GRAD)) {
    // rows [Kp1_src, Kp1) of the kernel-order operand are padding: zero them, split the rest in place
    DPD_CUDA_CALL(cudaMemsetAsync(base + b.w1nh, 0, (size_t)Kp1 * H * 2, st));
    DPD_CUDA_CALL(cudaMemsetAsync(base + b.w1nl, 0, (size_t)Kp1 * H * 2, st));
    DPD_LAUNCH("tc_pack_w", st, tc::split_f16_kernel<<<(unsigned)ceil_div<size_t>(n1, 256), 256, 0, st>>>(
        w1p, n1, ps + tc::P_W1, (__half*)(base + b.w1nh), (__half*)(base + b.w1nl)));
  }
  if (tc_train(c)) {
    const size_t n2e = (size_t)H * H;
    DPD_LAUNCH("tc_pack_w", st, tc::split_f16_kernel<<<(unsigned)ceil_div<size_t>(n2e, 256), 256, 0, st>>>(
        w2, n2e, ps + tc::P_W2, (__half*)(base + b.w2nh), (__half*)(base + b.w2nl)));
    DPD_LAUNCH("tc_pack_w", st, tc::split_f16_kernel<<<(unsigned)ceil_div<size_t>(n2e, 256), 256, 0, st>>>(
        w3, n2e, ps + tc::P_W3, (__half*)(base + b.w3nh), (__half*)(base + b.w3nl)));
  }
  DPD_CUDA_CHECK_LAUNCH("tc_pack_weights f16");
  return 0;
}

bool tc_backward_supported(const dpd_head_config& c, bool f16) { return f16 && tc::use_2cta() && tc_bwd_env() && tc_train(c); }

// One layer of the backward pass on the tensor cores (fp16x3, 2-CTA kernel), layer = 3, 2 or 1:
//   gw = A^T . dZ   with A = H2 (layer 3), H1 (layer 2) or the gathered layer-1 operand (layer 1); gb = column sums of dZ
//   dz_next = (dZ . W^T) * ReLU'(A)                                              (layers 3 and 2 only)
// dZ (fp32) is scaled by a power of two (its |.|max was recorded by the kernel that produced it) and split once into an
// fp16 (hi, lo) pair [rows, H].  Both products run through the forward GEMM kernel: dX with B = W as stored (K-major for
// this product) and the ReLU' bit-mask gate in the epilogue; dW with MN-major operands, i.e. A and dZ exactly as stored
// ([rows, features], the row axis being the reduction), split into slices of the active extent whose fp32 partials are
// summed in a fixed order.  No operand is ever transposed.
int tc_backward_layer(const dpd_head_config& c, int layer, const void* tc_blob, void* tc_ws, size_t ws_rows, int rows,
                      const GatherDesc* g, const float* dz, float* dz_next, const int* active, float* gw, float* gb,
                      cudaStream_t st) {
  const TcBlob b = tc_blob_layout(c, true);
  const TcWs w = tc_ws_layout(c, true, ws_rows);
  const char* blob = (const char*)tc_blob;
  char* ws = (char*)tc_ws;
  const int H = c.H, Kp1 = kp1_of(c, true);
  const int Mp = round_up(rows, 128);
  const int nblk = ceil_div(rows, 128);
  float* bsc = (float*)(ws + w.bsc);      // [0] -  [1] s_g  [2] 1/(s_g s_W)  [3] 1/(s_g s_A)  [8] active extent (int)  [16..19] |dZ|max slots
  int* extent = (int*)(bsc + 8);
  const float* ps = (const float*)(blob + b.scales);
  const float* sc = (const float*)(ws + w.scales);
  const float* act_scale = sc + (layer == 3 ? tc::S_A3 : layer == 2 ? tc::S_A2 : tc::S_A1);
  const float* w_scale = layer == 3 ? ps + tc::P_W3 : layer == 2 ? ps + tc::P_W2 : nullptr;
  __half* gh = (__half*)(ws + w.gh); __half* gl = (__half*)(ws + w.gl);
  unsigned* slots = (unsigned*)(bsc + 16);
  float* cpart = (float*)(ws + w.cpart);
  DPD_LAUNCH("bwd_scales", st, tc::bwd_scales_kernel<<<1, 1, 0, st>>>(bsc, w_scale, act_scale, slots + layer));
  DPD_LAUNCH("bwd_scales", st, tc::active_extent_kernel<<<1, 32, 0, st>>>(active, nblk, rows, extent));
  DPD_LAUNCH("bwd_split", st, tc::split_colsum_f16_kernel<<<dim3(Mp / 64, H / 64), 256, 0, st>>>(
      dz, rows, H, active, bsc + 1, gh, gl, extent, gw != nullptr ? cpart : nullptr));
  DPD_CUDA_CHECK_LAUNCH("tc_backward_layer prep");
  int rc;
  if (gw != nullptr) {
    const __half* ah; const __half* al;
    int Mo;       // rows of the weight gradient in kernel order
    if (layer == 1) {
      tc::GatherArgs ga;
      ga.rowinfo = nullptr; ga.lut = nullptr; ga.y_off = 0;
      ga.fv_hi = ws + w.fvh; ga.fv_lo = ws + w.fvl; ga.idx = g->idx; ga.off4_hi = ws + w.o4h; ga.off4_lo = ws + w.o4l; ga.row0 = g->row0;
      ga.y_off = tc_fv_y_off(c);
      ga.n_query = g->n_query; ga.G = g->G; ga.C = g->C; ga.k = g->k; ga.E = g->E;
      // the gather warps of the GEMM kernel assemble the MN-major A tiles on the fly: nothing is materialised
      int2* rinfo = (int2*)(ws + w.rinfo);
      DPD_LAUNCH("bwd_rowinfo", st, tc::rowinfo_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(
          g->idx, g->row0, g->n_query, g->G, g->C, g->k, rows, rinfo, (int2*)(ws + w.lut), Kp1 / 4, g->E));
      ga.lut = (const int2*)(ws + w.lut);
      DPD_CUDA_CHECK_LAUNCH("rowinfo_kernel");
      tc::BwdExtras bx;
      bx.mn_major = 1; bx.mn_rows = (unsigned long long)rows; ga.rowinfo = rinfo; bx.lut_chunks = Kp1 / 4;
      const int by_rows1 = Mp / (64 * 8) > 0 ? Mp / (64 * 8) : 1;
      bx.mode = 2; bx.slices = by_rows1 < TC_DW1_SLICES ? by_rows1 : TC_DW1_SLICES; bx.k_limit = extent;
      bx.slice_stride = (long long)Kp1 * H;
      float* part1 = (float*)(ws + w.part);
      if ((rc = tc::launch2(true, nullptr, nullptr, Kp1, Mp, gh, gl, H, nullptr, part1, nullptr, 0, bsc + 3, nullptr, &ga, st,
                            nullptr, nullptr, &bx))) return rc;
      DPD_LAUNCH("bwd_colsum", st, tc::colsum_blocks_final_kernel<<<H / 64, 1024, 0, st>>>(cpart, extent, H, gb));
      DPD_CUDA_CHECK_LAUNCH("tc_backward_layer colsum");
      return launch_reduce_partials(part1, nullptr, Kp1, g->E + 3, H, g->E, 1, gw, nullptr, st, bx.slices, g->k * g->k * g->k, g->C);
    } else {
      ah = (const __half*)(ws + (layer == 3 ? w.yh : w.xh)); al = (const __half*)(ws + (layer == 3 ? w.yl : w.xl)); Mo = H;
    }
    tc::BwdExtras bx;
    bx.mn_major = 1; bx.mn_rows = (unsigned long long)rows;
    // at most the configured number of slices, at least 8 K-blocks (512 rows) per slice: small batches need few partials
    const int max_slices = layer == 1 ? TC_DW1_SLICES : TC_DW_SLICES;
    const int by_rows = Mp / (64 * 8) > 0 ? Mp / (64 * 8) : 1;
    bx.mode = 2; bx.slices = by_rows < max_slices ? by_rows : max_slices; bx.k_limit = extent; bx.slice_stride = (long long)Mo * H;
    float* part = (float*)(ws + w.part);
    if ((rc = tc::launch2(false, ah, al, Mo, Mp, gh, gl, H, nullptr, part, nullptr, 0, bsc + 3, nullptr, nullptr, st, nullptr,
                          nullptr, &bx))) return rc;
    // gb = column sums of dZ over the active rows: per-block partials came out of the split pass
    DPD_LAUNCH("bwd_colsum", st, tc::colsum_blocks_final_kernel<<<H / 64, 1024, 0, st>>>(cpart, extent, H, gb));
    DPD_CUDA_CHECK_LAUNCH("tc_backward_layer colsum");
    if (layer == 1) rc = launch_reduce_partials(part, nullptr, Kp1, g->E + 3, H, g->E, 1, gw, nullptr, st, bx.slices, g->k * g->k * g->k, g->C);
    else rc = launch_reduce_partials(part, nullptr, H, H, H, 0, 0, gw, nullptr, st, bx.slices);
    if (rc) return rc;
  }
  if (dz_next != nullptr) {
    tc::BwdExtras bx;
    bx.mode = 1; bx.active = active; bx.absmax_bits = slots + (layer - 1);
    bx.relu_bits_in = (const uint4*)(ws + (layer == 3 ? w.rb2 : w.rb1));
    if ((rc = tc::launch2(false, gh, gl, rows, H, blob + (layer == 3 ? b.w3nh : b.w2nh), blob + (layer == 3 ? b.w3nl : b.w2nl), H,
                          nullptr, dz_next, nullptr, 0, bsc + 2, nullptr, nullptr, st, nullptr, nullptr, &bx))) return rc;
  }
  return 0;
}

int tc_kp1(const dpd_head_config& c) { return kp1_of(c, true); }

// |dZ|max slots of the tensor-core backward: zeroed at the start of a backward pass; slot 3 is filled by the layer-4
// backward kernel, slots 2 and 1 by the dX products
int tc_backward_begin(const dpd_head_config& c, void* tc_ws, size_t ws_rows, unsigned** slot3, cudaStream_t st) {
  const TcWs w = tc_ws_layout(c, true, ws_rows);
  float* bsc = (float*)((char*)tc_ws + w.bsc);
  DPD_CUDA_CALL(cudaMemsetAsync(bsc + 16, 0, 16, st));
  *slot3 = (unsigned*)(bsc + 16) + 3;
  return 0;
}

// dX1 = dZ1 . W1p^T for the input-gradient path.  prepare: measure / scale / split dZ1 once; rows: one group of rows.
int tc_backward_inputs_prepare(const dpd_head_config& c, const void* tc_blob, void* tc_ws, size_t ws_rows, int rows,
                               const float* dz1, const int* active, cudaStream_t st) {
  const TcBlob b = tc_blob_layout(c, true);
  const TcWs w = tc_ws_layout(c, true, ws_rows);
  char* ws = (char*)tc_ws;
  const int H = c.H, nblk = ceil_div(rows, 128);
  float* bsc = (float*)(ws + w.bsc);
  const float* ps = (const float*)((const char*)tc_blob + b.scales);
  DPD_LAUNCH("bwd_scales", st, tc::bwd_scales_kernel<<<1, 1, 0, st>>>(bsc, ps + tc::P_W1, ps + tc::P_W1, (const unsigned*)(bsc + 16) + 1));
  DPD_LAUNCH("bwd_split", st, tc::split_f16_active_kernel<<<nblk, 256, 0, st>>>(
      dz1, rows, H, active, bsc + 1, (__half*)(ws + w.gh), (__half*)(ws + w.gl)));
  DPD_CUDA_CHECK_LAUNCH("tc_backward_inputs_prepare");
  return 0;
}

int tc_backward_inputs_rows(const dpd_head_config& c, const void* tc_blob, void* tc_ws, size_t ws_rows, size_t r0, int nrows,
                            const int* active, float* dx1, cudaStream_t st) {
  const TcBlob b = tc_blob_layout(c, true);
  const TcWs w = tc_ws_layout(c, true, ws_rows);
  const char* blob = (const char*)tc_blob;
  char* ws = (char*)tc_ws;
  const int H = c.H, Kp1 = kp1_of(c, true);
  const float* bsc = (const float*)(ws + w.bsc);
  tc::BwdExtras bx;
  bx.mode = 2; bx.active = active + r0 / 128;
  return tc::launch2(false, (const __half*)(ws + w.gh) + r0 * H, (const __half*)(ws + w.gl) + r0 * H, nrows, H, blob + b.w1nh,
                     blob + b.w1nl, Kp1, nullptr, dx1, nullptr, 0, bsc + 2, nullptr, nullptr, st, nullptr, nullptr, &bx);
}

// mode 0: foreign fv (measure |fv|max, split)   1: |fv| <= 1 known (3DmFV output), split here
//      2: |fv| <= 1 and the FV kernel already wrote the (hi, lo) pair at tc_fv_split_ptrs with TC_FV_UNIT_SCALE
int tc_prepare_fv(const dpd_head_config& c, bool f16, const float* fv, const void* tc_blob, void* tc_ws, size_t ws_rows,
                  cudaStream_t st, int mode) {
  const TcWs w = tc_ws_layout(c, f16, ws_rows);
  char* ws = (char*)tc_ws;
  const size_t nfv = (size_t)c.n_clouds * c.G * c.G * c.G * c.C;
  if (!f16) {
    DPD_LAUNCH("tc_split_fv", st, tc::split_kernel<<<(unsigned)ceil_div<size_t>(nfv, 256), 256, 0, st>>>(
        fv, nfv, (float*)(ws + w.fvh), (float*)(ws + w.fvl)));
    DPD_CUDA_CHECK_LAUNCH("split_kernel");
    return 0;
  }
  const TcBlob b = tc_blob_layout(c, true);
  float* sc = (float*)(ws + w.scales);
  if (mode == 0) {
    DPD_CUDA_CALL(cudaMemsetAsync(sc, 0, tc::S_COUNT * 4, st));
    DPD_LAUNCH("tc_fv_absmax", st, tc::absmax_kernel<<<num_sms() * 4, 256, 0, st>>>(fv, nfv, (unsigned*)sc + tc::S_FVMAX_BITS));
  }
  // with the unit bound in1 = max(|fv|max, 1) = 1 exactly as the measured path would find for a 3DmFV tensor
  DPD_LAUNCH("tc_scales", st, tc::activation_scales_kernel<<<1, 1, 0, st>>>(sc, (const float*)((const char*)tc_blob + b.scales), mode != 0));
  if (mode != 2)
    DPD_LAUNCH("tc_split_fv", st, tc::split_fv_f16_kernel<<<(unsigned)ceil_div<size_t>(nfv, 256), 256, 0, st>>>(
        fv, nfv, c.C, tc_fv_y_off(c), sc + tc::S_A1, (__half*)(ws + w.fvh), (__half*)(ws + w.fvl)));
  DPD_CUDA_CHECK_LAUNCH("tc_prepare_fv f16");
  return 0;
}

void tc_fv_split_ptrs(const dpd_head_config& c, bool f16, void* tc_ws, size_t ws_rows, void** hi, void** lo, long long* y_off) {
  const TcWs w = tc_ws_layout(c, f16, ws_rows);
  *hi = (char*)tc_ws + w.fvh;
  *lo = (char*)tc_ws + w.fvl;
  *y_off = tc_fv_y_off(c);
}

int tc_head_layers(const dpd_head_config& c, bool f16, const GatherDesc& g, const float* mask, int rows, size_t ws_rows,
                   const void* tc_blob, const float* b1, const float* b2, const float* b3, float* ha, float* hb,
                   float* h3_out, void* tc_ws, const float** h3, cudaStream_t st, const float* w4, const float* b4,
                   float* fused_out) {
  const TcBlob b = tc_blob_layout(c, f16);
  const TcWs w = tc_ws_layout(c, f16, ws_rows);
  const char* blob = (const char*)tc_blob;
  char* ws = (char*)tc_ws;
  const int H = c.H, Kp1 = kp1_of(c, f16);
  const float* sc = (const float*)(ws + w.scales);
  tc::GatherArgs ga;
  ga.rowinfo = nullptr; ga.lut = nullptr; ga.y_off = 0;
  ga.fv_hi = ws + w.fvh; ga.fv_lo = ws + w.fvl; ga.idx = g.idx; ga.off4_hi = ws + w.o4h; ga.off4_lo = ws + w.o4l; ga.row0 = g.row0;
  ga.y_off = f16 ? tc_fv_y_off(c) : 0;
  ga.n_query = g.n_query; ga.G = g.G; ga.C = g.C; ga.k = g.k; ga.E = g.E;
  float* out3 = h3_out ? h3_out : ha;
  int rc;
  if (!f16) {
    DPD_LAUNCH("tc_split_off", st, tc::split_off4_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(g.offset, rows, (float*)(ws + w.o4h), (float*)(ws + w.o4l)));
    DPD_CUDA_CHECK_LAUNCH("split_off4_kernel");
    // layer 1: gathered A -> (ha = hi, hb = lo); layer 2 -> (xh, xl); layer 3 -> fp32
    if ((rc = tc::launch(true, false, nullptr, nullptr, rows, Kp1, blob + b.w1h, blob + b.w1l, H, b1, ha, hb, 1, nullptr, nullptr, &ga, st))) return rc;
    if ((rc = tc::launch(false, false, ha, hb, rows, H, blob + b.w2h, blob + b.w2l, H, b2, ws + w.xh, ws + w.xl, 1, nullptr, nullptr, nullptr, st))) return rc;
    if ((rc = tc::launch(false, false, ws + w.xh, ws + w.xl, rows, H, blob + b.w3h, blob + b.w3l, H, b3, out3, nullptr, 0, nullptr, nullptr, nullptr, st))) return rc;
  } else {
    DPD_LAUNCH("tc_split_off", st, tc::split_off4_f16_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(
        g.offset, mask, rows, sc + tc::S_A1, (__half*)(ws + w.o4h), (__half*)(ws + w.o4l),
        g.idx, g.row0, g.n_query, g.G, g.C, g.k, (int2*)(ws + w.rinfo), (int2*)(ws + w.lut), Kp1 / 4, g.E));
    DPD_CUDA_CHECK_LAUNCH("split_off4_f16_kernel");
    ga.rowinfo = (const int2*)(ws + w.rinfo); ga.lut = (const int2*)(ws + w.lut);
    const bool fused = fused_out != nullptr && tc::use_2cta() && tc::fuse_l4();
    // layer 1: gathered A -> (xh, xl) scaled by sA2; layer 2 -> (yh, yl) scaled by sA3; layer 3 -> fp32
    const bool bits = tc_train(c) && tc::use_2cta();    // ReLU' bit masks for the tensor-core backward
    if ((rc = tc::launch(true, true, nullptr, nullptr, rows, Kp1, blob + b.w1h, blob + b.w1l, H, b1, ws + w.xh, ws + w.xl, 1,
                         sc + tc::S_ACC1, sc + tc::S_A2, &ga, st, bits ? (uint4*)(ws + w.rb1) : nullptr, !tc_train(c)))) return rc;
    if ((rc = tc::launch(false, true, ws + w.xh, ws + w.xl, rows, H, blob + b.w2h, blob + b.w2l, H, b2, ws + w.yh, ws + w.yl, 1,
                         sc + tc::S_ACC2, sc + tc::S_A3, nullptr, st, bits ? (uint4*)(ws + w.rb2) : nullptr, !tc_train(c)))) return rc;
    if (fused) {
      // layer 3 with the output layer fused into its epilogue.  Inference: H3 never reaches HBM.  Training (h3_out set):
      // the fp32 activations are stored as well, for the backward pass.  `ha` (not read by this path; the SIMT backward
      // refills it later) holds the [rows, 2*H/256] float4 partials.
      const int nslots = 2 * (H / tc::BN);
      float* part4 = ha;
      if ((rc = tc::launch2(false, ws + w.yh, ws + w.yl, rows, H, blob + b.w3h, blob + b.w3l, H, b3, h3_out, nullptr, 0,
                            sc + tc::S_ACC3, nullptr, nullptr, st, w4, part4, nullptr, nullptr, !tc_train(c)))) return rc;
      DPD_LAUNCH("head_out_finish", st, tc::head_out_finish_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(
          (const float4*)part4, nslots, b4, mask, fused_out, rows));
      DPD_CUDA_CHECK_LAUNCH("head_out_finish_kernel");
      *h3 = nullptr;
      return 0;
    }
    if ((rc = tc::launch(false, true, ws + w.yh, ws + w.yl, rows, H, blob + b.w3h, blob + b.w3l, H, b3, out3, nullptr, 0,
                         sc + tc::S_ACC3, nullptr, nullptr, st))) return rc;
  }
  (void)hb;
  *h3 = out3;
  return 0;
}

// backward support: fp32 H1 -> ha, H2 -> hb from the (hi, lo) pairs the forward left behind
int tc_merge_activations(const dpd_head_config& c, bool f16, void* tc_ws, size_t ws_rows, int rows, float* ha, float* hb,
                         cudaStream_t st) {
  const TcWs w = tc_ws_layout(c, f16, ws_rows);
  char* ws = (char*)tc_ws;
  const size_t n = (size_t)rows * c.H;
  int rc;
  if (!f16) {   // H1 = (ha, hb), H2 = (xh, xl): H1 into ha, then H2 into hb
    if ((rc = launch_add_inplace(ha, hb, n, st))) return rc;
    DPD_CUDA_CALL(cudaMemcpyAsync(hb, ws + w.xh, n * 4, cudaMemcpyDeviceToDevice, st));
    return launch_add_inplace(hb, (const float*)(ws + w.xl), n, st);
  }
  const float* sc = (const float*)(ws + w.scales);
  DPD_LAUNCH("bwd_merge_hi_lo", st, tc::merge_f16_kernel<<<(unsigned)ceil_div<size_t>(n, 256), 256, 0, st>>>(
      (const __half*)(ws + w.xh), (const __half*)(ws + w.xl), sc + tc::S_A2, ha, n));
  DPD_LAUNCH("bwd_merge_hi_lo", st, tc::merge_f16_kernel<<<(unsigned)ceil_div<size_t>(n, 256), 256, 0, st>>>(
      (const __half*)(ws + w.yh), (const __half*)(ws + w.yl), sc + tc::S_A3, hb, n));
  DPD_CUDA_CHECK_LAUNCH("merge_f16_kernel");
  return 0;
}

// test hook: dense split-precision GEMM on caller-provided fp32 operands (see include/dpdist_b200.h)
int tc_debug_gemm(const float* a, int M, int K, const float* w, int N, const float* bias, float* out, void* scratch,
                  size_t scratch_bytes, int f16, cudaStream_t st) {
  const size_t need = ((size_t)M * K * 2 + (size_t)N * K * 2) * 4 + 256;
  DPD_REQUIRE(scratch_bytes >= need, DPD_E_WORKSPACE, "tc_debug_gemm: scratch %zu < %zu", scratch_bytes, need);
  char* s = (char*)scratch;
  if (!f16) {
    float* ah = (float*)s; float* al = ah + (size_t)M * K;
    float* bh = al + (size_t)M * K; float* bl = bh + (size_t)N * K;
    DPD_LAUNCH("tc_split_fv", st, tc::split_kernel<<<(unsigned)ceil_div<size_t>((size_t)M * K, 256), 256, 0, st>>>(a, (size_t)M * K, ah, al));
    DPD_LAUNCH("tc_pack_w", st, tc::transpose_split_kernel<<<dim3(ceil_div(N, 32), ceil_div(K, 32)), dim3(32, 8), 0, st>>>(w, K, N, bh, bl));
    DPD_CUDA_CHECK_LAUNCH("tc_debug_gemm prep");
    return tc::launch(false, false, ah, al, M, K, bh, bl, N, bias, out, nullptr, 0, nullptr, nullptr, nullptr, st);
  }
  // fp16x3 with scales from the measured maxima of both operands
  float* sc = (float*)s;                      // [0] sA  [1] sW  [2] 1/(sA*sW)  [4] |a|max bits  [5] |w|max bits
  __half* ah = (__half*)(s + 256); __half* al = ah + (size_t)M * K;
  __half* bh = al + (size_t)M * K; __half* bl = bh + (size_t)N * K;
  DPD_CUDA_CALL(cudaMemsetAsync(sc, 0, 64, st));
  tc::absmax_kernel<<<256, 256, 0, st>>>(a, (size_t)M * K, (unsigned*)sc + 4);
  tc::absmax_kernel<<<256, 256, 0, st>>>(w, (size_t)N * K, (unsigned*)sc + 5);
  tc::debug_scales_kernel<<<1, 1, 0, st>>>(sc);
  tc::split_f16_kernel<<<(unsigned)ceil_div<size_t>((size_t)M * K, 256), 256, 0, st>>>(a, (size_t)M * K, sc + 0, ah, al);
  tc::transpose_split_f16_kernel<<<dim3(ceil_div(N, 32), ceil_div(K, 32)), dim3(32, 8), 0, st>>>(w, K, K, N, sc + 1, bh, bl, 0, 0);
  DPD_CUDA_CHECK_LAUNCH("tc_debug_gemm f16 prep");
  return tc::launch(false, true, ah, al, M, K, bh, bl, N, bias, out, nullptr, 0, sc + 2, nullptr, nullptr, st);
}

}  // namespace dpd
