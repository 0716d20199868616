// Gradient of the 3DmFV encoding w.r.t. the points: what tf.gradients builds over get_3dmfv_tf
// (reference utils/dpdist_util.py:54-137) when DPDist is used as a loss for another network
// (pcrnet-registration/iterative_PCRNet_ours.py:229-257; train_multi_gpu_pc_compare_dist.py:433-463).
//
// One 256-thread CTA per cloud, nothing but the points and the upstream gradient is read:
//   phase A  (thread <-> Gaussian, loop over points) recomputes the forward statistics: per channel the
//            sum, max, min over the points, and HOW MANY points attain each max / min (TF's reduce_max /
//            reduce_min gradient is split evenly among ties)
//   phase A2 back-propagates the per-channel L2 normalisation (:124-126) and the signed square root
//            (:118-121) to the raw statistics; leaves, per Gaussian, one coefficient per statistic
//   phase B  (thread <-> point, 4 threads per point, loop over Gaussians) evaluates the pair values again with
//            the SAME instruction sequence (so equality with the recorded max / min is exact), routes the
//            coefficients to the pairs and pushes them through Q = softmax_g(-|z|^2 / 2), z = (x - mu) / sigma.
// With a0, am_d, as_d the coefficients of Q, Q z_d, Q (z_d^2 - 1) at pair (n, g):
//   gQ = a0 + sum_d am_d z_d + as_d (z_d^2 - 1),    S_n = sum_g Q gQ
//   dL/dx_d[n] = (1/sigma) sum_g Q (am_d + 2 as_d z_d - (gQ - S_n) z_d)
// Deterministic: all reductions run in a fixed order.
#include "fv.cuh"

namespace dpd {
namespace {

constexpr int BT = 256;        // threads per CTA
constexpr int BP = 64;         // points per table chunk
constexpr int CS = 48;         // floats per Gaussian in the coefficient table
// final layout of a Gaussian's row: cm[7] @0 (mean coefficients), cx[7] @8, vx[7] @16 (max: coefficient / tie count, value),
// cn[6] @24, vn[6] @32 (min).  While the channel norms are being reduced the row holds the raw statistics instead:
// r[20] @0, vx[7] @20, vn[6] @27, tie counts of the maxima [7] @33 and of the minima [6] @40.

struct PairValues { float v[7]; };   // Q, Q zx, Q zy, Q zz, Q (zx^2-1), Q (zy^2-1), Q (zz^2-1)

// pinned instruction sequence (no contraction differences between the two phases)
__device__ __forceinline__ PairValues pair_values(float qy, float qx, float qz, float zx, float zy, float zz) {
  PairValues r;
  const float Q = __fmul_rn(__fmul_rn(qy, qx), qz);
  r.v[0] = Q;
  r.v[1] = __fmul_rn(Q, zx); r.v[2] = __fmul_rn(Q, zy); r.v[3] = __fmul_rn(Q, zz);
  r.v[4] = __fmul_rn(Q, __fmaf_rn(zx, zx, -1.0f));
  r.v[5] = __fmul_rn(Q, __fmaf_rn(zy, zy, -1.0f));
  r.v[6] = __fmul_rn(Q, __fmaf_rn(zz, zz, -1.0f));
  return r;
}

struct BwdParams {
  const float* points;    // [n_clouds, N, 3]
  const float* grad_fv;   // [n_clouds, V, C] or [n_clouds, C*V]
  float* grad_points;     // [n_clouds, N, 3]
  int n_clouds, N, G, V, C, flatten;
  float sigma;
  float c[DPD_MAX_GRID];
};

// tables of one 64-point chunk: tq[axis][point][cell], tz[axis][point][cell]; axis 0 = x <-> i1, 1 = y <-> i0, 2 = z <-> i2
__device__ __forceinline__ void build_tables(const BwdParams& p, const float* pts, int n0, int np, float* tq, float* tz) {
  const int G = p.G;
  for (int t = threadIdx.x; t < np * 3; t += BT) {
    const int pi = t / 3, a = t - pi * 3;
    const float x = pts[(size_t)(n0 + pi) * 3 + a];
    float* q = tq + (a * BP + pi) * G;
    float* z = tz + (a * BP + pi) * G;
    float hmin = INFINITY;             // softmax shift, as in the forward kernels (no 0/0 for far points)
    for (int i = 0; i < G; ++i) {
      const float zz = (x - p.c[i]) / p.sigma;
      hmin = fminf(hmin, zz * zz);
    }
    float sum = 0.f;
    for (int i = 0; i < G; ++i) {
      const float zz = (x - p.c[i]) / p.sigma;
      const float e = expf(-0.5f * (zz * zz - hmin));
      z[i] = zz; q[i] = e; sum += e;
    }
    const float inv = 1.0f / sum;
    for (int i = 0; i < G; ++i) q[i] *= inv;
  }
}

template <bool FULL>
__global__ void __launch_bounds__(BT) fv_backward_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int G = p.G, V = p.V, N = p.N, C = p.C;
  float* coef = smem;                         // [V][CS]
  float* tq = coef + (size_t)V * CS;          // [3][BP][G]
  float* tz = tq + 3 * BP * G;                // [3][BP][G]
  float* red = tz + 3 * BP * G;               // [8 warps][2*20]
  float* chan = red + 8 * 40;                 // [20] rs, [20] rs^3 * dot (0 where the norm is clamped)
  float* comb = chan + 40;                    // [4][BP][7] per-part sums of phase B
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NCH = FULL ? 20 : 7;

  const float w = 1.0f / (float)V;
  const float sqrt_w = sqrtf(w);
  const float c_pi = 1.0f / (sqrt_w * (float)N), c_mu = 1.0f / sqrt_w, c_sg = 1.0f / sqrtf(2.0f * w);
  const float inv_n = 1.0f / (float)N;

  for (int cloud = blockIdx.x; cloud < p.n_clouds; cloud += gridDim.x) {
    const float* pts = p.points + (size_t)cloud * N * 3;
    const float* gup = p.grad_fv + (size_t)cloud * V * C;
    float ss_acc[NCH], dot_acc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) ss_acc[c] = dot_acc[c] = 0.f;

    // ---------------- phase A: forward statistics, thread <-> Gaussian ----------------
    for (int g0 = 0; g0 < V; g0 += BT) {
      const int g = g0 + tid;
      const bool active = g < V;
      const int i2 = g % G, i1 = (g / G) % G, i0 = (g / (G * G)) % G;
      float sum[7], mx[7], mn[7], cx[7], cn[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) { sum[j] = 0.f; mx[j] = -INFINITY; mn[j] = INFINITY; cx[j] = cn[j] = 0.f; }
      for (int n0 = 0; n0 < N; n0 += BP) {
        const int np = min(BP, N - n0);
        __syncthreads();
        build_tables(p, pts, n0, np, tq, tz);
        __syncthreads();
        if (active) {
          for (int pi = 0; pi < np; ++pi) {
            const PairValues pv = pair_values(tq[(1 * BP + pi) * G + i0], tq[(0 * BP + pi) * G + i1], tq[(2 * BP + pi) * G + i2],
                                              tz[(0 * BP + pi) * G + i1], tz[(1 * BP + pi) * G + i0], tz[(2 * BP + pi) * G + i2]);
#pragma unroll
            for (int j = 0; j < 7; ++j) {
              const float v = pv.v[j];
              sum[j] += v;
              if (FULL) {
                if (v > mx[j]) { mx[j] = v; cx[j] = 1.f; } else if (v == mx[j]) cx[j] += 1.f;
                if (j > 0) { if (v < mn[j]) { mn[j] = v; cn[j] = 1.f; } else if (v == mn[j]) cn[j] += 1.f; }
              }
            }
          }
        }
      }
      if (active) {
        // raw (scaled) statistics in the output channel order (:134-137)
        float r[NCH];
        if (FULL) {
          r[0] = (sum[0] * inv_n - w) * c_pi; r[1] = (mx[0] - w) * c_pi;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            r[2 + d] = sum[1 + d] * inv_n * c_mu; r[5 + d] = mx[1 + d] * c_mu; r[8 + d] = mn[1 + d] * c_mu;
            r[11 + d] = sum[4 + d] * inv_n * c_sg; r[14 + d] = mx[4 + d] * c_sg; r[17 + d] = mn[4 + d] * c_sg;
          }
        } else {
          r[0] = (sum[0] * inv_n - w) * c_pi;
#pragma unroll
          for (int d = 0; d < 3; ++d) { r[1 + d] = sum[1 + d] * inv_n * c_mu; r[4 + d] = sum[4 + d] * inv_n * c_sg; }
        }
        float* row = coef + (size_t)g * CS;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          row[c] = r[c];
          const float t = power_norm(r[c]);
          const float gu = p.flatten ? gup[(size_t)c * V + g] : gup[(size_t)g * C + c];
          ss_acc[c] = fmaf(t, t, ss_acc[c]);
          dot_acc[c] = fmaf(gu, t, dot_acc[c]);
        }
        if (FULL) {
#pragma unroll
          for (int j = 0; j < 7; ++j) { row[20 + j] = mx[j]; row[33 + j] = cx[j]; }
#pragma unroll
          for (int j = 0; j < 6; ++j) { row[27 + j] = mn[1 + j]; row[40 + j] = cn[1 + j]; }
        }
      }
    }
    // ---------------- channel reductions over the Gaussians (fixed order) ----------------
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float a = ss_acc[c], b = dot_acc[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if (lane == 0) { red[warp * 40 + c] = a; red[warp * 40 + 20 + c] = b; }
    }
    __syncthreads();
    if (tid < NCH) {
      float ss = 0.f, dot = 0.f;
      for (int wv = 0; wv < BT / 32; ++wv) { ss += red[wv * 40 + tid]; dot += red[wv * 40 + 20 + tid]; }
      // tf.nn.l2_normalize: y = t * rsqrt(max(ss, 1e-12)); below the clamp the norm is a constant
      const float rs = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
      chan[tid] = rs;
      chan[20 + tid] = (ss >= 1e-12f) ? rs * rs * rs * dot : 0.f;
    }
    __syncthreads();
    // ---------------- phase A2: d(raw statistic), then per-pair coefficients ----------------
    for (int g = tid; g < V; g += BT) {
      float* row = coef + (size_t)g * CS;
      float dr[NCH];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float r = row[c];
        const float t = power_norm(r);
        const float gu = p.flatten ? gup[(size_t)c * V + g] : gup[(size_t)g * C + c];
        const float dt = chan[c] * gu - t * chan[20 + c];
        // d/dr sign(r) sqrt(max(|r|, eps)) = 0.5 / sqrt(|r|) where |r| >= eps, else 0
        dr[c] = (fabsf(r) >= 1e-12f) ? dt * 0.5f / fabsf(t) : 0.f;
      }
      if (FULL) {
        float vx[7], vn[6], kx[7], kn[6];
#pragma unroll
        for (int j = 0; j < 7; ++j) { vx[j] = row[20 + j]; kx[j] = row[33 + j]; }
#pragma unroll
        for (int j = 0; j < 6; ++j) { vn[j] = row[27 + j]; kn[j] = row[40 + j]; }
        row[0] = dr[0] * c_pi * inv_n;
        row[8] = dr[1] * c_pi / kx[0]; row[16] = vx[0];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          row[1 + d] = dr[2 + d] * c_mu * inv_n;
          row[9 + d] = dr[5 + d] * c_mu / kx[1 + d];   row[17 + d] = vx[1 + d];
          row[24 + d] = dr[8 + d] * c_mu / kn[d];      row[32 + d] = vn[d];
          row[4 + d] = dr[11 + d] * c_sg * inv_n;
          row[12 + d] = dr[14 + d] * c_sg / kx[4 + d]; row[20 + d] = vx[4 + d];
          row[27 + d] = dr[17 + d] * c_sg / kn[3 + d]; row[35 + d] = vn[3 + d];
        }
      } else {
        row[0] = dr[0] * c_pi * inv_n;
#pragma unroll
        for (int d = 0; d < 3; ++d) { row[1 + d] = dr[1 + d] * c_mu * inv_n; row[4 + d] = dr[4 + d] * c_sg * inv_n; }
      }
    }
    // ---------------- phase B: thread <-> (point, quarter of the Gaussians) ----------------
    const int part = tid >> 6, pl = tid & 63;
    for (int n0 = 0; n0 < N; n0 += BP) {
      const int np = min(BP, N - n0);
      __syncthreads();
      build_tables(p, pts, n0, np, tq, tz);
      __syncthreads();
      float S = 0.f, T[3] = {0.f, 0.f, 0.f}, U[3] = {0.f, 0.f, 0.f};
      if (pl < np) {
        for (int g = part; g < V; g += 4) {
          const int i2 = g % G, i1 = (g / G) % G, i0 = g / (G * G);
          const float z[3] = {tz[(0 * BP + pl) * G + i1], tz[(1 * BP + pl) * G + i0], tz[(2 * BP + pl) * G + i2]};
          const PairValues pv = pair_values(tq[(1 * BP + pl) * G + i0], tq[(0 * BP + pl) * G + i1], tq[(2 * BP + pl) * G + i2],
                                            z[0], z[1], z[2]);
          const float* row = coef + (size_t)g * CS;
          float a[7];
          {
            const float4 m0 = *reinterpret_cast<const float4*>(row), m1 = *reinterpret_cast<const float4*>(row + 4);
            a[0] = m0.x; a[1] = m0.y; a[2] = m0.z; a[3] = m0.w; a[4] = m1.x; a[5] = m1.y; a[6] = m1.z;
          }
          if (FULL) {
            const float4 c0 = *reinterpret_cast<const float4*>(row + 8), c1 = *reinterpret_cast<const float4*>(row + 12);
            const float4 v0 = *reinterpret_cast<const float4*>(row + 16), v1 = *reinterpret_cast<const float4*>(row + 20);
            const float cxs[7] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z};
            const float vxs[7] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z};
#pragma unroll
            for (int j = 0; j < 7; ++j) if (pv.v[j] == vxs[j]) a[j] += cxs[j];
            const float4 d0 = *reinterpret_cast<const float4*>(row + 24), d1 = *reinterpret_cast<const float4*>(row + 28);
            const float4 w0 = *reinterpret_cast<const float4*>(row + 32), w1 = *reinterpret_cast<const float4*>(row + 36);
            const float cns[6] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y};
            const float vns[6] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y};
#pragma unroll
            for (int j = 0; j < 6; ++j) if (pv.v[1 + j] == vns[j]) a[1 + j] += cns[j];
          }
          const float Q = pv.v[0];
          float gQ = a[0];
#pragma unroll
          for (int d = 0; d < 3; ++d) gQ += a[1 + d] * z[d] + a[4 + d] * (z[d] * z[d] - 1.0f);
          S = fmaf(Q, gQ, S);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            T[d] = fmaf(Q, a[1 + d] + 2.0f * a[4 + d] * z[d] - gQ * z[d], T[d]);
            U[d] = fmaf(Q, z[d], U[d]);
          }
        }
      }
      // combine the four parts of a point in a fixed order through shared memory
      if (pl < np) {
        float* dst = comb + (part * BP + pl) * 7;
        dst[0] = S; dst[1] = T[0]; dst[2] = T[1]; dst[3] = T[2]; dst[4] = U[0]; dst[5] = U[1]; dst[6] = U[2];
      }
      __syncthreads();
      if (tid < np) {
        float acc[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[j] = 0.f;
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 7; ++j) acc[j] += comb[(q * BP + tid) * 7 + j];
        float* out = p.grad_points + ((size_t)cloud * N + n0 + tid) * 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) out[d] = (acc[1 + d] + acc[0] * acc[4 + d]) / p.sigma;
      }
    }
    __syncthreads();   // coef / tables are rewritten by the next cloud
  }
}

}  // namespace
}  // namespace dpd

extern "C" int dpd_fv_backward(const float* d_points, int n_clouds, int n_points, int G, const float* h_centers,
                               float sigma, int full_fv, int flatten, const float* d_grad_fv, float* d_grad_points,
                               void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_points && d_grad_fv && d_grad_points && h_centers, DPD_E_INVALID, "dpd_fv_backward: null pointer");
  DPD_REQUIRE(n_clouds >= 0 && n_points > 0, DPD_E_INVALID, "dpd_fv_backward: bad sizes (%d clouds, %d points)", n_clouds, n_points);
  DPD_REQUIRE(G >= 2 && G <= DPD_MAX_GRID, DPD_E_UNSUPPORTED, "dpd_fv_backward: G=%d outside [2,%d]", G, DPD_MAX_GRID);
  DPD_REQUIRE(sigma > 0.f, DPD_E_INVALID, "dpd_fv_backward: sigma must be > 0");
  if (n_clouds == 0) return 0;
  BwdParams p;
  p.points = d_points; p.grad_fv = d_grad_fv; p.grad_points = d_grad_points;
  p.n_clouds = n_clouds; p.N = n_points; p.G = G; p.V = G * G * G; p.flatten = flatten ? 1 : 0;
  p.C = full_fv ? DPD_FV_CHANNELS_FULL : DPD_FV_CHANNELS_SMALL;
  p.sigma = sigma;
  for (int i = 0; i < DPD_MAX_GRID; ++i) p.c[i] = i < G ? h_centers[i] : 0.f;
  const size_t smem = ((size_t)p.V * CS + 6 * BP * G + 8 * 40 + 40 + 4 * BP * 7) * sizeof(float);
  DPD_REQUIRE(smem <= 227 * 1024, DPD_E_UNSUPPORTED,
              "dpd_fv_backward: G=%d needs %zu B of shared memory per CTA (limit 227 KB, G <= 10)", G, smem);
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    DPD_CUDA_CALL(cudaFuncSetAttribute(fv_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DPD_CUDA_CALL(cudaFuncSetAttribute(fv_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int grid = n_clouds < 4 * num_sms() ? n_clouds : 4 * num_sms();
  if (full_fv) DPD_LAUNCH("fv_backward", st, fv_backward_kernel<true><<<grid, BT, smem, st>>>(p));
  else DPD_LAUNCH("fv_backward", st, fv_backward_kernel<false><<<grid, BT, smem, st>>>(p));
  DPD_CUDA_CHECK_LAUNCH("fv_backward_kernel");
  return 0;
}
