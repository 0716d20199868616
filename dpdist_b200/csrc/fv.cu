// 3DmFV encoding kernels (replaces get_3dmfv_tf, reference utils/dpdist_util.py:22-141).
//
// Arithmetic: the reference GMM is an axis-aligned G^3 grid with one isotropic sigma and equal
// weights, so the responsibility factorises per axis (oracle/fv_separable_np.py):
//   Q[n,(i0,i1,i2)] = qy[n,i0]*qx[n,i1]*qz[n,i2],  q_a[n,i] = e_a[n,i]/sum_i e_a[n,i],
//   e_a[n,i] = exp(-0.5 z^2), z = (p_a - l[i])/sigma,  m = q*z (d/dmu), s = q*(z^2-1) (d/dsigma)
// Per point only 3*G exps are evaluated; the N*G^3 pair work is multiplies + sum/max/min.
//
// fv_generic_kernel: any G <= 16, any N, full or small FV.  One CTA per cloud.
#include "fv.cuh"
#include <stdlib.h>

namespace dpd {

constexpr int FV_THREADS = 256;
constexpr int FV_PCHUNK = 32;  // points per smem table chunk (generic kernel)

__global__ void __launch_bounds__(FV_THREADS) fv_generic_kernel(const FvParams p) {
  extern __shared__ float smem[];
  const int G = p.G, V = p.V, N = p.N, C = p.C;
  // tables: [axis(3)][type(3: q,m,s)][FV_PCHUNK][G]
  float* tab = smem;
  float* chan_ss = smem + 9 * FV_PCHUNK * G;  // [C] sum of squares per channel
  float* warp_ss = chan_ss + DPD_FV_CHANNELS_FULL;  // [FV_THREADS/32][C] per-warp partials (fixed-order sum)
  const int cloud = blockIdx.x;
  const float* pts = p.points + (size_t)cloud * N * 3;
  float* out = p.fv + (size_t)cloud * V * C;
  const int tid = threadIdx.x;

  const float w = 1.0f / (float)V;            // tf.ones/n_gaussians (:49)
  const float sqrt_w = sqrtf(w);
  const float c_pi = 1.0f / (sqrt_w * (float)N);   // (:78)
  const float c_mu = 1.0f / sqrt_w;                // (:98)
  const float c_sg = 1.0f / sqrtf(2.0f * w);       // (:109)
  const float inv_n = 1.0f / (float)N;

  for (int i = tid; i < C; i += FV_THREADS) chan_ss[i] = 0.f;

  for (int g0 = 0; g0 < V; g0 += FV_THREADS) {
    const int g = g0 + tid;
    const bool active = g < V;
    const int i2 = g % G, i1 = (g / G) % G, i0 = g / (G * G);
    float sQ = 0.f, mQ = -INFINITY;
    float sm[3] = {0.f, 0.f, 0.f}, ss[3] = {0.f, 0.f, 0.f};
    float xm[3], nm[3], xs[3], ns[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { xm[d] = xs[d] = -INFINITY; nm[d] = ns[d] = INFINITY; }

    for (int n0 = 0; n0 < N; n0 += FV_PCHUNK) {
      const int np = min(FV_PCHUNK, N - n0);
      __syncthreads();
      // build the per-axis tables for this chunk: one thread per (point, axis)
      for (int t = tid; t < np * 3; t += FV_THREADS) {
        const int pi = t / 3, a = t % 3;
        const float x = pts[(size_t)(n0 + pi) * 3 + a];
        float* q = tab + ((a * 3 + 0) * FV_PCHUNK + pi) * G;
        float* m = tab + ((a * 3 + 1) * FV_PCHUNK + pi) * G;
        float* s = tab + ((a * 3 + 2) * FV_PCHUNK + pi) * G;
        // exponentials shifted by the axis' smallest exponent (softmax shift: same q wherever the reference is finite, the
        // limit value instead of 0/0 = NaN for a point far outside the cube; see fv_ws.cu)
        float hmin = INFINITY;
        for (int i = 0; i < G; ++i) {
          const float z = (x - p.c[i]) / p.sigma;
          hmin = fminf(hmin, z * z);
        }
        float sum = 0.f;
        for (int i = 0; i < G; ++i) {
          const float z = (x - p.c[i]) / p.sigma;
          const float e = expf(-0.5f * (z * z - hmin));
          q[i] = e;
          sum += e;
        }
        const float inv = 1.0f / sum;
        for (int i = 0; i < G; ++i) {
          const float z = (x - p.c[i]) / p.sigma;
          const float qq = q[i] * inv;
          q[i] = qq;
          m[i] = qq * z;
          s[i] = qq * (z * z - 1.0f);
        }
      }
      __syncthreads();
      if (active) {
        // axis 0 = x <-> i1, axis 1 = y <-> i0, axis 2 = z <-> i2
        for (int pi = 0; pi < np; ++pi) {
          const float qx = tab[((0 * 3 + 0) * FV_PCHUNK + pi) * G + i1];
          const float mx = tab[((0 * 3 + 1) * FV_PCHUNK + pi) * G + i1];
          const float sx = tab[((0 * 3 + 2) * FV_PCHUNK + pi) * G + i1];
          const float qy = tab[((1 * 3 + 0) * FV_PCHUNK + pi) * G + i0];
          const float my = tab[((1 * 3 + 1) * FV_PCHUNK + pi) * G + i0];
          const float sy = tab[((1 * 3 + 2) * FV_PCHUNK + pi) * G + i0];
          const float qz = tab[((2 * 3 + 0) * FV_PCHUNK + pi) * G + i2];
          const float mz = tab[((2 * 3 + 1) * FV_PCHUNK + pi) * G + i2];
          const float sz = tab[((2 * 3 + 2) * FV_PCHUNK + pi) * G + i2];
          const float qyx = qy * qx, qyz = qy * qz, qxz = qx * qz;
          const float Q = qyx * qz;
          sQ += Q;
          mQ = fmaxf(mQ, Q);
          const float v[6] = {mx * qyz, my * qxz, mz * qyx, sx * qyz, sy * qxz, sz * qyx};
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            sm[d] += v[d];
            xm[d] = fmaxf(xm[d], v[d]);
            nm[d] = fminf(nm[d], v[d]);
            ss[d] += v[3 + d];
            xs[d] = fmaxf(xs[d], v[3 + d]);
            ns[d] = fminf(ns[d], v[3 + d]);
          }
        }
      }
    }
    // scale + power-normalise, park un-normalised values in the output buffer
    float vals[DPD_FV_CHANNELS_FULL];
    int nc = 0;
    vals[nc++] = (sQ * inv_n - w) * c_pi;
    if (p.full_fv) vals[nc++] = (mQ - w) * c_pi;
#pragma unroll
    for (int d = 0; d < 3; ++d) vals[nc++] = sm[d] * inv_n * c_mu;
    if (p.full_fv) {
#pragma unroll
      for (int d = 0; d < 3; ++d) vals[nc++] = xm[d] * c_mu;
#pragma unroll
      for (int d = 0; d < 3; ++d) vals[nc++] = nm[d] * c_mu;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) vals[nc++] = ss[d] * inv_n * c_sg;
    if (p.full_fv) {
#pragma unroll
      for (int d = 0; d < 3; ++d) vals[nc++] = xs[d] * c_sg;
#pragma unroll
      for (int d = 0; d < 3; ++d) vals[nc++] = ns[d] * c_sg;
    }
    for (int ch = 0; ch < C; ++ch) {
      float x = active ? power_norm(vals[ch]) : 0.f;
      if (active) {
        if (p.flatten) out[(size_t)ch * V + g] = x;
        else out[(size_t)g * C + ch] = x;
      }
      // block-wide sum of squares for this channel
      float sq = x * x;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if ((tid & 31) == 0) warp_ss[(tid >> 5) * DPD_FV_CHANNELS_FULL + ch] = sq;
    }
    __syncthreads();
    if (tid < C) {   // deterministic: warps in order, Gaussian batches in order
      float acc = chan_ss[tid];
      for (int wv = 0; wv < FV_THREADS / 32; ++wv) acc += warp_ss[wv * DPD_FV_CHANNELS_FULL + tid];
      chan_ss[tid] = acc;
    }
  }
  __syncthreads();
  // L2-normalise each channel across the V Gaussians (tf.nn.l2_normalize(dim=1), :124-126)
  // the three groups (pi, mu, sigma) are normalised per channel independently, so per-channel it is.
  for (int i = tid; i < V * C; i += FV_THREADS) {
    const int ch = p.flatten ? (i / V) : (i % C);
    const float inv = 1.0f / sqrtf(fmaxf(chan_ss[ch], 1e-12f));
    out[i] = out[i] * inv;
  }
}

int fv_forward_dispatch(const FvParams& p, cudaStream_t st, bool* split_done) {
  if (split_done) *split_done = false;
  // DPD_FV_IMPL=old selects the previous one-role G = 8 kernel (A/B timing); default is the warp-specialised one
  const char* impl_env = getenv("DPD_FV_IMPL");
  const bool use_old = impl_env && !strcmp(impl_env, "old");
  int r = use_old ? fv_forward_optimized(p, st) : fv_forward_ws(p, st);
  if (r <= 0) {
    if (r == 0 && split_done) *split_done = p.fv_hi != nullptr && !p.flatten;
    return r;
  }
  size_t smem = (size_t)(9 * FV_PCHUNK * p.G + DPD_FV_CHANNELS_FULL * (1 + FV_THREADS / 32)) * sizeof(float);
  DPD_LAUNCH("fv_generic", st, fv_generic_kernel<<<p.n_clouds, FV_THREADS, smem, st>>>(p));
  DPD_CUDA_CHECK_LAUNCH("fv_generic_kernel");
  return 0;
}

}  // namespace dpd

extern "C" int dpd_fv_forward(const float* d_points, int n_clouds, int n_points, int G,
                              const float* h_centers, float sigma, int full_fv, int flatten,
                              float* d_fv, void* stream) {
  using namespace dpd;
  DPD_REQUIRE(d_points && d_fv && h_centers, DPD_E_INVALID, "dpd_fv_forward: null pointer");
  DPD_REQUIRE(n_clouds >= 0 && n_points > 0, DPD_E_INVALID, "dpd_fv_forward: bad sizes (%d clouds, %d points)", n_clouds, n_points);
  DPD_REQUIRE(G >= 2 && G <= DPD_MAX_GRID, DPD_E_UNSUPPORTED, "dpd_fv_forward: G=%d outside [2,%d]", G, DPD_MAX_GRID);
  DPD_REQUIRE(sigma > 0.f, DPD_E_INVALID, "dpd_fv_forward: sigma must be > 0");
  DPD_REQUIRE(aligned16(d_fv), DPD_E_INVALID, "dpd_fv_forward: d_fv must be 16-byte aligned");
  if (n_clouds == 0) return 0;
  FvParams p;
  fill_fv_params(p, d_points, n_clouds, n_points, G, h_centers, sigma, full_fv, flatten, d_fv);
  return fv_forward_dispatch(p, (cudaStream_t)stream, nullptr);
}
