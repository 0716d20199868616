// 3DmFV kernel specialised for the reference default grid (G = 8, 512 Gaussians, full 20-channel FV).
// Replaces get_3dmfv_tf (reference utils/dpdist_util.py:22-141) for that configuration.
//
// One 128-thread CTA encodes one cloud at a time (grid-stride over clouds).
//   phase 1  per 64-point chunk: per-axis tables q, m = q*z, s = q*(z^2-1) in shared memory
//            (3*8 exps per point; 8 lanes cooperate on one (point, axis) and reduce with shuffles)
//   phase 2  thread (col = i0*8+i1, h) owns Gaussians (i0, i1, 4h..4h+3): per point 5 LDS.128, 5 products,
//            then 7 channels x 4 Gaussians of multiply / add / max / min.  Products and sums use the
//            packed fp32x2 instructions (FMUL2 / FADD2), max/min fold two points per FMNMX3.
//   phase 3  scale + power-normalise into a channel-major staging tile, per-channel L2 norm over the 512
//            Gaussians (fixed-order warp reductions), coalesced float4 copy-out.
// Algorithmic HBM traffic: 4*(3N + 20*512) bytes per cloud (read points once, write the FV once).
#include "fv.cuh"

namespace dpd {
namespace {

constexpr int G8 = 8, V8 = 512, C20 = 20;
constexpr int T8 = 128;          // threads per CTA
constexpr int PC = 64;           // points per table chunk
constexpr int PITCH = V8 + 4;    // staging pitch (floats) per channel

struct __align__(16) Smem {
  union {
    struct {
      float4 tx[PC][G8];         // x axis (<-> i1): {q, m, s, 0}
      float4 ty[PC][G8];         // y axis (<-> i0)
      float4 qz[PC][2], mz[PC][2], sz[PC][2];   // z axis (<-> i2), 8 values as two float4
    } t;
    float stage[C20 * PITCH];    // [channel][gaussian], written after the tables are dead
  } u;
  float pts[PC * 3];
  float inv_norm[C20];
};

typedef unsigned long long u64;
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<u64*>(&d)) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<u64*>(&d)) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)));
  return d;
}
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }

// sign(x)*sqrt(max(|x|,1e-12)) with the 1-ulp hardware square root (sqrt.approx): 5 instructions instead
// of ~25 for the IEEE sqrtf; 80 of these per thread per cloud.
__device__ __forceinline__ float power_norm_fast(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(fabsf(x), 1e-12f)));
  return x == 0.f ? 0.f : copysignf(r, x);
}

// per-thread running statistics for its 4 Gaussians (two packed pairs)
struct Acc {
  float2 s[7][2];     // sums:   Q, mu xyz, sigma xyz
  float2 mx[7][2];    // maxima
  float2 mn[6][2];    // minima: mu xyz, sigma xyz
};

// the 7 per-pair values of one point for Gaussian pair jp: [Q, dmx, dmy, dmz, dsx, dsy, dsz]
struct PointTerms {
  float a, bx, by, cx, cy;       // qy*qx, qy*mx, my*qx, qy*sx, sy*qx
  float4 qz, mz, sz;
};

__device__ __forceinline__ PointTerms load_terms(const Smem& sm, int p, int i0, int i1, int h) {
  const float4 X = sm.u.t.tx[p][i1], Y = sm.u.t.ty[p][i0];
  PointTerms t;
  t.a = Y.x * X.x; t.bx = Y.x * X.y; t.cx = Y.x * X.z; t.by = Y.y * X.x; t.cy = Y.z * X.x;
  t.qz = sm.u.t.qz[p][h]; t.mz = sm.u.t.mz[p][h]; t.sz = sm.u.t.sz[p][h];
  return t;
}

__device__ __forceinline__ void pair_values(const PointTerms& t, int jp, float2 (&v)[7]) {
  const float2 qz = jp ? make_float2(t.qz.z, t.qz.w) : make_float2(t.qz.x, t.qz.y);
  const float2 mz = jp ? make_float2(t.mz.z, t.mz.w) : make_float2(t.mz.x, t.mz.y);
  const float2 sz = jp ? make_float2(t.sz.z, t.sz.w) : make_float2(t.sz.x, t.sz.y);
  v[0] = mul2(qz, bc(t.a));
  v[1] = mul2(qz, bc(t.bx));
  v[2] = mul2(qz, bc(t.by));
  v[3] = mul2(mz, bc(t.a));
  v[4] = mul2(qz, bc(t.cx));
  v[5] = mul2(qz, bc(t.cy));
  v[6] = mul2(sz, bc(t.a));
}

__global__ void __launch_bounds__(T8) fv_g8_kernel(const FvParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = tid >> 6, col = tid & 63, i0 = col >> 3, i1 = col & 7;
  const int N = p.N;
  const float ci = p.c[tid & 7];                 // this thread's table column in phase 1
  const float w = 1.0f / (float)V8;              // tf.ones/n_gaussians (:49)
  const float sqrt_w = sqrtf(w);
  const float c_pi = 1.0f / (sqrt_w * (float)N); // (:78)
  const float c_mu = 1.0f / sqrt_w;              // (:98)
  const float c_sg = 1.0f / sqrtf(2.0f * w);     // (:109)
  const float inv_n = 1.0f / (float)N;
  const float inv_sigma = 1.0f / p.sigma;        // exact for the reference's power-of-two sigmas

  for (int cloud = blockIdx.x; cloud < p.n_clouds; cloud += gridDim.x) {
    const float* pts = p.points + (size_t)cloud * N * 3;
    Acc acc;
#pragma unroll
    for (int c = 0; c < 7; ++c)
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        acc.s[c][jp] = make_float2(0.f, 0.f);
        acc.mx[c][jp] = make_float2(-INFINITY, -INFINITY);
        if (c < 6) acc.mn[c][jp] = make_float2(INFINITY, INFINITY);
      }

    for (int n0 = 0; n0 < N; n0 += PC) {
      const int np = min(PC, N - n0);
      __syncthreads();   // previous chunk's tables / previous cloud's staging are dead
      for (int i = tid; i < np * 3; i += T8) sm.pts[i] = pts[(size_t)n0 * 3 + i];
      __syncthreads();
      // ---- phase 1: tables.  task = (point, axis, cell); 8 consecutive lanes share (point, axis)
      const int ntask = np * 24;
      for (int t0 = 0; t0 < ntask; t0 += T8) {
        const int t = t0 + tid;
        const bool ok = t < ntask;
        const int pa = ok ? (t >> 3) : 0;
        const int pp = pa / 3, a = pa - pp * 3;
        const float x = sm.pts[pp * 3 + a];
        const float z = (x - ci) * inv_sigma;
        const float e = __expf(-0.5f * z * z);
        float S = e;
        S += __shfl_xor_sync(0xffffffffu, S, 1);
        S += __shfl_xor_sync(0xffffffffu, S, 2);
        S += __shfl_xor_sync(0xffffffffu, S, 4);
        const float q = __fdividef(e, S);
        const float m = q * z, s = q * (z * z - 1.0f);
        if (ok) {
          const int i = tid & 7;
          if (a == 0) sm.u.t.tx[pp][i] = make_float4(q, m, s, 0.f);
          else if (a == 1) sm.u.t.ty[pp][i] = make_float4(q, m, s, 0.f);
          else {
            reinterpret_cast<float*>(&sm.u.t.qz[pp][0])[i] = q;
            reinterpret_cast<float*>(&sm.u.t.mz[pp][0])[i] = m;
            reinterpret_cast<float*>(&sm.u.t.sz[pp][0])[i] = s;
          }
        }
      }
      __syncthreads();
      // ---- phase 2: accumulate, two points per iteration
      int pp = 0;
      for (; pp + 1 < np; pp += 2) {
        const PointTerms t0 = load_terms(sm, pp, i0, i1, h);
        const PointTerms t1 = load_terms(sm, pp + 1, i0, i1, h);
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
          float2 v0[7], v1[7];
          pair_values(t0, jp, v0);
          pair_values(t1, jp, v1);
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            acc.s[c][jp] = add2(add2(acc.s[c][jp], v0[c]), v1[c]);
            acc.mx[c][jp].x = fmaxf(acc.mx[c][jp].x, fmaxf(v0[c].x, v1[c].x));
            acc.mx[c][jp].y = fmaxf(acc.mx[c][jp].y, fmaxf(v0[c].y, v1[c].y));
            if (c > 0) {
              acc.mn[c - 1][jp].x = fminf(acc.mn[c - 1][jp].x, fminf(v0[c].x, v1[c].x));
              acc.mn[c - 1][jp].y = fminf(acc.mn[c - 1][jp].y, fminf(v0[c].y, v1[c].y));
            }
          }
        }
      }
      if (pp < np) {   // odd tail
        const PointTerms t0 = load_terms(sm, pp, i0, i1, h);
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
          float2 v0[7];
          pair_values(t0, jp, v0);
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            acc.s[c][jp] = add2(acc.s[c][jp], v0[c]);
            acc.mx[c][jp].x = fmaxf(acc.mx[c][jp].x, v0[c].x);
            acc.mx[c][jp].y = fmaxf(acc.mx[c][jp].y, v0[c].y);
            if (c > 0) {
              acc.mn[c - 1][jp].x = fminf(acc.mn[c - 1][jp].x, v0[c].x);
              acc.mn[c - 1][jp].y = fminf(acc.mn[c - 1][jp].y, v0[c].y);
            }
          }
        }
      }
    }
    __syncthreads();   // tables dead -> staging
    // ---- phase 3a: scale, power-normalise, stage channel-major.  Output channel order (:134-137):
    // [pi mean, pi max, mu mean xyz, mu max xyz, mu min xyz, sigma mean xyz, sigma max xyz, sigma min xyz]
    {
      float* st = sm.u.stage + col * 8 + h * 4;
      auto put = [&](int ch, float2 a, float2 b, float scale, float bias_w) {
        float4 o;
        o.x = power_norm_fast((a.x - bias_w) * scale); o.y = power_norm_fast((a.y - bias_w) * scale);
        o.z = power_norm_fast((b.x - bias_w) * scale); o.w = power_norm_fast((b.y - bias_w) * scale);
        *reinterpret_cast<float4*>(st + ch * PITCH) = o;
      };
      auto mean2 = [&](float2 a) { return make_float2(a.x * inv_n, a.y * inv_n); };
      put(0, mean2(acc.s[0][0]), mean2(acc.s[0][1]), c_pi, w);
      put(1, acc.mx[0][0], acc.mx[0][1], c_pi, w);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        put(2 + d, mean2(acc.s[1 + d][0]), mean2(acc.s[1 + d][1]), c_mu, 0.f);
        put(5 + d, acc.mx[1 + d][0], acc.mx[1 + d][1], c_mu, 0.f);
        put(8 + d, acc.mn[d][0], acc.mn[d][1], c_mu, 0.f);
        put(11 + d, mean2(acc.s[4 + d][0]), mean2(acc.s[4 + d][1]), c_sg, 0.f);
        put(14 + d, acc.mx[4 + d][0], acc.mx[4 + d][1], c_sg, 0.f);
        put(17 + d, acc.mn[3 + d][0], acc.mn[3 + d][1], c_sg, 0.f);
      }
    }
    __syncthreads();
    // ---- phase 3b: per-channel L2 norm over the 512 Gaussians (tf.nn.l2_normalize(dim=1), :124-126)
    for (int ch = warp; ch < C20; ch += T8 / 32) {
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < V8 / 32; ++k) {
        const float x = sm.u.stage[ch * PITCH + lane + 32 * k];
        ss = fmaf(x, x, ss);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) sm.inv_norm[ch] = rsqrtf(fmaxf(ss, 1e-12f));
    }
    __syncthreads();
    // ---- phase 3c: copy-out, float4, coalesced
    float* out = p.fv + (size_t)cloud * V8 * C20;
    if (p.flatten) {
      for (int e4 = tid; e4 < V8 * C20 / 4; e4 += T8) {
        const int ch = e4 / (V8 / 4), g = (e4 - ch * (V8 / 4)) * 4;
        float4 v = *reinterpret_cast<const float4*>(&sm.u.stage[ch * PITCH + g]);
        const float s = sm.inv_norm[ch];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        reinterpret_cast<float4*>(out)[e4] = v;
      }
    } else {
      for (int e4 = tid; e4 < V8 * C20 / 4; e4 += T8) {
        const int g = e4 / (C20 / 4), ch = (e4 - g * (C20 / 4)) * 4;
        float4 v;
        v.x = sm.u.stage[(ch + 0) * PITCH + g] * sm.inv_norm[ch + 0];
        v.y = sm.u.stage[(ch + 1) * PITCH + g] * sm.inv_norm[ch + 1];
        v.z = sm.u.stage[(ch + 2) * PITCH + g] * sm.inv_norm[ch + 2];
        v.w = sm.u.stage[(ch + 3) * PITCH + g] * sm.inv_norm[ch + 3];
        reinterpret_cast<float4*>(out)[e4] = v;
      }
    }
  }
}

}  // namespace

int fv_forward_optimized(const FvParams& p, cudaStream_t stream) {
  if (p.G != G8 || !p.full_fv) return 1;
  static bool attr_done = false;
  if (!attr_done) {
    DPD_CUDA_CALL(cudaFuncSetAttribute(fv_g8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    attr_done = true;
  }
  const int max_ctas = num_sms() * 4;
  const int grid = p.n_clouds < max_ctas ? p.n_clouds : max_ctas;
  DPD_LAUNCH("fv_g8", stream, fv_g8_kernel<<<grid, T8, sizeof(Smem), stream>>>(p));
  DPD_CUDA_CHECK_LAUNCH("fv_g8_kernel");
  return 0;
}

}  // namespace dpd
