// 3DmFV kernel specialised for the reference default grid (G = 8, 512 Gaussians, full 20-channel FV).
// Replaces get_3dmfv_tf (reference utils/dpdist_util.py:22-141) for that configuration.
//
// One 128-thread CTA encodes one cloud at a time (grid-stride over clouds); work item = (cloud, 64-point chunk).
//   phase 0  the next item's points are prefetched into the other half of a double buffer with cp.async
//   phase 1  per-axis tables q, m = q*z, s = q*(z^2-1) in shared memory: one thread per (point, axis)
//            computes the 8 cells in registers (8 independent exps, no shuffles)
//   phase 2  thread (col = i0*8+i1, h) owns Gaussians (i0, i1, 4h..4h+3): per point 5 LDS.128, 5 products,
//            then 7 channels x 4 Gaussians of multiply / add / max / min.  Products and sums use the
//            packed fp32x2 instructions (FMUL2 / FADD2), max/min fold two points per FMNMX3.
//   phase 3  scale + power-normalise into a channel-major staging tile, per-channel L2 norm over the 512
//            Gaussians (fixed-order reductions, deterministic), coalesced float4 copy-out.
// Algorithmic HBM traffic: 4*(3N + 20*512) bytes per cloud (read points once, write the FV once).
#include "fv.cuh"
#include <cuda_fp16.h>

namespace dpd {
namespace {

constexpr int G8 = 8, V8 = 512, C20 = 20;
constexpr int T8 = 128;          // threads per CTA
constexpr int PC = 64;           // points per table chunk
constexpr int PITCH = V8 + 4;    // staging pitch (floats) per channel

struct __align__(16) Smem {
  union {
    struct {
      // row pitches of 9 / 3 float4 (not 8 / 2) so that the phase-1 stores of consecutive points,
      // which come from different lanes of one warp, fall into different 16-byte bank groups
      float4 tx[PC][G8 + 1];     // x axis (<-> i1): {q, m, s, 0}
      float4 ty[PC][G8 + 1];     // y axis (<-> i0)
      float4 qz[PC][3], mz[PC][3], sz[PC][3];   // z axis (<-> i2), 8 values as two float4 (+1 pad)
    } t;
    float stage[C20 * PITCH];    // [channel][gaussian], written after the tables are dead
  } u;
  float pts[2][PC * 3];          // double-buffered point chunks
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

// sign(x)*sqrt(max(|x|,1e-12)) with the hardware square root (sqrt.approx, ~1 ulp)
__device__ __forceinline__ float power_norm_fast(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(fabsf(x), 1e-12f)));
  return x == 0.f ? 0.f : copysignf(r, x);
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// per-thread running statistics for its 4 Gaussians (two packed pairs)
struct Acc {
  float2 s[7][2];     // sums:   Q, mu xyz, sigma xyz
  float2 mx[7][2];    // maxima
  float2 mn[6][2];    // minima: mu xyz, sigma xyz
};

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

// the 7 per-pair values of one point for Gaussian pair jp: [Q, dmx, dmy, dmz, dsx, dsy, dsz]
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
  const int nchunk = (N + PC - 1) / PC;
  const float w = 1.0f / (float)V8;              // tf.ones/n_gaussians (:49)
  const float sqrt_w = sqrtf(w);
  const float c_pi = 1.0f / (sqrt_w * (float)N); // (:78)
  const float c_mu = 1.0f / sqrt_w;              // (:98)
  const float c_sg = 1.0f / sqrtf(2.0f * w);     // (:109)
  const float inv_n = 1.0f / (float)N;
  const float inv_sigma = 1.0f / p.sigma;        // exact for the reference's power-of-two sigmas

  auto prefetch = [&](int cloud, int chunk, int buf) {
    const int n0 = chunk * PC, np = min(PC, N - n0);
    const float* src = p.points + ((size_t)cloud * N + n0) * 3;
    for (int i = tid; i < np * 3; i += T8) cp_async4(&sm.pts[buf][i], src + i);
    cp_async_commit();
  };

  int buf = 0;
  if ((int)blockIdx.x < p.n_clouds) prefetch(blockIdx.x, 0, 0);

  for (int cloud = blockIdx.x; cloud < p.n_clouds; cloud += gridDim.x) {
    Acc acc;
#pragma unroll
    for (int c = 0; c < 7; ++c)
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        acc.s[c][jp] = make_float2(0.f, 0.f);
        acc.mx[c][jp] = make_float2(-INFINITY, -INFINITY);
        if (c < 6) acc.mn[c][jp] = make_float2(INFINITY, INFINITY);
      }

    for (int chunk = 0; chunk < nchunk; ++chunk) {
      const int np = min(PC, N - chunk * PC);
      cp_async_wait_all();
      __syncthreads();   // this item's points have landed; previous tables / staging are dead
      {                  // phase 0: prefetch the next item into the other buffer
        int ncloud = cloud, nchk = chunk + 1;
        if (nchk == nchunk) { nchk = 0; ncloud += gridDim.x; }
        if (ncloud < p.n_clouds) prefetch(ncloud, nchk, buf ^ 1);
      }
      // ---- phase 1: tables, one thread per (point, axis); task index == offset into the chunk's xyz list
      for (int task = tid; task < np * 3; task += T8) {
        const int pp = task / 3, a = task - pp * 3;
        const float x = sm.pts[buf][task];
        float q[8], m[8], s[8];
        float S = 0.f, hmin = INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = (x - p.c[i]) * inv_sigma;
          m[i] = z;
          hmin = fminf(hmin, z * z);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {      // softmax shift: no 0/0 for far points (see fv_ws.cu)
          q[i] = __expf(-0.5f * (m[i] * m[i] - hmin));
          S += q[i];
        }
        const float inv = __fdividef(1.0f, S);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = m[i];
          q[i] *= inv;
          m[i] = q[i] * z;
          s[i] = q[i] * (z * z - 1.0f);
        }
        if (a == 2) {
          sm.u.t.qz[pp][0] = make_float4(q[0], q[1], q[2], q[3]); sm.u.t.qz[pp][1] = make_float4(q[4], q[5], q[6], q[7]);
          sm.u.t.mz[pp][0] = make_float4(m[0], m[1], m[2], m[3]); sm.u.t.mz[pp][1] = make_float4(m[4], m[5], m[6], m[7]);
          sm.u.t.sz[pp][0] = make_float4(s[0], s[1], s[2], s[3]); sm.u.t.sz[pp][1] = make_float4(s[4], s[5], s[6], s[7]);
        } else {
          float4* dst = (a == 0) ? sm.u.t.tx[pp] : sm.u.t.ty[pp];
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[i] = make_float4(q[i], m[i], s[i], 0.f);
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
      buf ^= 1;
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
    {
      float ss[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {      // the warp's five channels are loaded and reduced together
        const float4* row = reinterpret_cast<const float4*>(&sm.u.stage[(warp + 4 * j) * PITCH]);
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = row[lane + 32 * k];
        float s0 = v[0].x * v[0].x, s1 = v[1].x * v[1].x, s2 = v[2].x * v[2].x, s3 = v[3].x * v[3].x;
        s0 = fmaf(v[0].y, v[0].y, s0); s1 = fmaf(v[1].y, v[1].y, s1); s2 = fmaf(v[2].y, v[2].y, s2); s3 = fmaf(v[3].y, v[3].y, s3);
        s0 = fmaf(v[0].z, v[0].z, s0); s1 = fmaf(v[1].z, v[1].z, s1); s2 = fmaf(v[2].z, v[2].z, s2); s3 = fmaf(v[3].z, v[3].z, s3);
        s0 = fmaf(v[0].w, v[0].w, s0); s1 = fmaf(v[1].w, v[1].w, s1); s2 = fmaf(v[2].w, v[2].w, s2); s3 = fmaf(v[3].w, v[3].w, s3);
        ss[j] = (s0 + s1) + (s2 + s3);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < 5; ++j) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], o);
      if (lane < 5) {
        float sel = ss[0];
#pragma unroll
        for (int j = 1; j < 5; ++j) sel = (lane == j) ? ss[j] : sel;
        sm.inv_norm[warp + 4 * lane] = rsqrtf(fmaxf(sel, 1e-12f));
      }
    }
    __syncthreads();
    // ---- phase 3c: copy-out, float4, coalesced
    float* out = p.fv + (size_t)cloud * V8 * C20;
    if (p.flatten) {
#pragma unroll 4
      for (int e4 = tid; e4 < V8 * C20 / 4; e4 += T8) {
        const int ch = e4 >> 7, g = (e4 & 127) * 4;
        float4 v = *reinterpret_cast<const float4*>(&sm.u.stage[ch * PITCH + g]);
        const float s = sm.inv_norm[ch];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        reinterpret_cast<float4*>(out)[e4] = v;
      }
    } else if (tid < 125) {
      // thread = (g_local = tid/5, c4 = tid%5): fixed channel quad -> its 4 norms live in registers;
      // each iteration the CTA writes 125 consecutive float4 (25 Gaussians x 20 channels)
      const int gl = tid / 5, c4 = tid - gl * 5;
      const float n0 = sm.inv_norm[c4 * 4], n1 = sm.inv_norm[c4 * 4 + 1], n2 = sm.inv_norm[c4 * 4 + 2], n3 = sm.inv_norm[c4 * 4 + 3];
      const float* s0 = &sm.u.stage[(c4 * 4) * PITCH];
#pragma unroll 3
      for (int g = gl; g < V8; g += 25) {
        float4 v;
        v.x = s0[g] * n0; v.y = s0[PITCH + g] * n1; v.z = s0[2 * PITCH + g] * n2; v.w = s0[3 * PITCH + g] * n3;
        reinterpret_cast<float4*>(out)[g * 5 + c4] = v;
        if (p.fv_hi != nullptr) {   // scaled fp16 (hi, lo) copy for the tensor-core head
          const float sc = p.split_scale;
          const float a0 = v.x * sc, a1 = v.y * sc, a2 = v.z * sc, a3 = v.w * sc;
          const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
          // channel-split layout (fv.cuh): quads 0..3 -> X record of 16 channels, quad 4 -> Y record of 4 channels
          const size_t rec = (size_t)cloud * V8 + g;
          const size_t e4 = c4 < 4 ? rec * 4 + c4 : (size_t)(p.split_y_off / 4) + rec;
          uint2 uh, ul;
          uh.x = *reinterpret_cast<const unsigned*>(&h01); uh.y = *reinterpret_cast<const unsigned*>(&h23);
          ul.x = *reinterpret_cast<const unsigned*>(&l01); ul.y = *reinterpret_cast<const unsigned*>(&l23);
          reinterpret_cast<uint2*>(p.fv_hi)[e4] = uh;
          reinterpret_cast<uint2*>(p.fv_lo)[e4] = ul;
        }
      }
    }
  }
}

}  // namespace

int fv_forward_optimized(const FvParams& p, cudaStream_t stream) {
  if (p.G != G8 || !p.full_fv) return 1;
  static PerDeviceOnce attr_once;
  if (attr_once.need()) {
    DPD_CUDA_CALL(cudaFuncSetAttribute(fv_g8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  }
  const int max_ctas = num_sms() * 4;
  const int grid = p.n_clouds < max_ctas ? p.n_clouds : max_ctas;
  DPD_LAUNCH("fv_g8", stream, fv_g8_kernel<<<grid, T8, sizeof(Smem), stream>>>(p));
  DPD_CUDA_CHECK_LAUNCH("fv_g8_kernel");
  return 0;
}

}  // namespace dpd
