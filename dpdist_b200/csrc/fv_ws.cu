// 3DmFV kernel for the reference default grid (G = 8, 512 Gaussians, full 20-channel FV), warp-specialised.
// Replaces get_3dmfv_tf (reference utils/dpdist_util.py:22-141) for that configuration.
//
// One persistent 384-thread CTA per SM runs two independent cloud pipelines ("slots"); per slot three roles of two
// warps each work on different clouds at the same time, handing buffers over through mbarriers:
//   builder      (warps 8-11)  point chunk c+1: the chunk's 768 bytes arrive by ONE bulk copy (cp.async.bulk +
//                              mbarrier complete_tx; prefetched one chunk ahead into a 4-deep ring); one thread per
//                              point evaluates the 3 x 8 per-axis responsibilities q, m = q z, s = q (z^2 - 1) (:54-75,
//                              separable form, oracle/fv_separable_np.py) into a double-buffered table
//   accumulator  (warps 0-3)   cloud c: thread (i0, i1) owns the 8 Gaussians of its z column: per point 8 LDS.128 and
//                              5 products, then 7 channels x 8 Gaussians of multiply / add / max / min (:78-109) with
//                              packed fp32x2 products and sums and FMNMX3 folding two points; 160 running statistics
//                              live in registers (setmaxnreg 208) and are dumped raw into a staging tile
//   finaliser    (warps 4-7)   cloud c-1: scale, signed square root (:118-121), per-channel L2 norm over the 512
//                              Gaussians (:124-126, fixed-order reductions: deterministic), float4 copy-out (and the
//                              scaled fp16 (hi, lo) copy the tensor-core head gathers from)
// so the table build and the normalisation / copy-out (40 % of the instructions of the previous one-role kernel, which
// ran them serialised at 1.4 instructions per clock) fill the issue slots the pair loop leaves free.
// Algorithmic HBM traffic: 4*(3N + 20*512) bytes per cloud (points read once, FV written once).
//
// 0/0 policy: the reference evaluates exp() unshifted, so a point farther than ~13 sigma from every
// Gaussian gives w_p = 0 for all of them and Q = 0/0 = NaN for the whole cloud (:73-74).  Here every axis is shifted
// by its smallest exponent before exp() (a softmax shift: identical values wherever the reference is finite), so such
// a point gets the limit responsibilities instead of NaN.  Tested against the fp64 twin of the oracle.
#include "fv.cuh"
#include <cuda_fp16.h>

namespace dpd {
namespace {

constexpr int G8 = 8, V8 = 512, C20 = 20;
constexpr int PC = 64;            // points per table chunk
constexpr int SLOTS = 2;          // cloud pipelines per CTA
constexpr int NT = 384;           // 4 accumulator + 4 finaliser + 4 builder warps
constexpr int TXP = G8 + 1;       // float4 pitch of the x / y tables (conflict-free builder stores)
constexpr int STAGE_F4 = V8 * 5 + 64;   // staging tile in float4 units: unit(g, c4) = 5 g + c4 + (g >> 3)
constexpr int PTS_RING = 4;
constexpr int REGS_ACC = 232, REGS_AUX = 96;

struct __align__(16) Tables {
  float4 tx[PC][TXP];             // x axis (<-> i1): {q, m, s, 0}
  float4 ty[PC][TXP];             // y axis (<-> i0)
  float4 qz[PC][3], mz[PC][3], sz[PC][3];   // z axis (<-> i2): 8 values as two float4 (+1 pad)
};

struct __align__(16) Slot {
  Tables tab[2];
  float4 stage[STAGE_F4];
  float pts[PTS_RING][PC * 3];
  float ss_part[2][C20][32];      // per-lane partial sums of squares of the two finaliser warps
  float4 inv_norm[5];
  unsigned long long pts_full[PTS_RING], tab_full[2], tab_empty[2], stage_full, stage_empty;
  int tab_np[2];
  int pad[2];
};

struct __align__(16) Smem {
  Slot slot[SLOTS];
};

typedef unsigned long long u64;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
// one bulk copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* src, unsigned bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  if (id == 1) asm volatile("bar.sync 1, 64;" ::: "memory"); else asm volatile("bar.sync 2, 64;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

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
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// sign(x)*sqrt(max(|x|,1e-12)), sign(0) = 0 (:118-121), hardware square root (sqrt.approx, ~1 ulp)
__device__ __forceinline__ float power_norm_fast(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(fabsf(x), 1e-12f)));
  return x == 0.f ? 0.f : copysignf(r, x);
}

// ------------------------------------------------------------------------------------------------------------------
// builder: per-axis tables of one point
// ------------------------------------------------------------------------------------------------------------------
struct AxisTab { float q[8], m[8], s[8]; };

__device__ __forceinline__ void build_axis(float x, const float (&c)[DPD_MAX_GRID], float inv_sigma, AxisTab& t) {
  float h[8], hmin = INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float z = (x - c[i]) * inv_sigma;
    t.m[i] = z;
    h[i] = z * z;
    hmin = fminf(hmin, h[i]);
  }
  float S = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // exp(-(z^2 - zmin^2)/2): the shift cancels in q = e / sum(e) and keeps the nearest cell at e = 1
    t.q[i] = ex2_approx((hmin - h[i]) * 0.72134752044448170368f);
    S += t.q[i];
  }
  const float inv = __fdividef(1.0f, S);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float z = t.m[i];
    t.q[i] *= inv;
    t.m[i] = t.q[i] * z;
    t.s[i] = t.q[i] * (h[i] - 1.0f);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// accumulator
// ------------------------------------------------------------------------------------------------------------------
struct Acc {
  float2 s[7][4];     // sums:   Q, mu xyz, sigma xyz  (4 Gaussian pairs)
  float2 mx[7][4];    // maxima
  float2 mn[6][4];    // minima: mu xyz, sigma xyz
};

struct PointTerms { float a, bx, by, cx, cy; };   // qy*qx, qy*mx, my*qx, qy*sx, sy*qx

__device__ __forceinline__ PointTerms point_terms(const Tables& T, int p, int i0, int i1) {
  const float4 X = T.tx[p][i1], Y = T.ty[p][i0];
  PointTerms t;
  t.a = Y.x * X.x; t.bx = Y.x * X.y; t.cx = Y.x * X.z; t.by = Y.y * X.x; t.cy = Y.z * X.x;
  return t;
}

// the 7 per-pair values of one point for one Gaussian pair: [Q, dmx, dmy, dmz, dsx, dsy, dsz]
__device__ __forceinline__ void pair_values(const PointTerms& t, float2 qz, float2 mz, float2 sz, float2 (&v)[7]) {
  v[0] = mul2(qz, bc(t.a));
  v[1] = mul2(qz, bc(t.bx));
  v[2] = mul2(qz, bc(t.by));
  v[3] = mul2(mz, bc(t.a));
  v[4] = mul2(qz, bc(t.cx));
  v[5] = mul2(qz, bc(t.cy));
  v[6] = mul2(sz, bc(t.a));
}

__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }

__device__ __forceinline__ void accumulate_chunk(const Tables& T, int np, int i0, int i1, Acc& acc) {
  int pp = 0;
  for (; pp + 1 < np; pp += 2) {
    const PointTerms t0 = point_terms(T, pp, i0, i1);
    const PointTerms t1 = point_terms(T, pp + 1, i0, i1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 q0 = T.qz[pp][h], m0 = T.mz[pp][h], s0 = T.sz[pp][h];
      const float4 q1 = T.qz[pp + 1][h], m1 = T.mz[pp + 1][h], s1 = T.sz[pp + 1][h];
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        const int J = 2 * h + jp;
        float2 v0[7], v1[7];
        pair_values(t0, jp ? hi2(q0) : lo2(q0), jp ? hi2(m0) : lo2(m0), jp ? hi2(s0) : lo2(s0), v0);
        pair_values(t1, jp ? hi2(q1) : lo2(q1), jp ? hi2(m1) : lo2(m1), jp ? hi2(s1) : lo2(s1), v1);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          acc.s[c][J] = add2(add2(acc.s[c][J], v0[c]), v1[c]);
          acc.mx[c][J].x = fmaxf(acc.mx[c][J].x, fmaxf(v0[c].x, v1[c].x));
          acc.mx[c][J].y = fmaxf(acc.mx[c][J].y, fmaxf(v0[c].y, v1[c].y));
          if (c > 0) {
            acc.mn[c - 1][J].x = fminf(acc.mn[c - 1][J].x, fminf(v0[c].x, v1[c].x));
            acc.mn[c - 1][J].y = fminf(acc.mn[c - 1][J].y, fminf(v0[c].y, v1[c].y));
          }
        }
      }
    }
  }
  if (pp < np) {   // odd tail
    const PointTerms t0 = point_terms(T, pp, i0, i1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 q0 = T.qz[pp][h], m0 = T.mz[pp][h], s0 = T.sz[pp][h];
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        const int J = 2 * h + jp;
        float2 v0[7];
        pair_values(t0, jp ? hi2(q0) : lo2(q0), jp ? hi2(m0) : lo2(m0), jp ? hi2(s0) : lo2(s0), v0);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          acc.s[c][J] = add2(acc.s[c][J], v0[c]);
          acc.mx[c][J].x = fmaxf(acc.mx[c][J].x, v0[c].x);
          acc.mx[c][J].y = fmaxf(acc.mx[c][J].y, v0[c].y);
          if (c > 0) {
            acc.mn[c - 1][J].x = fminf(acc.mn[c - 1][J].x, v0[c].x);
            acc.mn[c - 1][J].y = fminf(acc.mn[c - 1][J].y, v0[c].y);
          }
        }
      }
    }
  }
}

// Raw statistics of Gaussian pair J (Gaussians g, g+1) -> staging, in the output channel order (:134-137):
// [pi mean, pi max, mu mean xyz, mu max xyz, mu min xyz, sigma mean xyz, sigma max xyz, sigma min xyz]
__device__ __forceinline__ void dump_gaussian(float4* st, const Acc& a, int J, bool second) {
#define DPD_PICK(v) (second ? (v).y : (v).x)
  st[0] = make_float4(DPD_PICK(a.s[0][J]), DPD_PICK(a.mx[0][J]), DPD_PICK(a.s[1][J]), DPD_PICK(a.s[2][J]));
  st[1] = make_float4(DPD_PICK(a.s[3][J]), DPD_PICK(a.mx[1][J]), DPD_PICK(a.mx[2][J]), DPD_PICK(a.mx[3][J]));
  st[2] = make_float4(DPD_PICK(a.mn[0][J]), DPD_PICK(a.mn[1][J]), DPD_PICK(a.mn[2][J]), DPD_PICK(a.s[4][J]));
  st[3] = make_float4(DPD_PICK(a.s[5][J]), DPD_PICK(a.s[6][J]), DPD_PICK(a.mx[4][J]), DPD_PICK(a.mx[5][J]));
  st[4] = make_float4(DPD_PICK(a.mx[6][J]), DPD_PICK(a.mn[3][J]), DPD_PICK(a.mn[4][J]), DPD_PICK(a.mn[5][J]));
#undef DPD_PICK
}

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) fv_g8_ws_kernel(const FvParams p, const int use_bulk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int role = tid >> 7;                 // 0 accumulate, 1 finalise, 2 build  (one warpgroup each)
  const int slot_id = (tid >> 6) & 1;
  const int t = tid & 63;
  Slot& S = sm.slot[slot_id];
  const int N = p.N;
  const int nchunk = (N + PC - 1) / PC;
  const int slot_global = blockIdx.x * SLOTS + slot_id;
  const int slot_stride = gridDim.x * SLOTS;

  if (t == 0 && role == 0) {
    for (int i = 0; i < PTS_RING; ++i) mbar_init(&S.pts_full[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&S.tab_full[i], 2); mbar_init(&S.tab_empty[i], 2); }
    mbar_init(&S.stage_full, 2);
    mbar_init(&S.stage_empty, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (role == 2) {
    // ================================ builder ================================
    reg_dec<REGS_AUX>();
    const float inv_sigma = 1.0f / p.sigma;    // exact for the reference's power-of-two sigmas
    auto issue = [&](int k, int cloud, int chunk) {   // bulk copy of item k = (cloud, chunk) into ring slot k & 3
      const int n0 = chunk * PC, np = min(PC, N - n0);
      u64* bar = &S.pts_full[k & (PTS_RING - 1)];
      mbar_expect_tx(bar, (unsigned)np * 12u);
      bulk_load(S.pts[k & (PTS_RING - 1)], p.points + ((size_t)cloud * N + n0) * 3, (unsigned)np * 12u, bar);
    };
    int k = 0;
    if (use_bulk && t == 0 && slot_global < p.n_clouds) issue(0, slot_global, 0);
    for (int cloud = slot_global; cloud < p.n_clouds; cloud += slot_stride) {
      for (int chunk = 0; chunk < nchunk; ++chunk, ++k) {
        const int n0 = chunk * PC, np = min(PC, N - n0);
        float x = 0.f, y = 0.f, z = 0.f;
        if (use_bulk) {
          if (t == 0) {   // prefetch the next item; its ring slot was last read three items ago
            int nc = cloud, nk = chunk + 1;
            if (nk == nchunk) { nk = 0; nc += slot_stride; }
            if (nc < p.n_clouds) issue(k + 1, nc, nk);
          }
          mbar_wait(&S.pts_full[k & (PTS_RING - 1)], (k >> 2) & 1);
          if (t < np) {
            const float* src = &S.pts[k & (PTS_RING - 1)][t * 3];
            x = src[0]; y = src[1]; z = src[2];
          }
        } else if (t < np) {
          const float* src = p.points + ((size_t)cloud * N + n0 + t) * 3;
          x = __ldg(src); y = __ldg(src + 1); z = __ldg(src + 2);
        }
        mbar_wait(&S.tab_empty[k & 1], ((k >> 1) & 1) ^ 1);
        Tables& T = S.tab[k & 1];
        if (t < np) {
          AxisTab a;
          build_axis(x, p.c, inv_sigma, a);
#pragma unroll
          for (int i = 0; i < 8; ++i) T.tx[t][i] = make_float4(a.q[i], a.m[i], a.s[i], 0.f);
          build_axis(y, p.c, inv_sigma, a);
#pragma unroll
          for (int i = 0; i < 8; ++i) T.ty[t][i] = make_float4(a.q[i], a.m[i], a.s[i], 0.f);
          build_axis(z, p.c, inv_sigma, a);
          T.qz[t][0] = make_float4(a.q[0], a.q[1], a.q[2], a.q[3]); T.qz[t][1] = make_float4(a.q[4], a.q[5], a.q[6], a.q[7]);
          T.mz[t][0] = make_float4(a.m[0], a.m[1], a.m[2], a.m[3]); T.mz[t][1] = make_float4(a.m[4], a.m[5], a.m[6], a.m[7]);
          T.sz[t][0] = make_float4(a.s[0], a.s[1], a.s[2], a.s[3]); T.sz[t][1] = make_float4(a.s[4], a.s[5], a.s[6], a.s[7]);
        }
        if (t == 0) S.tab_np[k & 1] = np;
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.tab_full[k & 1]);
      }
    }
  } else if (role == 0) {
    // ================================ accumulator ================================
    reg_inc<REGS_ACC>();
    const int i0 = t >> 3, i1 = t & 7;
    int k = 0, it = 0;
    for (int cloud = slot_global; cloud < p.n_clouds; cloud += slot_stride, ++it) {
      Acc acc;
#pragma unroll
      for (int c = 0; c < 7; ++c)
#pragma unroll
        for (int J = 0; J < 4; ++J) {
          acc.s[c][J] = make_float2(0.f, 0.f);
          acc.mx[c][J] = make_float2(-INFINITY, -INFINITY);
          if (c < 6) acc.mn[c][J] = make_float2(INFINITY, INFINITY);
        }
      for (int chunk = 0; chunk < nchunk; ++chunk, ++k) {
        mbar_wait(&S.tab_full[k & 1], (k >> 1) & 1);
        accumulate_chunk(S.tab[k & 1], S.tab_np[k & 1], i0, i1, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.tab_empty[k & 1]);
      }
      mbar_wait(&S.stage_empty, (it & 1) ^ 1);
      // Gaussian g = t*8 + j at float4 unit 5 g + c4 + t (the + t skew spreads the 640-byte thread stride over banks)
      float4* st = S.stage + t * 41;
#pragma unroll
      for (int J = 0; J < 4; ++J) {
        dump_gaussian(st + (2 * J) * 5, acc, J, false);
        dump_gaussian(st + (2 * J + 1) * 5, acc, J, true);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.stage_full);
    }
  } else {
    // ================================ finaliser ================================
    reg_dec<REGS_AUX>();
    const float w = 1.0f / (float)V8;              // tf.ones/n_gaussians (:49)
    const float sqrt_w = sqrtf(w);
    const float c_pi = 1.0f / (sqrt_w * (float)N); // (:78)
    const float c_mu = 1.0f / sqrt_w;              // (:98)
    const float c_sg = 1.0f / sqrtf(2.0f * w);     // (:109)
    const float inv_n = 1.0f / (float)N;
    const int wslot = t >> 5;
    int it = 0;
    for (int cloud = slot_global; cloud < p.n_clouds; cloud += slot_stride, ++it) {
      mbar_wait(&S.stage_full, it & 1);
      // ---- pass 1: scale + signed square root in place, per-thread sums of squares
      float ss[C20];
#pragma unroll
      for (int c = 0; c < C20; ++c) ss[c] = 0.f;
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        const int g = t + 64 * i;
        float4* u = S.stage + 5 * g + (g >> 3);
        float4 r[5];
#pragma unroll
        for (int c4 = 0; c4 < 5; ++c4) r[c4] = u[c4];
        float* v = reinterpret_cast<float*>(r);
        v[0] = (v[0] * inv_n - w) * c_pi;
        v[1] = (v[1] - w) * c_pi;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          v[2 + d] = v[2 + d] * inv_n * c_mu;
          v[5 + d] *= c_mu;
          v[8 + d] *= c_mu;
          v[11 + d] = v[11 + d] * inv_n * c_sg;
          v[14 + d] *= c_sg;
          v[17 + d] *= c_sg;
        }
#pragma unroll
        for (int c = 0; c < C20; ++c) {
          v[c] = power_norm_fast(v[c]);
          ss[c] = fmaf(v[c], v[c], ss[c]);
        }
#pragma unroll
        for (int c4 = 0; c4 < 5; ++c4) u[c4] = r[c4];
      }
#pragma unroll
      for (int c = 0; c < C20; ++c) S.ss_part[wslot][c][lane] = ss[c];
      if (slot_id == 0) named_bar_sync(1, 64); else named_bar_sync(2, 64);
      // ---- per-channel L2 norm over the 512 Gaussians (tf.nn.l2_normalize(dim=1), :124-126), fixed order
      if (t < C20) {
        float tot = 0.f;
#pragma unroll
        for (int ww = 0; ww < 2; ++ww) {
          const float4* row = reinterpret_cast<const float4*>(S.ss_part[ww][t]);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) { const float4 x4 = row[q]; a0 += x4.x; a1 += x4.y; a2 += x4.z; a3 += x4.w; }
          tot += (a0 + a1) + (a2 + a3);
        }
        reinterpret_cast<float*>(S.inv_norm)[t] = rsqrtf(fmaxf(tot, 1e-12f));
      }
      if (slot_id == 0) named_bar_sync(1, 64); else named_bar_sync(2, 64);
      // ---- pass 2: copy-out, float4, coalesced
      float* out = p.fv + (size_t)cloud * V8 * C20;
      if (!p.flatten) {
        int g = t / 5, c4 = t - g * 5;
        const bool split = p.fv_hi != nullptr;
        const float sc = p.split_scale;
#pragma unroll 4
        for (int i = 0; i < 40; ++i) {
          const int F = t + 64 * i;
          const float4 nrm = S.inv_norm[c4];
          float4 v = S.stage[F + (g >> 3)];
          v.x *= nrm.x; v.y *= nrm.y; v.z *= nrm.z; v.w *= nrm.w;
          reinterpret_cast<float4*>(out)[F] = v;
          if (split) {   // scaled fp16 (hi, lo) copy for the tensor-core head
            const float a0 = v.x * sc, a1 = v.y * sc, a2 = v.z * sc, a3 = v.w * sc;
            const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
            const size_t e4 = (size_t)cloud * (V8 * C20 / 4) + F;
            uint2 uh, ul;
            uh.x = *reinterpret_cast<const unsigned*>(&h01); uh.y = *reinterpret_cast<const unsigned*>(&h23);
            ul.x = *reinterpret_cast<const unsigned*>(&l01); ul.y = *reinterpret_cast<const unsigned*>(&l23);
            reinterpret_cast<uint2*>(p.fv_hi)[e4] = uh;
            reinterpret_cast<uint2*>(p.fv_lo)[e4] = ul;
          }
          g += 12; c4 += 4;                      // F += 64 = 12 * 5 + 4
          if (c4 >= 5) { c4 -= 5; g += 1; }
        }
      } else {
        // channel-major [20, 512]: thread writes 4 consecutive Gaussians of one channel
        const float* stf = reinterpret_cast<const float*>(S.stage);
        const float* nrm = reinterpret_cast<const float*>(S.inv_norm);
        for (int e4 = t; e4 < V8 * C20 / 4; e4 += 64) {
          const int ch = e4 >> 7, g = (e4 & 127) * 4;
          const float s = nrm[ch];
          float4 v;
          v.x = stf[(5 * (g + 0) + ((g + 0) >> 3)) * 4 + ch] * s;
          v.y = stf[(5 * (g + 1) + ((g + 1) >> 3)) * 4 + ch] * s;
          v.z = stf[(5 * (g + 2) + ((g + 2) >> 3)) * 4 + ch] * s;
          v.w = stf[(5 * (g + 3) + ((g + 3) >> 3)) * 4 + ch] * s;
          reinterpret_cast<float4*>(out)[e4] = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.stage_empty);
    }
  }
}

}  // namespace

int fv_forward_ws(const FvParams& p, cudaStream_t stream) {
  if (p.G != G8 || !p.full_fv) return 1;
  static PerDeviceOnce attr_once;
  static int regs_ok = -1;
  if (attr_once.need()) {
    DPD_CUDA_CALL(cudaFuncSetAttribute(fv_g8_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  }
  if (regs_ok < 0) {
    // setmaxnreg.inc waits for registers the other warpgroups release: the launch allocation must cover the sum
    cudaFuncAttributes fa;
    DPD_CUDA_CALL(cudaFuncGetAttributes(&fa, fv_g8_ws_kernel));
    regs_ok = (fa.numRegs * NT >= 128 * REGS_ACC + 256 * REGS_AUX) ? 1 : 0;
  }
  if (!regs_ok) return set_error(DPD_E_UNSUPPORTED, "fv_g8_ws_kernel: register pool too small for its setmaxnreg plan");
  // one bulk copy per 64-point chunk needs 16-byte aligned sources and sizes: N % 4 == 0 and an aligned base
  const int use_bulk = (p.N % 4 == 0) && aligned16(p.points);
  const int want = (p.n_clouds + SLOTS - 1) / SLOTS;
  const int grid = want < num_sms() ? want : num_sms();
  DPD_LAUNCH("fv_g8_ws", stream, fv_g8_ws_kernel<<<grid, NT, sizeof(Smem), stream>>>(p, use_bulk));
  DPD_CUDA_CHECK_LAUNCH("fv_g8_ws_kernel");
  return 0;
}

}  // namespace dpd
