// 3DmFV kernel for the reference default grid (G = 8, 512 Gaussians, full 20-channel FV), warp-specialised.
// Replaces get_3dmfv_tf (reference utils/dpdist_util.py:22-141) for that configuration.
//
// One persistent 512-thread CTA per SM runs three independent cloud pipelines ("slots"); per slot three roles work on
// different clouds / point chunks at the same time, handing buffers over through mbarriers:
//   builder      (warp 12, serves all slots round-robin)  point chunk c+1: the chunk's 384 bytes arrive by ONE bulk copy
//                (cp.async.bulk + mbarrier complete_tx, prefetched one chunk ahead); one lane per point evaluates the
//                3 x 8 per-axis responsibilities q, m = q z, s = q (z^2 - 1) (:54-75, separable form,
//                oracle/fv_separable_np.py) into a double-buffered table of 32 points
//   accumulator  (warps 4 s .. 4 s + 3 of slot s)  chunk c: thread (i0, i1, h) owns Gaussians (i0, i1, 4h..4h+3): per point
//                5 LDS.128 and 5 products, then 7 channels x 4 Gaussians of multiply / add / max / min (:78-109) with
//                packed fp32x2 products and sums and FMNMX3 folding two points; the 80 running statistics live in
//                registers and are dumped raw into a staging tile when the cloud is complete
//   finaliser    (warp 13 + s)  cloud c-1: scale, signed square root (:118-121), per-channel L2 norm over the 512
//                Gaussians (:124-126, fixed-order reductions: deterministic), float4 copy-out (and the scaled fp16
//                (hi, lo) copy the tensor-core head gathers from)
// The previous one-role kernel ran these phases serialised and in lockstep on every CTA of an SM: its pair loop was half
// of the time (3.1 instructions per clock and SM) and the other half crawled at 1.4.  Here twelve accumulator warps (three
// per scheduler) keep the fp32 and max/min pipes busy while the four support warps fill the issue slots left over.
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
constexpr int PC = 32;            // points per table chunk (two chunks in flight per slot)
constexpr int SLOTS = 3;          // cloud pipelines per CTA
constexpr int ACC_WARPS = 4;      // accumulator warps per slot
constexpr int NT = 32 * (SLOTS * ACC_WARPS + 1 + SLOTS);   // 12 accumulator warps + 1 builder + 3 finalisers = 512
constexpr int TXP = G8 + 1;       // float4 pitch of the x / y tables (conflict-free builder stores)
constexpr int STAGE_F4 = V8 * 5 + 64;   // staging tile in float4 units: unit(g, c4) = 5 g + c4 + (g >> 3)
constexpr int PTS_RING = 4;

struct __align__(16) Tables {
  float4 tx[PC][TXP];             // x axis (<-> i1): {q, m, s, 0}
  float4 ty[PC][TXP];             // y axis (<-> i0)
  float4 qz[PC][3], mz[PC][3], sz[PC][3];   // z axis (<-> i2): 8 values as two float4 (+1 pad)
};

struct __align__(16) Slot {
  Tables tab[2];
  float4 stage[STAGE_F4];
  float pts[PTS_RING][PC * 3];
  float ss_part[C20][32];         // per-lane partial sums of squares of the finaliser warp
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
__device__ __forceinline__ unsigned mbar_try_wait(unsigned addr, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  return done;
}
// SLEEP_NS > 0: the waiting role is off the critical path (builder, finaliser) and must not burn the issue slots of the
// accumulator warp that shares its scheduler: back off between polls (a hot try_wait loop was 36 % of all instructions)
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  while (!mbar_try_wait(addr, parity)) {
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
  }
}
// one bulk copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* src, unsigned bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

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
  float2 s[7][2];     // sums:   Q, mu xyz, sigma xyz  (2 Gaussian pairs)
  float2 mx[7][2];    // maxima
  float2 mn[6][2];    // minima: mu xyz, sigma xyz
};

struct PointTerms {
  float a, bx, by, cx, cy;       // qy*qx, qy*mx, my*qx, qy*sx, sy*qx
  float4 qz, mz, sz;
};

__device__ __forceinline__ PointTerms load_terms(const Tables& T, int p, int i0, int i1, int h) {
  const float4 X = T.tx[p][i1], Y = T.ty[p][i0];
  PointTerms t;
  t.a = Y.x * X.x; t.bx = Y.x * X.y; t.cx = Y.x * X.z; t.by = Y.y * X.x; t.cy = Y.z * X.x;
  t.qz = T.qz[p][h]; t.mz = T.mz[p][h]; t.sz = T.sz[p][h];
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

__device__ __forceinline__ void accumulate_chunk(const Tables& T, int np, int i0, int i1, int h, Acc& acc) {
  int pp = 0;
  for (; pp + 1 < np; pp += 2) {
    const PointTerms t0 = load_terms(T, pp, i0, i1, h);
    const PointTerms t1 = load_terms(T, pp + 1, i0, i1, h);
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
    const PointTerms t0 = load_terms(T, pp, i0, i1, h);
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

// Scale constants of the statistics (:78, :98, :109) and the mean's 1/N.
struct Scales { float inv_n, w, c_pi, c_mu, c_sg; };

// Statistics of Gaussian pair J (Gaussians g, g+1) -> staging, scaled and power-normalised (:78-121), in the output
// channel order (:134-137): [pi mean, pi max, mu mean xyz, mu max xyz, mu min xyz, sigma mean xyz, sigma max xyz,
// sigma min xyz].  Done by the accumulator threads (four warps in parallel, values still in registers) so that the
// single finaliser warp of the slot only has the cross-Gaussian norm and the copy-out left.
__device__ __forceinline__ void dump_gaussian(float4* st, const Acc& a, int J, bool second, const Scales& k) {
#define DPD_PICK(v) (second ? (v).y : (v).x)
#define DPD_MEAN(v, c) power_norm_fast(DPD_PICK(v) * k.inv_n * (c))
#define DPD_EXT(v, c) power_norm_fast(DPD_PICK(v) * (c))
  st[0] = make_float4(power_norm_fast((DPD_PICK(a.s[0][J]) * k.inv_n - k.w) * k.c_pi), power_norm_fast((DPD_PICK(a.mx[0][J]) - k.w) * k.c_pi),
                      DPD_MEAN(a.s[1][J], k.c_mu), DPD_MEAN(a.s[2][J], k.c_mu));
  st[1] = make_float4(DPD_MEAN(a.s[3][J], k.c_mu), DPD_EXT(a.mx[1][J], k.c_mu), DPD_EXT(a.mx[2][J], k.c_mu), DPD_EXT(a.mx[3][J], k.c_mu));
  st[2] = make_float4(DPD_EXT(a.mn[0][J], k.c_mu), DPD_EXT(a.mn[1][J], k.c_mu), DPD_EXT(a.mn[2][J], k.c_mu), DPD_MEAN(a.s[4][J], k.c_sg));
  st[3] = make_float4(DPD_MEAN(a.s[5][J], k.c_sg), DPD_MEAN(a.s[6][J], k.c_sg), DPD_EXT(a.mx[4][J], k.c_sg), DPD_EXT(a.mx[5][J], k.c_sg));
  st[4] = make_float4(DPD_EXT(a.mx[6][J], k.c_sg), DPD_EXT(a.mn[3][J], k.c_sg), DPD_EXT(a.mn[4][J], k.c_sg), DPD_EXT(a.mn[5][J], k.c_sg));
#undef DPD_PICK
#undef DPD_MEAN
#undef DPD_EXT
}

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) fv_g8_ws_kernel(const FvParams p, const int use_bulk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.N;
  const int nchunk = (N + PC - 1) / PC;
  const int slot_stride = gridDim.x * SLOTS;

  if (tid < SLOTS) {
    Slot& S = sm.slot[tid];
    for (int i = 0; i < PTS_RING; ++i) mbar_init(&S.pts_full[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&S.tab_full[i], 1); mbar_init(&S.tab_empty[i], ACC_WARPS); }
    mbar_init(&S.stage_full, ACC_WARPS);
    mbar_init(&S.stage_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == SLOTS * ACC_WARPS) {
    // ================================ builder (one warp, all slots round-robin) ================================
    const float inv_sigma = 1.0f / p.sigma;    // exact for the reference's power-of-two sigmas
    auto issue = [&](Slot& S, int k, int cloud, int chunk) {   // bulk copy of item k = (cloud, chunk) into ring slot k & 3
      const int n0 = chunk * PC, np = min(PC, N - n0);
      u64* bar = &S.pts_full[k & (PTS_RING - 1)];
      mbar_expect_tx(bar, (unsigned)np * 12u);
      bulk_load(S.pts[k & (PTS_RING - 1)], p.points + ((size_t)cloud * N + n0) * 3, (unsigned)np * 12u, bar);
    };
    if (use_bulk && lane < SLOTS) {
      const int c0 = blockIdx.x * SLOTS + lane;
      if (c0 < p.n_clouds) issue(sm.slot[lane], 0, c0, 0);
    }
    __syncwarp();
    // item k of a slot = (its k / nchunk-th cloud, chunk k % nchunk); the slots' items are interleaved
    for (int k = 0;; ++k) {
      const int it = k / nchunk, chunk = k - it * nchunk;
      if (blockIdx.x * SLOTS + it * slot_stride >= p.n_clouds) break;        // slot 0 has the smallest cloud index
#pragma unroll 1
      for (int sl = 0; sl < SLOTS; ++sl) {
        const int cloud = blockIdx.x * SLOTS + sl + it * slot_stride;
        if (cloud >= p.n_clouds) break;
        Slot& S = sm.slot[sl];
        const int n0 = chunk * PC, np = min(PC, N - n0);
        float x = 0.f, y = 0.f, z = 0.f;
        if (use_bulk) {
          if (lane == 0) {   // prefetch this slot's next item
            int nc = cloud, nk = chunk + 1;
            if (nk == nchunk) { nk = 0; nc += slot_stride; }
            if (nc < p.n_clouds) issue(S, k + 1, nc, nk);
          }
          mbar_wait<40>(&S.pts_full[k & (PTS_RING - 1)], (k >> 2) & 1);
          if (lane < np) {
            const float* src = &S.pts[k & (PTS_RING - 1)][lane * 3];
            x = src[0]; y = src[1]; z = src[2];
          }
        } else if (lane < np) {
          const float* src = p.points + ((size_t)cloud * N + n0 + lane) * 3;
          x = __ldg(src); y = __ldg(src + 1); z = __ldg(src + 2);
        }
        mbar_wait<40>(&S.tab_empty[k & 1], ((k >> 1) & 1) ^ 1);
        Tables& T = S.tab[k & 1];
        if (lane < np) {
          AxisTab a;
          build_axis(x, p.c, inv_sigma, a);
#pragma unroll
          for (int i = 0; i < 8; ++i) T.tx[lane][i] = make_float4(a.q[i], a.m[i], a.s[i], 0.f);
          build_axis(y, p.c, inv_sigma, a);
#pragma unroll
          for (int i = 0; i < 8; ++i) T.ty[lane][i] = make_float4(a.q[i], a.m[i], a.s[i], 0.f);
          build_axis(z, p.c, inv_sigma, a);
          T.qz[lane][0] = make_float4(a.q[0], a.q[1], a.q[2], a.q[3]); T.qz[lane][1] = make_float4(a.q[4], a.q[5], a.q[6], a.q[7]);
          T.mz[lane][0] = make_float4(a.m[0], a.m[1], a.m[2], a.m[3]); T.mz[lane][1] = make_float4(a.m[4], a.m[5], a.m[6], a.m[7]);
          T.sz[lane][0] = make_float4(a.s[0], a.s[1], a.s[2], a.s[3]); T.sz[lane][1] = make_float4(a.s[4], a.s[5], a.s[6], a.s[7]);
        }
        if (lane == 0) S.tab_np[k & 1] = np;
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.tab_full[k & 1]);
      }
    }
  } else if (warp < SLOTS * ACC_WARPS) {
    // ================================ accumulator ================================
    const int slot_id = warp / ACC_WARPS;
    Slot& S = sm.slot[slot_id];
    const int t = tid - slot_id * (ACC_WARPS * 32);          // 0..127
    const int h = t >> 6, col = t & 63, i0 = col >> 3, i1 = col & 7;
    Scales sc;
    sc.w = 1.0f / (float)V8;                       // tf.ones/n_gaussians (:49)
    sc.c_pi = 1.0f / (sqrtf(sc.w) * (float)N);     // (:78)
    sc.c_mu = 1.0f / sqrtf(sc.w);                  // (:98)
    sc.c_sg = 1.0f / sqrtf(2.0f * sc.w);           // (:109)
    sc.inv_n = 1.0f / (float)N;
    int k = 0, it = 0;
    for (int cloud = blockIdx.x * SLOTS + slot_id; cloud < p.n_clouds; cloud += slot_stride, ++it) {
      Acc acc;
#pragma unroll
      for (int c = 0; c < 7; ++c)
#pragma unroll
        for (int J = 0; J < 2; ++J) {
          acc.s[c][J] = make_float2(0.f, 0.f);
          acc.mx[c][J] = make_float2(-INFINITY, -INFINITY);
          if (c < 6) acc.mn[c][J] = make_float2(INFINITY, INFINITY);
        }
      for (int chunk = 0; chunk < nchunk; ++chunk, ++k) {
        mbar_wait<0>(&S.tab_full[k & 1], (k >> 1) & 1);
        accumulate_chunk(S.tab[k & 1], S.tab_np[k & 1], i0, i1, h, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.tab_empty[k & 1]);
      }
      mbar_wait<0>(&S.stage_empty, (it & 1) ^ 1);
      // Gaussian g = col*8 + 4h + j at float4 unit 5 g + c4 + col (the + col skew spreads the 640-byte column stride
      // over the banks)
      float4* st = S.stage + col * 41 + h * 20;
#pragma unroll
      for (int J = 0; J < 2; ++J) {
        dump_gaussian(st + (2 * J) * 5, acc, J, false, sc);
        dump_gaussian(st + (2 * J + 1) * 5, acc, J, true, sc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.stage_full);
    }
  } else {
    // ================================ finaliser (one warp per slot) ================================
    const int slot_id = warp - SLOTS * ACC_WARPS - 1;
    Slot& S = sm.slot[slot_id];
    int it = 0;
    for (int cloud = blockIdx.x * SLOTS + slot_id; cloud < p.n_clouds; cloud += slot_stride, ++it) {
      mbar_wait<40>(&S.stage_full, it & 1);
      // ---- pass 1: per-lane sums of squares of the (already power-normalised) statistics
      float ss[C20];
#pragma unroll
      for (int c = 0; c < C20; ++c) ss[c] = 0.f;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int g = lane + 32 * i;
        const float4* u = S.stage + 5 * g + (g >> 3);
#pragma unroll
        for (int c4 = 0; c4 < 5; ++c4) {
          const float4 r = u[c4];
          ss[4 * c4 + 0] = fmaf(r.x, r.x, ss[4 * c4 + 0]);
          ss[4 * c4 + 1] = fmaf(r.y, r.y, ss[4 * c4 + 1]);
          ss[4 * c4 + 2] = fmaf(r.z, r.z, ss[4 * c4 + 2]);
          ss[4 * c4 + 3] = fmaf(r.w, r.w, ss[4 * c4 + 3]);
        }
      }
#pragma unroll
      for (int c = 0; c < C20; ++c) S.ss_part[c][lane] = ss[c];
      __syncwarp();
      // ---- per-channel L2 norm over the 512 Gaussians (tf.nn.l2_normalize(dim=1), :124-126), fixed order
      if (lane < C20) {
        const float4* row = reinterpret_cast<const float4*>(S.ss_part[lane]);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) { const float4 x4 = row[q]; a0 += x4.x; a1 += x4.y; a2 += x4.z; a3 += x4.w; }
        reinterpret_cast<float*>(S.inv_norm)[lane] = rsqrtf(fmaxf((a0 + a1) + (a2 + a3), 1e-12f));
      }
      __syncwarp();
      // ---- pass 2: copy-out, float4, coalesced
      float* out = p.fv + (size_t)cloud * V8 * C20;
      if (!p.flatten) {
        int g = lane / 5, c4 = lane - g * 5;
        const bool split = p.fv_hi != nullptr;
        const float sc = p.split_scale;
#pragma unroll 4
        for (int i = 0; i < 80; ++i) {
          const int F = lane + 32 * i;
          const float4 nrm = S.inv_norm[c4];
          float4 v = S.stage[F + (g >> 3)];
          v.x *= nrm.x; v.y *= nrm.y; v.z *= nrm.z; v.w *= nrm.w;
          reinterpret_cast<float4*>(out)[F] = v;
          if (split) {   // scaled fp16 (hi, lo) copy for the tensor-core head
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
          g += 6; c4 += 2;                       // F += 32 = 6 * 5 + 2
          if (c4 >= 5) { c4 -= 5; g += 1; }
        }
      } else {
        // channel-major [20, 512]: a lane writes 4 consecutive Gaussians of one channel
        const float* stf = reinterpret_cast<const float*>(S.stage);
        const float* nrm = reinterpret_cast<const float*>(S.inv_norm);
        for (int e4 = lane; e4 < V8 * C20 / 4; e4 += 32) {
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
  if (attr_once.need()) {
    DPD_CUDA_CALL(cudaFuncSetAttribute(fv_g8_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  }
  // one bulk copy per 32-point chunk needs 16-byte aligned sources and sizes: N % 4 == 0 and an aligned base
  const int use_bulk = (p.N % 4 == 0) && aligned16(p.points);
  const int want = (p.n_clouds + SLOTS - 1) / SLOTS;
  const int grid = want < num_sms() ? want : num_sms();
  DPD_LAUNCH("fv_g8_ws", stream, fv_g8_ws_kernel<<<grid, NT, sizeof(Smem), stream>>>(p, use_bulk));
  DPD_CUDA_CHECK_LAUNCH("fv_g8_ws_kernel");
  return 0;
}

}  // namespace dpd
